p='/root/repo/srsran_b200/csrc/map_f16.cuh'
s=open(p).read()
def rep(a,b,cnt=1):
    global s
    c=s.count(a)
    assert c==cnt,(a,c)
    s=s.replace(a,b)
rep("""template <int T>
struct F16Lay {""","""// L2 eviction-priority policies: the LLR planes are streamed (each row is read once per pass, far apart), the beta
// checkpoints are written and read back within one launch -> keep those in L2, let the stream pass through
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_tile4_hint(unsigned dst_s, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar_s, uint64_t pol)
{
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;\\n" ::"r"(dst_s),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_s), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s, uint64_t pol)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\\n" ::"r"(dst_s), "l"(src),
               "r"(bytes), "r"(bar_s), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg128_hint(void* p, uint4 v, uint64_t pol)
{
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;\\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

template <int T, int STAGES>
struct F16Lay {""")
rep("  static constexpr int kStages     = 2;","  static constexpr int kStages     = STAGES;")
rep("""template <int N, int MODE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_map_f16(const MapArgs a)""","""template <int N, int MODE, int NT, int MINB, int STAGES>
__global__ void __launch_bounds__(NT, MINB) k_map_f16(const MapArgs a)""")
rep("  using Lay = F16Lay<T>;","  using Lay = F16Lay<T, STAGES>;\n  constexpr int kStages = STAGES;")
old_issue = s[s.index("  auto bar_of = [&](int idx)"):s.index("  // x (systematic + a-priori) and y (parity) of box row i")]
new_issue = """  const bool     hints     = (a.mode & 0x100) == 0; // (bit 8 of mode: L2 hints off, for A/B measurements)
  const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
  auto bar_of = [&](int stage) -> unsigned { return sm_s + 4u * (unsigned)(Lay::kBarOff + 2 * stage); };
  int wr_idx = 0, wr_stage = 0; // next tile of the sequence to request / its stage
  auto issue = [&]() {
    __syncwarp(); // every lane is done with the stage about to be refilled
    if (lane == 0 && wr_idx < n_seq) {
      const unsigned bar = bar_of(wr_stage);
      const unsigned dst = sm_s + 4u * (unsigned)(wr_stage * Lay::kStageWords);
      int            t;
      bool           aux = false;
      if (wr_idx < s1)
        t = 4 - wr_idx;
      else if (wr_idx < s2)
        t = nT - 1 - (wr_idx - s1);
      else if (wr_idx < s3)
        t = a0 + (wr_idx - s2);
      else {
        t   = wr_idx - s3;
        aux = true;
      }
      const int      r1        = (8 * t + 8) < W ? 8 : W - 8 * t;
      const unsigned lut_bytes = (unsigned)r1 * T * 4u;
      mbar_expect_tx(bar, aux ? kBoxBytes + lut_bytes + 1024u : kBoxBytes);
      if (hints)
        tma_tile4_hint(dst, tmap, 0, blk0, 8 * t, plane0, bar, pol_first);
      else
        tma_tile4(dst, tmap, 0, blk0, 8 * t, plane0, bar);
      if (aux) {
        bulk_g2s(dst + 4u * (unsigned)Lay::kLutOff, lut + (size_t)8 * t * T, lut_bytes, bar);
        if (hints)
          bulk_g2s_hint(dst + 4u * (unsigned)Lay::kCkOff, ck_warp + (size_t)(t + 1) * ck_stride, 1024u, bar, pol_first);
        else
          bulk_g2s(dst + 4u * (unsigned)Lay::kCkOff, ck_warp + (size_t)(t + 1) * ck_stride, 1024u, bar);
      }
    }
    wr_idx++;
    wr_stage = wr_stage + 1 == kStages ? 0 : wr_stage + 1;
  };
  int      rd_stage = 0;
  unsigned rd_phase = 0; // bit s = parity the consumer waits for on stage s
  // acquire(): keep kStages - 1 tiles in flight behind the one being consumed, wait for the next one, return its lane view
  auto acquire = [&]() -> const u32* {
    issue();
    mbar_wait(bar_of(rd_stage), (rd_phase >> rd_stage) & 1u);
    rd_phase ^= 1u << rd_stage;
    const u32* tb = my + rd_stage * Lay::kStageWords;
    rd_stage      = rd_stage + 1 == kStages ? 0 : rd_stage + 1;
    return tb;
  };
"""
s=s.replace(old_issue,new_issue)
rep("""  if (lane == 0) {
    mbar_init(bar_of(0), 1);
    mbar_init(bar_of(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\\n" ::: "memory");
  }
  __syncwarp();
  issue(0);
""","""  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kStages; i++)
      mbar_init(bar_of(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\\n" ::: "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kStages - 1; i++)
    issue();
""")
rep("""  auto ck_store = [&](int sl, const u32 (&v)[8]) {
    if (live) {
      uint4* g = reinterpret_cast<uint4*>(ck_warp + (size_t)sl * ck_stride) + lane;
      g[0]     = make_uint4(v[0], v[1], v[2], v[3]);
      g[32]    = make_uint4(v[4], v[5], v[6], v[7]);
    }
  };""","""  auto ck_store = [&](int sl, const u32 (&v)[8]) {
    if (live) {
      uint4* g = reinterpret_cast<uint4*>(ck_warp + (size_t)sl * ck_stride) + lane;
      if (hints) {
        stg128_hint(g, make_uint4(v[0], v[1], v[2], v[3]), pol_last);
        stg128_hint(g + 32, make_uint4(v[4], v[5], v[6], v[7]), pol_last);
      } else {
        g[0]  = make_uint4(v[0], v[1], v[2], v[3]);
        g[32] = make_uint4(v[4], v[5], v[6], v[7]);
      }
    }
  };""")
open(p,'w').write(s)
p='/root/repo/srsran_b200/csrc/engine.cu'
s=open(p).read()
rep("""template <int N, int NT, int MINB>
static cudaError_t launch_map_f16(MapArgs a, int n_slots, uint32_t n_iter, cudaStream_t st)
{
  constexpr int T = N / 2, G = 32 / T;
  const int     warps  = (n_slots + G - 1) / G;
  const int     blocks = (warps + (NT / 32) - 1) / (NT / 32);
  const size_t  smem   = (size_t)(NT / 32) * F16Lay<T>::kWarpWords * 4;
  void (*kern)(const MapArgs) = (n_iter & 1) ? k_map_f16<N, 2, NT, MINB> : (n_iter ? k_map_f16<N, 1, NT, MINB> : k_map_f16<N, 0, NT, MINB>);""","""template <int N, int NT, int MINB, int STAGES>
static cudaError_t launch_map_f16(MapArgs a, int n_slots, uint32_t n_iter, cudaStream_t st)
{
  constexpr int T = N / 2, G = 32 / T;
  const int     warps  = (n_slots + G - 1) / G;
  const int     blocks = (warps + (NT / 32) - 1) / (NT / 32);
  const size_t  smem   = (size_t)(NT / 32) * F16Lay<T, STAGES>::kWarpWords * 4;
  void (*kern)(const MapArgs) = (n_iter & 1) ? k_map_f16<N, 2, NT, MINB, STAGES>
                                             : (n_iter ? k_map_f16<N, 1, NT, MINB, STAGES> : k_map_f16<N, 0, NT, MINB, STAGES>);""")
rep("""        else if (opt_map_cfg == 20)
          e = launch_map_f16<16, 256, 2>(a, cls[c].n_slots, p.iter0 + it, stream);
        else if (opt_map_cfg == 21)
          e = launch_map_f16<16, 128, 4>(a, cls[c].n_slots, p.iter0 + it, stream);
        else if (opt_map_cfg == 22)
          e = launch_map_f16<16, 256, 1>(a, cls[c].n_slots, p.iter0 + it, stream);""","""        else if (opt_map_cfg >= 20) {
          if (opt_map_cfg & 0x100)
            a.mode |= 0x100;
          switch (opt_map_cfg & 0xff) {
            case 20: e = launch_map_f16<16, 256, 2, 2>(a, cls[c].n_slots, p.iter0 + it, stream); break;
            case 21: e = launch_map_f16<16, 128, 3, 3>(a, cls[c].n_slots, p.iter0 + it, stream); break;
            case 22: e = launch_map_f16<16, 256, 1, 2>(a, cls[c].n_slots, p.iter0 + it, stream); break;
            case 23: e = launch_map_f16<16, 256, 1, 3>(a, cls[c].n_slots, p.iter0 + it, stream); break;
            default: e = launch_map_f16<16, 256, 1, 4>(a, cls[c].n_slots, p.iter0 + it, stream); break;
          }
        }""")
rep("  const bool     use_tmaps = opt_map_cfg >= 10;","  const bool     use_tmaps = (opt_map_cfg & 0xff) >= 10;")
open(p,'w').write(s)
