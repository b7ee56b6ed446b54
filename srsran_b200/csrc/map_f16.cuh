// k_map_f16 -- the hot kernel: windowed max-log-MAP of the int16 decoders (tdec_win{sse16,avx16},
// reference include/srslte/phy/fec/turbodecoder_win.h:551-868 + the half-iteration glue of turbodecoder_iter.h:104-128)
// in the Fast16 arithmetic (native packed wrapping ops under the range monitor, arith.cuh / map_core.cuh).
//
// What differs from the generic k_map_win (kernels.cuh), which stays the exact-replay / int8 path:
//   * staging by ONE TMA tensor copy per warp and tile of 8 trellis steps (TensorSrc comment in kernels.cuh explains the
//     box); two stages, one tile ahead.  No LDGSTS, no per-thread address arithmetic for loads.
//   * tiles are aligned to row 0 of the sub-block, so the step parity (normalisation cadence of the reference) is a
//     compile-time property of the unrolled bodies; only a partial TOP tile (W % 8 != 0) runs guarded code.
//   * the beta values of a tile are recomputed from its checkpoint once: four of them stay in registers, three take a
//     round trip through shared memory, the eighth IS the checkpoint -> 32 instead of 64 registers for the segment,
//     which is what lets 16 warps share an SM.
//   * the a-posteriori LLR uses the factored form llr_factored() (20 instead of 28 packed operations per step).
//   * the first steps of lane 0 (known start state [0, -INF x 7]) are covered by their own range monitor instead of
//     being run with the emulated saturating arithmetic.
//   * compile-time specialisation on the constituent decoder: MODE 0 = DEC1 without a-priori input (first
//     half-iteration), 1 = DEC1, 2 = DEC2.
#pragma once
#include "map_core.cuh"

namespace b200 {

// A-posteriori LLR of one forward step against beta_{k+1} in factored form (turbodecoder_win.h:769-813 computes
// max_i(beta_i + alpha_s(i) + gamma_i) per hypothesis with 16 + 14 operations; grouping the branches by their gamma
// in {0, y} / {x, x+y} moves gamma out of the inner maximum).  Equal to fwd_step_llr() whenever no intermediate
// leaves int16, which is what the Fast16 range monitor certifies: the new intermediates beta_i + alpha_s (without
// gamma) are bounded by the same Sp_beta + Sp_alpha + 3g as the original operands.
template <class P, class Mon>
B200_HD u32 llr_factored(const u32 (&o)[8], const u32 (&b)[8], u32 x, u32 y, u32 xy, Mon& mon)
{
  // hypothesis 0: gamma = 0 on (b0,o0) (b3,o7) (b4,o1) (b7,o6); gamma = y on (b1,o3) (b2,o4) (b5,o2) (b6,o5)
  u32 tA = P::add(b[0], o[0]);
  tA     = P::addmax(b[3], o[7], tA);
  tA     = P::addmax(b[4], o[1], tA);
  tA     = P::addmax(b[7], o[6], tA);
  u32 tB = P::add(b[1], o[3]);
  tB     = P::addmax(b[2], o[4], tB);
  tB     = P::addmax(b[5], o[2], tB);
  tB     = P::addmax(b[6], o[5], tB);
  const u32 m0 = P::addmax(tB, y, tA);
  // hypothesis 1: gamma = x on (b1,o2) (b2,o5) (b5,o3) (b6,o4); gamma = x + y on (b0,o1) (b3,o6) (b4,o0) (b7,o7)
  u32 tC = P::add(b[1], o[2]);
  tC     = P::addmax(b[2], o[5], tC);
  tC     = P::addmax(b[5], o[3], tC);
  tC     = P::addmax(b[6], o[4], tC);
  u32 tD = P::add(b[0], o[1]);
  tD     = P::addmax(b[3], o[6], tD);
  tD     = P::addmax(b[4], o[0], tD);
  tD     = P::addmax(b[7], o[7], tD);
  const u32 m1 = P::addmax(tD, xy, P::add(tC, x));
  const u32 d  = P::sub(m1, m0);
  if (P::kMonitor)
    mon.track_sub(m1, m0, d);
  return P::out(d);
}

#if defined(__CUDACC__)

// L2 eviction-priority policies: the LLR planes are streamed (each row is read once per pass, far apart), the beta
// checkpoints are written and read back within one launch -> keep those in L2, let the stream pass through
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t lds_u16(unsigned addr_s)
{
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];\n" : "=h"(v) : "r"(addr_s));
  return v;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal()
{
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_tile4_hint(unsigned dst_s, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar_s, uint64_t pol)
{
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;\n" ::"r"(dst_s),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_s), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s, uint64_t pol)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(dst_s), "l"(src),
               "r"(bytes), "r"(bar_s), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void stg128_hint(void* p, uint4 v, uint64_t pol)
{
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ void stg16_hint(void* p, uint16_t v, uint64_t pol)
{
  asm volatile("st.global.L2::cache_hint.b16 [%0], %1, %2;\n" ::"l"(p), "h"(v), "l"(pol) : "memory");
}

template <int T, int STAGES, int PLANES>
struct F16Lay {
  static constexpr int kRows       = 8;                                  // trellis steps per tile
  static constexpr int kPlaneWords = kRows * 32;                         // one plane of a tile: 8 box rows of 32 words
  static constexpr int kLutOff     = PLANES * kPlaneWords;               // (2 input planes, 3 for DEC1 with a-priori) | QPP rows [8][T]
  static constexpr int kCkOff      = kLutOff + kRows * T;                // checkpoint [2 halves][32 lanes][4 words]
  static constexpr int kStageWords = (kCkOff + 256 + 31) / 32 * 32;      // stages stay 128-byte aligned
  static constexpr int kStages     = STAGES;
  static constexpr int kYOff       = kStages * kStageWords;              // beta spill [3][2 halves][32 lanes][4 words]
  static constexpr int kBarOff     = kYOff + 3 * 256;
  static constexpr int kWarpWords  = (kBarOff + 2 * kStages + 31) / 32 * 32;
};

// MODE: 0 = DEC1 without a-priori input, 1 = DEC1 with a-priori input, 2 = DEC2
template <class P, int N, int MODE, int NT, int MINB, int STAGES>
__global__ void __launch_bounds__(NT, MINB) k_map_f16(const MapArgs a)
{
  constexpr int  T = N / 2, G = 32 / T;
  constexpr int  kNP = P::kNormPeriod; // normalisation cadence of the reference: 2 (int16, by state 0) or 1 (int8, by the max)
  constexpr bool kDec2 = MODE == 2, kApr = MODE == 1;
  using Lay = F16Lay<T, STAGES, kApr ? 3 : 2>;
  constexpr int kStages = STAGES;
  extern __shared__ __align__(128) u32 smem_f[];

  if (a.iter > 0 && a.counters[4 + a.iter - 1] == 0)
    return; // nothing left to decode in this batch
  const int lane  = threadIdx.x & 31;
  const int wib   = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * (NT / 32) + wib;
  const int slot  = gwarp * G + lane / T;
  const int j     = lane % T;
  const int cb    = slot < a.n_slots ? a.work[slot] : -1;
  bool      live  = cb >= 0;
  uint32_t  d_W = 0, d_K = 0, d_ps = 0, d_qpp = 0, d_sat = 0, n_iter0 = 0;
  uint64_t  d_ws = 0;
  if (live) {
    const CbDev*   dp = a.cbs + cb;
    const CbState* sp = a.state + cb;
    d_W     = dp->W;
    d_K     = dp->K;
    d_ps    = dp->ps;
    d_qpp   = dp->qpp_off;
    d_sat   = dp->sat_end;
    d_ws    = dp->ws_off;
    n_iter0 = sp->n_iter;
    if (sp->done || n_iter0 >= dp->max_iter)
      live = false;
  }
  // a warp runs as long as one of its code blocks is live; the lanes of the others follow as ghosts (they compute on
  // whatever the box delivered and never store), so the warp stays converged for the warp-wide staging protocol
  const unsigned live_mask = __ballot_sync(0xffffffffu, live);
  if (live_mask == 0)
    return;
  const int leader = __ffs(live_mask) - 1;
  const int W      = __shfl_sync(0xffffffffu, (int)d_W, leader);
  const int K      = __shfl_sync(0xffffffffu, (int)d_K, leader);
  const int niter  = __shfl_sync(0xffffffffu, (int)n_iter0, leader);
  const int qoff   = __shfl_sync(0xffffffffu, (int)d_qpp, leader);
  {
    const int want = (niter & 1) ? 2 : (niter > 0 ? 1 : 0);
    if (want != MODE || (live && ((int)d_W != W || (int)n_iter0 != niter)))
      __trap(); // host planning launches the variant of this half-iteration; a warp never mixes sizes or parities
  }
  const unsigned gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane / T * T));

  int16_t*        ws = a.ws + d_ws;
  const size_t    ps = d_ps;
  const int16_t*  tl = a.tails + (size_t)(live ? cb : 0) * 12;
  const uint16_t* q  = a.qpp + qoff;

  // ---- staging
  u32*           sm   = smem_f + wib * Lay::kWarpWords;
  const unsigned sm_s = (unsigned)__cvta_generic_to_shared(sm);
  const u32*     my   = sm + lane; // a box row holds one word per lane: (block in warp) * T + j == lane
  const CUtensorMap* tmap   = a.tmaps + 2 * a.winfo[2 * gwarp] + (kApr ? 0 : 1);
  const int          blk0   = a.winfo[2 * gwarp + 1];
  constexpr int      plane0 = kDec2 ? kPlApp2 : kPlSyst;
  constexpr unsigned kBoxBytes = (kApr ? 3u : 2u) * Lay::kRows * 128u;
  const u32*         lut  = (const u32*)(kDec2 ? q : q + K); // rev[] pairs for DEC1, fwd[] pairs for DEC2
  const size_t       ck_stride = (size_t)gridDim.x * (NT / 32) * 256; // words per checkpoint slot
  u32*               ck_warp   = a.ck_scratch + (size_t)gwarp * 256;  // [2 halves][32 lanes][4 words] of slot 0

  // tile sequence: beta warm-up (tiles 4..0), beta main (top..0), alpha warm-up (a0..top), alpha main (0..top)
  const int nT  = (W + 7) >> 3;
  const int a0  = (W - kWinOverlap) >> 3;
  const int nAW = nT - a0;
  const int s1 = 5, s2 = s1 + nT, s3 = s2 + nAW, n_seq = s3 + nT;
  const bool     write_post = (a.mode & kMapSkipPost) == 0; // the a-posteriori plane is only read by a hard decision
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
      tma_tile4_hint(dst, tmap, 0, blk0, 8 * t, plane0, bar, pol_first);
      if (aux) {
        bulk_g2s(dst + 4u * (unsigned)Lay::kLutOff, lut + (size_t)8 * t * T, lut_bytes, bar);
        bulk_g2s_hint(dst + 4u * (unsigned)Lay::kCkOff, ck_warp + (size_t)(t + 1) * ck_stride, 1024u, bar, pol_first);
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
  // x (systematic + a-priori) and y (parity) of box row i
  auto row = [&](const u32* tb, int i, u32& x, u32& y) {
    const u32 vin = tb[i * 32];
    y             = tb[Lay::kPlaneWords + i * 32];
    x             = kApr ? P::add(tb[2 * Lay::kPlaneWords + i * 32], vin) : vin;
  };

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kStages; i++)
      mbar_init(bar_of(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kStages - 1; i++)
    issue();

  RangeMon mon_b, mon_a, mon_h;
  mon_b.reset();
  mon_a.reset();
  mon_h.reset();
  u32 st[8];

  // =============================================================== backward
  // ---- warm-up: steps 39..0 of the lane's own sub-block from the all-"unknown" state (win.h:622-630)
#pragma unroll
  for (int s = 0; s < 8; s++)
    st[s] = splat16(-P::kInf);
  for (int t = 4; t >= 0; t--) {
    const u32* tb = acquire();
#pragma unroll
    for (int i = 7; i >= 0; i--) {
      u32 x, y;
      row(tb, i, x, y);
      bwd_step<P>(st, x, y, P::add(x, y));
      if (P::kMonitor && (i & 1) == 0 && (t < 4 || i < 6))
        mon_b.track(st); // k < 38: the first two steps start from eight equal values (spread 0, covered by g)
      if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
        P::normalize_now(st);
    }
  }
  // hand the estimate to the lane below; tail trellis for the last lane (win.h:580-612, 500-548)
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const u32 nx = __shfl_down_sync(gmask, st[s], 1, T);
    st[s]        = shift_down_lanes(st[s], nx);
  }
  if (j == T - 1) {
    int32_t tt[8];
    tail_trellis<P>(kDec2 ? tl + 6 : tl, kDec2 ? tl + 9 : tl + 3, tt);
#pragma unroll
    for (int s = 0; s < 8; s++)
      st[s] = (st[s] & 0xffffu) | ((u32)(uint16_t)tt[s] << 16);
  }
  // ---- main pass with one checkpoint per tile: slot t = beta[8t] before normalisation, slot nT = beta[W]
  auto ck_store = [&](int sl, const u32 (&v)[8]) {
    if (live) {
      uint4* g = reinterpret_cast<uint4*>(ck_warp + (size_t)sl * ck_stride) + lane;
      stg128_hint(g, make_uint4(v[0], v[1], v[2], v[3]), pol_last);
      stg128_hint(g + 32, make_uint4(v[4], v[5], v[6], v[7]), pol_last);
    }
  };
  if (P::kMonitor)
    mon_b.track(st);
  ck_store(nT, st);
  {
    int t = nT - 1;
    if (W & 7) { // partial top tile: guarded, rolled
      const u32* tb = acquire();
#pragma unroll 1
      for (int i = (W & 7) - 1; i >= 0; i--) {
        u32 x, y;
        row(tb, i, x, y);
        bwd_step<P>(st, x, y, P::add(x, y));
        if (i == 0)
          ck_store(t, st);
        if (P::kMonitor && (i & 1) == 0)
          mon_b.track(st);
        if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
          P::normalize_now(st);
      }
      t--;
    }
    for (; t >= 0; t--) {
      const u32* tb = acquire();
#pragma unroll
      for (int i = 7; i >= 0; i--) {
        u32 x, y;
        row(tb, i, x, y);
        bwd_step<P>(st, x, y, P::add(x, y));
        if (i == 0)
          ck_store(t, st);
        if (P::kMonitor && (i & 1) == 0)
          mon_b.track(st);
        if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
          P::normalize_now(st);
      }
    }
  }
  // the checkpoints were written through the generic proxy and come back through the async proxy
  asm volatile("fence.proxy.async.global;\n" ::: "memory");

  // bound on every |branch metric| of this call: max|a-priori| + max|systematic| + max|parity|
  int*      gm = a.gmax + (size_t)(live ? cb : 0) * 4;
  const int g  = !P::kMonitor ? 0 : kDec2 ? gm[3] + gm[2] : (kApr ? gm[3] : 0) + gm[0] + gm[1];
  if (P::kMonitor) {
    const bool bad = !fast16_beta_ok(mon_b.spread_lo(), g) || !fast16_beta_ok(mon_b.spread_hi(), g);
    if (__any_sync(gmask, bad && live)) { // the whole code block is replayed with the exact policy (mode 2 launch)
      if (j == 0 && live)
        a.state[cb].redo = 1;
      live = false;
    }
  }

  // =============================================================== forward
  // ---- warm-up: steps W-40..W-1 of the lane's own sub-block (win.h:747-756); normalisation follows the loop counter
#pragma unroll
  for (int s = 0; s < 8; s++)
    st[s] = splat16(-P::kInf);
  {
    int kk = 0; // loop counter of the pass
    for (int t = a0; t < nT; t++) {
      const u32* tb = acquire();
      const int  i0 = t == a0 ? (W - kWinOverlap) - 8 * a0 : 0;
      const int  i1 = (8 * t + 8) <= W ? 8 : W - 8 * t;
#pragma unroll 1
      for (int i = i0; i < i1; i++, kk++) {
        u32 x, y;
        row(tb, i, x, y);
        fwd_step<P>(st, x, y, P::add(x, y));
        if (P::kMonitor && (kk & 1) == 0 && kk > 2)
          mon_a.track(st);
        if ((kNP == 1 || (kk & 1) == 0) && kk != 0)
          P::normalize_now(st);
      }
    }
  }
  // hand the estimate to the lane above; lane 0 starts from the known state [0, -INF x 7]
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const u32 pv = __shfl_up_sync(gmask, st[s], 1, T);
    st[s]        = shift_up_lanes(pv, st[s]);
  }
  if (j == 0) {
    st[0] = st[0] & 0xffff0000u;
#pragma unroll
    for (int s = 1; s < 8; s++)
      st[s] = (st[s] & 0xffff0000u) | (u32)(uint16_t)(-P::kInf);
  }
  // The first four steps (and the start state) are tracked by mon_h: the -INF entries of lane 0 widen the spread
  // only there (after three steps every state is reachable from state 0), and both monitors are checked below.
  if (P::kMonitor)
    mon_h.track(st);

  // ---- output pass
  u32* const     post  = (u32*)(ws + kPlPost * ps);
  int16_t* const post16 = ws + kPlPost * ps;
  int16_t* const ext   = kDec2 ? ws + kPlApr * ps : ws + kPlApp2 * ps; // extrinsic output, scattered through the QPP table
  u32            ehi = 0, elo = 0, e_even = 0;
  u32            al[8];
#pragma unroll
  for (int s = 0; s < 8; s++)
    al[s] = st[s];

  // one forward step with output: LLR against b = beta_{p+1}, state update, glue epilogue (iter.h:107-127)
  auto out_step = [&](const u32* tb, int t, int i, const u32 (&b)[8], RangeMon& mon, bool norm) {
    u32 x, y;
    row(tb, i, x, y);
    const u32 xy = P::add(x, y);
    u32       llr;
    if (P::kMonitor) { // wrapping arithmetic under the range monitor: factored form
      llr = llr_factored<P>(al, b, x, y, xy, mon);
      fwd_step<P>(al, x, y, xy);
    } else { // saturating arithmetic: the operation order of the reference is part of the result
      llr = fwd_step_llr<P>(al, b, x, y, xy, mon);
    }
    if (P::kMonitor && (i & 1) == 0)
      mon.track(al);
    if ((kNP == 1 || (i & 1) == 0) && norm)
      P::normalize_now(al);
    // the two scatter targets of this thread (QPP table word = two uint16 indices), read as halves: two LDS.U16 on the
    // load/store pipe instead of LDS + LOP3 + SHF on the max-type integer pipe, which is the busiest one
    const uint16_t* r16 = reinterpret_cast<const uint16_t*>(tb + (Lay::kLutOff + i * T + j - lane));
    const uint32_t  t0 = r16[0], t1 = r16[1];
    // running max / min of the extrinsic values, two steps per VIMNMX3
    auto track_e = [&](u32 e) {
      if (P::kMonitor) {
        if ((i & 1) == 0) {
          e_even = e;
        } else {
          ehi = p_max3(ehi, e_even, e);
          elo = p_min3(elo, e_even, e);
        }
      }
    };
    if (!kDec2) {
      // a-posteriori -> post (linear); extrinsic - a-priori -> app2[rev[.]]
      const uint32_t w2 = 2u * (uint32_t)((8 * t + i) * T + j);
      const u32      e  = kApr ? P::glue_sub(llr, tb[2 * Lay::kPlaneWords + i * 32], w2 < d_sat, w2 + 1 < d_sat) : llr;
      track_e(e);
      if (live && write_post)
        post[(8 * t + i) * T + j] = llr;
      if (live) {
        ext[t0] = (int16_t)lo16(e);
        ext[t1] = (int16_t)hi16(e);
      }
    } else {
      // a-posteriori -> post[fwd[.]]; a-posteriori - own input -> a-priori[fwd[.]]
      const u32 e = P::glue_sub(llr, x, t0 < d_sat, t1 < d_sat);
      track_e(e);
      if (live) {
        ext[t0]    = (int16_t)lo16(e);
        ext[t1]    = (int16_t)hi16(e);
      }
      if (live && write_post) {
        post16[t0] = (int16_t)lo16(llr);
        post16[t1] = (int16_t)hi16(llr);
      }
    }
  };
  auto ck_load = [&](const u32* tb, u32 (&v)[8]) { // checkpoint of the tile: [half][lane][4 words]
    const uint4* c = reinterpret_cast<const uint4*>(tb - lane + Lay::kCkOff) + lane;
    const uint4  lo = c[0], hi = c[32];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  };
  uint4* const ysp = reinterpret_cast<uint4*>(sm + Lay::kYOff) + lane; // beta spill: entry y at ysp[64 y], ysp[64 y + 32]

  const int n_full = W >> 3;
  for (int t = 0; t < n_full; t++) {
    const u32* tb = acquire();
    u32        bs[4][8];
    // ---- recompute beta_{8t+7} .. beta_{8t+1} from the checkpoint beta_{8t+8}
    ck_load(tb, st);
    if (8 * (t + 1) < W)
      P::normalize_now(st); // the recursion continued from the normalised value; beta[W] itself was never normalised
#pragma unroll
    for (int kk = 7; kk >= 1; kk--) {
      u32 x, y;
      row(tb, kk, x, y);
      bwd_step<P>(st, x, y, P::add(x, y));
      if (kk >= 5) {
        ysp[64 * (kk - 5)]      = make_uint4(st[0], st[1], st[2], st[3]);
        ysp[64 * (kk - 5) + 32] = make_uint4(st[4], st[5], st[6], st[7]);
      } else {
#pragma unroll
        for (int s = 0; s < 8; s++)
          bs[kk - 1][s] = st[s];
      }
      if (kNP == 1 || (kk & 1) == 0)
        P::normalize_now(st);
    }
    // ---- steps 8t .. 8t+3 against beta_{8t+1} .. beta_{8t+4}
#pragma unroll
    for (int i = 0; i < 4; i++)
      out_step(tb, t, i, bs[i], mon_a, i != 0 || t != 0);
    if (P::kMonitor && t == 0) { // what was tracked so far belongs to the head monitor
      mon_h.hi = p_max(mon_h.hi, mon_a.hi);
      mon_h.lo = p_min(mon_h.lo, mon_a.lo);
      mon_a.hi = 0;
      mon_a.lo = 0;
    }
    // ---- steps 8t+4 .. 8t+7 against beta_{8t+5} .. beta_{8t+7} (spilled by this lane) and the checkpoint beta_{8t+8}
#pragma unroll
    for (int y = 0; y < 3; y++) {
      const uint4 lo = ysp[64 * y], hi = ysp[64 * y + 32];
      bs[y][0] = lo.x; bs[y][1] = lo.y; bs[y][2] = lo.z; bs[y][3] = lo.w;
      bs[y][4] = hi.x; bs[y][5] = hi.y; bs[y][6] = hi.z; bs[y][7] = hi.w;
    }
    ck_load(tb, bs[3]);
#pragma unroll
    for (int i = 4; i < 8; i++)
      out_step(tb, t, i, bs[i - 4], mon_a, true);
  }
  if (W & 7) {
    // partial top tile, guarded: beta_{p+1} of each step is recomputed from the checkpoint beta[W] (at most 6 steps)
    const int  t  = n_full;
    const int  nv = W & 7;
    const u32* tb = acquire();
#pragma unroll 1
    for (int i = 0; i < nv; i++) {
      u32 b[8];
      ck_load(tb, b);
#pragma unroll 1
      for (int kk = nv - 1; kk > i; kk--) { // -> beta_{8t+kk}, normalised on the way except the one that is used
        u32 x, y;
        row(tb, kk, x, y);
        bwd_step<P>(b, x, y, P::add(x, y));
        if (kk > i + 1 && (kNP == 1 || (kk & 1) == 0))
          P::normalize_now(b);
      }
      out_step(tb, t, i, b, mon_a, true); // (t >= 5 here: a lane has at least 40 steps)
    }
  }

  if (P::kMonitor) {
    // max |extrinsic| handed to the next half-iteration (its a-priori / systematic input)
    ehi    = p_max(ehi, e_even); // (an odd number of steps leaves the last one pending; counting one twice is harmless)
    elo    = p_min(elo, e_even);
    int ge = max(max(lo16(ehi), hi16(ehi)), max(-lo16(elo), -hi16(elo)));
#pragma unroll
    for (int o = T / 2; o >= 1; o >>= 1)
      ge = max(ge, __shfl_xor_sync(gmask, ge, o, T));
    const bool bad = !fast16_alpha_ok(mon_a.spread_lo(), mon_b.spread_lo(), g) || !fast16_alpha_ok(mon_a.spread_hi(), mon_b.spread_hi(), g) ||
                     !fast16_alpha_ok(mon_h.spread_lo(), mon_b.spread_lo(), g) || !fast16_alpha_ok(mon_h.spread_hi(), mon_b.spread_hi(), g) ||
                     ((mon_a.ovf | mon_h.ovf) & 0x80008000u) != 0;
    if (__any_sync(gmask, bad && live)) {
      if (j == 0 && live)
        a.state[cb].redo = 1;
      return;
    }
    if (j == 0 && live)
      gm[3] = ge;
  }
}

#endif // __CUDACC__

} // namespace b200
