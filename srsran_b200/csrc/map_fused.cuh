// k_map_fused -- the persistent per-code-block pipeline: ALL half-iterations of a group of code blocks in one launch.
//
// Replaces the reference's per-code-block loop (lib/src/phy/phch/sch.c:420-450: srslte_tdec_iteration -> hard decision ->
// srslte_crc_checksum_byte -> stop on CRC == 0 or at the iteration limit; srslte_tdec_run_all, turbodecoder.c:537-578, is the
// same loop without the CRC) around the windowed max-log-MAP of include/srslte/phy/fec/turbodecoder_win.h:551-868 and the
// half-iteration glue of turbodecoder_iter.h:104-128.
//
// One WARP owns a group of G = 64/N code blocks of equal K (one thread = two adjacent sub-block lanes of one block, int16x2)
// from its first half-iteration to its last:
//
//     fetch group (atomic counter; groups are sorted by K, longest first)
//     repeat   DEC1 | DEC2 half-iteration: beta warm-up, beta pass (one checkpoint per 8-step tile), alpha warm-up, alpha
//                pass with the a-posteriori LLRs, the glue (extrinsic scatter through the QPP rows) and the HARD DECISIONS
//                packed as bits into a K-bit string per block in shared memory
//              CRC24 of the decided bytes by the block's own T lanes, early stop per block, decided bytes written once
//     until    every block of the group passed its CRC or reached its iteration limit
//
// so a batch costs ONE launch per decoder class instead of 2 x max_iterations launches, the a-posteriori plane and the
// decision kernel's read-back of it are gone, and there is no wave quantisation: the grid is one CTA of 12 warps per SM and a
// warp that finishes a group takes the next one.  Staging (TMA tensor tiles, three stages, one mbarrier per warp and
// stage), the checkpoint/recompute schedule, the factored LLR and the Fast16 range monitor are those of DESIGN.md 5.1 / 5.2.
// A block the monitor flags is PARKED (its inputs of that half-iteration are intact): the same kernel compiled with the
// exact saturating policy (Sat16) takes the parked groups in a second launch and finishes them.
#pragma once
#include <type_traits>

#include "map_f16.cuh"

namespace b200 {

#if defined(__CUDACC__)

struct FusedArgs {
  const int*         work;      // code block per slot, -1 = padding; the G slots of a group share K
  int                n_groups;
  const CbDev*       cbs;
  CbState*           state;
  int16_t*           ws;
  const int16_t*     tails;
  const uint16_t*    qpp;       // per (K, lanes): fwd[K] | rev[K] | per 8-step tile: fwd rows, natural bit index of their targets
  int*               gmax;
  u32*               ck_scratch; // resident warps x ck_words: beta checkpoints of the half-iteration in flight
  uint32_t           ck_words;
  uint32_t*          counters;  // [0] half-iterations run with the exact policy after parking, [1] half-iterations run,
                                // [2] groups parked by the Fast16 launch, [ctr_fetch] group fetch counter of this launch
  int                ctr_fetch;
  int*               parked;    // group ids holding parked blocks (written by mode 1, consumed by mode 2)
  // time slicing (batches with CRC early stop, modes 0 / 1): a warp runs a group for slice_first (first visit) / slice_next
  // half-iterations and, when other groups are waiting, hands the unfinished group back through queue[] instead of keeping
  // it: every group advances at the same pace, so the blocks that need all their iterations do not start them late, and a
  // warp is never idle while work is waiting.  slice_first == 0: a group stays with its warp until it is finished
  int*               queue;     // [n_groups * max_iter] handed-back group ids (-1 = not yet published) | [n_groups] "already on the parked list" flags
  int                queue_cap;
  int                slice_first, slice_next;
  int                ctr_tail;  // counters[ctr_tail]: number of groups handed back so far
  int                ctr_avail; // counters[ctr_avail] (signed) + n_groups: entries not yet claimed
  int                mode;      // 0: every active block; 1: Fast16 attempt, flagged blocks are parked; 2: parked blocks only
  const int*         winfo;     // per group {index of its K-group's pair of tensor maps, block coordinate of its first slot}
  const CUtensorMap* tmaps;     // per K-group: [2g] box of 3 planes, [2g+1] box of 2 planes
  uint8_t*           cb_out;
  const uint32_t*    crc_tab;   // [2][256]: CRC24A, CRC24B byte tables (global memory copy, read through L1)
  int                warp_words; // shared memory words per warp (FusedLay<T>::kFixedWords + G * bits words per block)
  int                bits_words; // words of the K-bit decision string per block (max over the class)
  uint32_t           dump_off;   // word offset, inside a warp's ck_scratch share, of the area ghost lanes scatter into
  int                ck_policy;  // beta checkpoints: 0 = stored evict_last / read back evict_first; 1 = no eviction hint (the stream's access policy window decides)
};

template <int T>
struct FusedLay {
  static constexpr int kRows       = 8;
  static constexpr int kPlaneWords = kRows * 32;
  static constexpr int kLutWords   = kRows * T;
  // stage: three plane slices | QPP rows | natural-index rows (decoder 2 only).  Decoder 1 with a-priori input stages 3
  // planes, every other half-iteration 2 (the third slice is zeroed).  The beta checkpoint of an alpha tile is requested one tile later than its planes and lands in a
  // ring of two (it only has to arrive before its tile is consumed; the third copy of it was 1 KB of shared memory per warp)
  static constexpr int kStageWords = (3 * kPlaneWords + 2 * kLutWords + 31) / 32 * 32;
  static constexpr int kStages     = 3;
  static constexpr int kCkRing     = kStages * kStageWords;  // [2][2 halves][32 lanes][4 words]
  static constexpr int kYOff       = kCkRing + 2 * 256;      // beta spill [3][2 halves][32 lanes][4 words]
  static constexpr int kBarOff     = kYOff + 3 * 256;        // kStages mbarriers
  static constexpr int kBitsOff    = kBarOff + 8;            // decision bits: G blocks x bits_words
  static constexpr int kFixedWords = kBitsOff;
};

// CRC24 of the decided bytes of one code block by the block's own T lanes (T = 4, 8, 16 consecutive lanes of the warp).
// bits: the block's K decisions, natural bit n at word n / 32, bit n % 32; byte b of the message holds bits 8b..8b+7, first
// bit = MSB (turbodecoder_win.h:925-993).  Each lane runs the reference's byte recurrence (crc.h:56-63) over one chunk of
// ceil(nbytes / T) bytes (the message is virtually left-padded with zero bytes, which do not change a zero-initialised CRC);
// chunks are combined with crc(A || B) = crc(A) x^(8|B|) + crc(B) mod g, xq[l] = x^(8 * chunk * 2^l) mod g from the host.
template <int T>
__device__ __forceinline__ uint32_t group_crc24(const u32* bits, uint32_t nbytes, const uint32_t* __restrict__ tab, uint32_t poly, const uint32_t* xq,
                                                int j, unsigned gmask)
{
  const uint32_t cbk = (nbytes + T - 1) / T, pad = T * cbk - nbytes;
  uint32_t       crc = 0;
  for (uint32_t q = 0; q < cbk; q++) {
    const uint32_t pos = (uint32_t)j * cbk + q;
    if (pos >= pad) {
      const uint32_t b    = pos - pad;
      const uint32_t byte = __brev((bits[b >> 2] >> (8u * (b & 3u))) & 0xffu) >> 24;
      crc                 = ((crc << 8) ^ __ldg(tab + (((crc >> 16) & 0xffu) ^ byte))) & 0xffffffu;
    }
  }
#pragma unroll
  for (int lv = 0; (1 << lv) < T; lv++) {
    const int      l     = 1 << lv;
    const uint32_t other = __shfl_down_sync(gmask, crc, l, T);
    if ((j & (2 * l - 1)) == 0)
      crc = crc24_mulmod(crc, xq[lv], poly) ^ other;
  }
  return __shfl_sync(gmask, crc, 0, T);
}

// per-warp state that lives across half-iterations and groups
template <int T>
struct FusedWarp {
  u32*     sm;        // this warp's shared memory
  unsigned sm_s;
  int      lane, j;
  unsigned gmask;
  int      wr_stage, rd_stage;
  unsigned rd_phase;
  uint64_t pol_first, pol_last, pol_ck_st, pol_ck_ld;
  u32*     ck_warp;   // checkpoint scratch of this resident warp: slot s at ck_warp + 256 s
  u32*     bits;      // decision bits of this thread's block
};

// ONE half-iteration of the group -- ONE code body for both constituent decoders.  The instruction cache behind an SM's
// twelve warps is 32 KB (L1.5) and the warps sit in different phases of different half-iterations: a build with one
// specialised body per (decoder, a-priori, decisions) combination ran the int8 kernel at a 59 % instruction-cache hit rate
// and 5 stall cycles per instruction waiting for instructions.  So everything that differs between the decoders is data:
//   dec2     decoder 2: inputs app2 | par1, extrinsic = a-posteriori - own input scattered through fwd[] into the a-priori
//            plane, decisions at natural positions nat[.]; else decoder 1: inputs syst | par0 | a-priori, extrinsic =
//            a-posteriori - a-priori scattered through rev[] into app2, decisions in this thread's own lanes
//   apr3     the tiles carry three planes (decoder 1 with a-priori input).  Otherwise they carry two and the a-priori slice
//            of every stage is zeroed once: x = 0 + input and e = llr - 0 are then the reference's values
//   bits     the hard decisions of this half-iteration are needed (a CRC check or the end of the run follows)
// and the unrolled bodies are shared: the two beta passes (warm-up, main) run the same 8-step tile, the alpha tile is two
// trips through one 4-step body.
// Returns true for the threads of blocks whose Fast16 range monitor cannot rule a saturation out.
// DEC: -1 = the decoder is a runtime property (one body: the int8 and exact-int16 kernels, whose code is large); 0 / 1 = this
// instance is decoder 1 / decoder 2 (the Fast16 kernel affords two bodies: 2 x 25 KB of hot loops still hit the instruction
// cache at 97 %, and decoder 2 then skips the add of its zero a-priori slice)
template <class P, int N, int DEC>
__device__ __forceinline__ bool fused_half(FusedWarp<N / 2>& w, const FusedArgs& a, const CUtensorMap* tmap, int blk0, int W, int K, const uint16_t* q,
                                           int16_t* ws, size_t ps, const int16_t* tl, int g_in, uint32_t d_sat, bool live, int& ge_out, bool dec2_rt,
                                           bool apr3, bool bits)
{
  const bool dec2 = DEC < 0 ? dec2_rt : DEC == 1;
  constexpr int  T = N / 2;
  constexpr int  kNP = P::kNormPeriod;
  using Lay = FusedLay<T>;
  constexpr int  kStages = Lay::kStages;
  constexpr int  kLutOff = 3 * Lay::kPlaneWords;
  const int      plane0    = dec2 ? kPlApp2 : kPlSyst;
  const unsigned kBoxBytes = (apr3 ? 3u : 2u) * Lay::kRows * 128u;
  const int      lane = w.lane, j = w.j;
  const unsigned gmask = w.gmask;
  const u32*     my   = w.sm + lane; // a box row holds one word per lane
  // QPP rows of a tile: rev[] pairs for decoder 1; for decoder 2 one record per tile, fwd[] pairs then the natural bit index
  // of their targets
  const u32*     lut        = (const u32*)(dec2 ? q + 2 * (size_t)K : q + K);
  const unsigned lut_stride = dec2 ? 2u * Lay::kLutWords : (unsigned)Lay::kLutWords;
  const unsigned lut_bytes  = (dec2 && bits) ? 2u * Lay::kLutWords * 4u : Lay::kLutWords * 4u;

  if (!apr3) {
#pragma unroll
    for (int sidx = 0; sidx < kStages; sidx++)
#pragma unroll
      for (int i = 0; i < Lay::kRows; i++)
        w.sm[sidx * Lay::kStageWords + 2 * Lay::kPlaneWords + i * 32 + lane] = 0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }

  // tile sequence: beta warm-up (tiles 4..0), beta main (top..0), alpha warm-up (a0..top), alpha main (0..top)
  const int nT  = (W + 7) >> 3;
  const int a0  = (W - kWinOverlap) >> 3;
  const int nAW = nT - a0;
  const int s1 = 5, s2 = s1 + nT, s3 = s2 + nAW, n_seq = s3 + nT;
  auto bar_of = [&](int stage) -> unsigned { return w.sm_s + 4u * (unsigned)(Lay::kBarOff + 2 * stage); };
  int  wr_idx = 0;
  // kAux = false_type: the caller knows that the tile requested now is not an alpha-main tile (both beta passes: the
  // request runs two tiles ahead of the consumer, so they only ever request beta and alpha warm-up tiles) -- no table rows,
  // no checkpoint, a third of the code
  auto issue = [&](auto kAux) {
    constexpr bool aux_possible = decltype(kAux)::value;
    __syncwarp(); // every lane is done with the stage (and the checkpoint slot) about to be refilled
    if (aux_possible && lane == 0 && wr_idx > s3 && wr_idx <= n_seq) {
      // checkpoint beta[8(t+1)] of the alpha tile whose planes the PREVIOUS call requested: same mbarrier, ring slot t & 1
      const int t     = wr_idx - 1 - s3;
      const int stage = w.wr_stage == 0 ? kStages - 1 : w.wr_stage - 1;
      bulk_g2s_hint(w.sm_s + 4u * (unsigned)(Lay::kCkRing + (t & 1) * 256), w.ck_warp + (size_t)(t + 1) * 256, 1024u, bar_of(stage), w.pol_ck_ld);
    }
    if (wr_idx < n_seq) {
      if (lane == 0) {
        const unsigned bar = bar_of(w.wr_stage);
        const unsigned dst = w.sm_s + 4u * (unsigned)(w.wr_stage * Lay::kStageWords);
        int            t;
        bool           aux = false;
        if (wr_idx < s1)
          t = 4 - wr_idx;
        else if (wr_idx < s2)
          t = nT - 1 - (wr_idx - s1);
        else if (!aux_possible || wr_idx < s3)
          t = a0 + (wr_idx - s2);
        else {
          t   = wr_idx - s3;
          aux = true;
        }
        // (the tables are padded to whole tiles: a partial top tile copies rows nobody reads)
        mbar_expect_tx(bar, aux ? kBoxBytes + lut_bytes + 1024u : kBoxBytes);
        tma_tile4_hint(dst, tmap, 0, blk0, 8 * t, plane0, bar, w.pol_first);
        if (aux_possible && aux)
          bulk_g2s(dst + 4u * (unsigned)kLutOff, lut + (size_t)t * lut_stride, lut_bytes, bar);
      }
      wr_idx++;
      w.wr_stage = w.wr_stage + 1 == kStages ? 0 : w.wr_stage + 1;
    } else if (aux_possible && wr_idx == n_seq) {
      wr_idx++; // (the last checkpoint has just been requested)
    }
  };
  auto acquire = [&](auto kAux) -> const u32* {
    issue(kAux);
    mbar_wait(bar_of(w.rd_stage), (w.rd_phase >> w.rd_stage) & 1u);
    w.rd_phase ^= 1u << w.rd_stage;
    const u32* tb = my + w.rd_stage * Lay::kStageWords;
    w.rd_stage    = w.rd_stage + 1 == kStages ? 0 : w.rd_stage + 1;
    return tb;
  };
  constexpr std::false_type kBeta{};  // requests made while a beta pass consumes
  constexpr std::true_type  kAlpha{}; // requests made while an alpha pass consumes
  // x (input + a-priori) and y (parity) of box row i
  auto row = [&](const u32* tb, int i, u32& x, u32& y) {
    y = tb[Lay::kPlaneWords + i * 32];
    x = DEC == 1 ? tb[i * 32] : P::add(tb[2 * Lay::kPlaneWords + i * 32], tb[i * 32]);
  };

#pragma unroll
  for (int i = 0; i < kStages - 1; i++)
    issue(kBeta);

  RangeMon mon_b, mon_a, mon_h;
  mon_b.reset();
  mon_a.reset();
  mon_h.reset();
  u32 st[8];
  // the recursion steps run with the cheaper rules for normalised input metrics where the policy has them (Sat8F, arith.cuh);
  // a step whose input is NOT normalised is followed by PS::fix (marked "un-normalised input" below)
  using PS = typename StepPolicy<P>::type;

  // =============================================================== backward
  // pass 0: warm-up, steps 39..0 of the lane's own sub-block from the all-"unknown" state (win.h:622-630)
  // pass 1: main pass from the neighbour's estimate / the tail trellis, one checkpoint per tile: slot t = beta[8t] before
  //         normalisation, slot nT = beta[W]  (every lane stores, ghosts too: the scratch belongs to the warp)
  auto ck_store = [&](int sl, const u32 (&v)[8]) {
    uint4* g = reinterpret_cast<uint4*>(w.ck_warp + (size_t)sl * 256) + lane;
    stg128_hint(g, make_uint4(v[0], v[1], v[2], v[3]), w.pol_ck_st);
    stg128_hint(g + 32, make_uint4(v[4], v[5], v[6], v[7]), w.pol_ck_st);
  };
#pragma unroll
  for (int s = 0; s < 8; s++)
    st[s] = splat16(-P::kInf);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    int t = 4;
    if (pass == 1) {
      // hand the estimate to the lane below; tail trellis for the last lane (win.h:580-612, 500-548)
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const u32 nx = __shfl_down_sync(gmask, st[s], 1, T);
        st[s]        = shift_down_lanes(st[s], nx);
      }
      if (j == T - 1) {
        int32_t tt[8];
        tail_trellis<P>(dec2 ? tl + 6 : tl, dec2 ? tl + 9 : tl + 3, tt);
#pragma unroll
        for (int s = 0; s < 8; s++)
          st[s] = (st[s] & 0xffffu) | ((u32)(uint16_t)tt[s] << 16);
      }
      if (P::kMonitor)
        mon_b.track(st);
      ck_store(nT, st);
      t = nT - 1;
      if (W & 7) { // partial top tile: guarded, rolled
        const u32* tb = acquire(kBeta);
#pragma unroll 1
        for (int i = (W & 7) - 1; i >= 0; i--) {
          u32 x, y;
          row(tb, i, x, y);
          bwd_step<PS>(st, x, y, P::add(x, y));
          if (PS::kNeedsFix && i == (W & 7) - 1)
            PS::fix(st); // un-normalised input: the neighbour's estimate / the tail trellis
          if (i == 0)
            ck_store(t, st);
          if (P::kMonitor && (i & 1) == 0)
            mon_b.track(st);
          if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
            P::normalize_now(st);
        }
        t--;
      }
    }
    const bool main_pass = pass == 1;
    bool       raw_in    = main_pass && (W & 7) == 0; // the main pass starts from an un-normalised state
#pragma unroll 1
    for (; t >= 0; t--) {
      const u32* tb = acquire(kBeta);
#pragma unroll
      for (int i = 7; i >= 0; i--) {
        u32 x, y;
        row(tb, i, x, y);
        bwd_step<PS>(st, x, y, P::add(x, y));
        if (PS::kNeedsFix && i == 7 && raw_in) {
          PS::fix(st); // un-normalised input: the neighbour's estimate / the tail trellis
          raw_in = false;
        }
        if (i == 0 && main_pass)
          ck_store(t, st);
        // warm-up, k = 38 (t = 4, i = 6): the first two steps start from eight equal values (spread 0, covered by g)
        if (P::kMonitor && (i & 1) == 0 && (i != 6 || main_pass || t < 4))
          mon_b.track(st);
        if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
          P::normalize_now(st);
      }
    }
  }
  // the checkpoints were written through the generic proxy and come back through the async proxy
  asm volatile("fence.proxy.async.global;\n" ::: "memory");

  // bound on every |branch metric| of this call: max|a-priori| + max|systematic| + max|parity| (g_in, from the caller)
  const int g = g_in;
  bool      flagged = false;
  if (P::kMonitor) {
    const bool bad = !fast16_beta_ok(mon_b.spread_lo(), g) || !fast16_beta_ok(mon_b.spread_hi(), g);
    if (__any_sync(gmask, bad && live)) {
      flagged = live;
      live    = false;
    }
  }

  // =============================================================== forward
  // ---- warm-up: steps W-40..W-1 of the lane's own sub-block (win.h:747-756); normalisation follows the loop counter
#pragma unroll
  for (int s = 0; s < 8; s++)
    st[s] = splat16(-P::kInf);
  {
    int kk = 0; // loop counter of the pass
#pragma unroll 1
    for (int t = a0; t < nT; t++) {
      const u32* tb = acquire(kAlpha);
      const int  i0 = t == a0 ? (W - kWinOverlap) - 8 * a0 : 0;
      const int  i1 = (8 * t + 8) <= W ? 8 : W - 8 * t;
#pragma unroll 1
      for (int i = i0; i < i1; i++, kk++) {
        u32 x, y;
        row(tb, i, x, y);
        fwd_step<PS>(st, x, y, P::add(x, y));
        if (PS::kNeedsFix && kk == 1)
          PS::fix(st); // un-normalised input: step 0 is not followed by a normalisation
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
  // The first four steps (and the start state) are tracked by mon_h
  if (P::kMonitor)
    mon_h.track(st);

  // ---- output pass
  // (ghost lanes -- blocks of the group that are finished or parked -- run the same instruction stream without a branch:
  //  their extrinsic values go to the warp's dump area, their decision bits to their own unused bit string)
  char* const ext = live ? reinterpret_cast<char*>(dec2 ? ws + kPlApr * ps : ws + kPlApp2 * ps) : reinterpret_cast<char*>(w.ck_warp + a.dump_off);
  u32         ehi = 0, elo = 0, e_even = 0;
  u32         al[8];
#pragma unroll
  for (int s = 0; s < 8; s++)
    al[s] = st[s];
  u32         acc = 0; // decoder 1: decisions of up to 16 steps, both lanes, first step = MSB
  int         acc_n = 0;
  char* const bits_c = reinterpret_cast<char*>(w.bits);
  // (the flags go through an empty asm in every tile so that the compiler cannot unswitch the loop into one copy per flag
  //  combination: code size is what this body is about)
  int         flags = (dec2 ? 1 : 0) | (bits && !dec2 ? 2 : 0) | (bits && dec2 ? 4 : 0);
  bool        bits1 = false, bits2 = false, d2 = false;
  // flush the accumulated decisions of steps [p_end - acc_n, p_end) into the natural-order bit string
  auto flush_bits = [&](int p_end) {
    if (acc_n == 0)
      return;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const u32      v  = h ? (acc >> 16) : (acc & 0xffffu);
      const u32      r  = __brev(v) >> (32 - acc_n); // first step at bit 0
      const uint32_t n0 = (uint32_t)(2 * j + h) * (uint32_t)W + (uint32_t)(p_end - acc_n);
      const uint32_t sh = n0 & 31u;
      atomicOr(w.bits + (n0 >> 5), r << sh);
      if (sh + (uint32_t)acc_n > 32u)
        atomicOr(w.bits + (n0 >> 5) + 1, r >> (32u - sh));
    }
    acc   = 0;
    acc_n = 0;
  };

  // one forward step with output: LLR against b = beta_{p+1}, state update, glue epilogue (iter.h:107-127).
  // tp / tq: this lane's view of the staged plane rows / table rows at row 0 of the step's half tile; p: trellis step;
  // i: step within the half tile (compile-time; its parity is the step's)
  auto out_step = [&](const u32* tp, const u32* tq, int p, int i, const u32 (&b)[8], RangeMon& mon, bool norm) {
    u32 x, y;
    row(tp, i, x, y);
    const u32 xy = P::add(x, y);
    u32       llr;
    if (P::kMonitor) { // wrapping arithmetic under the range monitor: factored form
      llr = llr_factored<P>(al, b, x, y, xy, mon);
      fwd_step<P>(al, x, y, xy);
    } else { // saturating arithmetic: the operation order of the reference is part of the result
      if (PS::kNeedsFix && i == 1 && p == 1)
        llr = fwd_step_llr<P>(al, b, x, y, xy, mon); // un-normalised input: step 0 is not followed by a normalisation
      else
        llr = fwd_step_llr<PS>(al, b, x, y, xy, mon);
    }
    if (P::kMonitor && (i & 1) == 0)
      mon.track(al);
    if (P::kMonitor && i == 0) {
      // (step 0 of a half tile normalises in every tile but the very first: without a branch -- subtracting zero leaves the
      //  metrics as they are -- the two paths need no register moves to meet again)
      const u32 n0 = norm ? al[0] : 0u;
#pragma unroll
      for (int s = 0; s < 8; s++)
        al[s] = p_sub_wrap(al[s], n0);
    } else if ((kNP == 1 || (i & 1) == 0) && norm)
      P::normalize_now(al);
    // (two 16-bit loads, not one word taken apart with LOP3 + PRMT: the load / store unit has room, the max-type pipe does not)
    const unsigned  tq_s = (unsigned)__cvta_generic_to_shared(tq + i * T);
    const uint32_t  t0 = lds_u16(tq_s), t1 = lds_u16(tq_s + 2u);
    // decoder 1: extrinsic - a-priori -> app2[rev[.]]; decoder 2: a-posteriori - own input -> a-priori[fwd[.]]
    // (decoder 2 has a zero a-priori slice, so x is its own input)
    const u32 sub = d2 ? x : tp[2 * Lay::kPlaneWords + i * 32];
    u32       e;
    if (P::kBits == 8) { // srslte_vec_sub_bbb saturates below an element index that depends on the build (arith.cuh)
      const uint32_t w2 = 2u * (uint32_t)(p * T + j);
      e = P::glue_sub(llr, sub, (d2 ? t0 : w2) < d_sat, (d2 ? t1 : w2 + 1) < d_sat);
    } else {
      e = P::glue_sub(llr, sub, false, false);
    }
    if (P::kMonitor) { // running max / min of the extrinsic values, two steps per VIMNMX3
      if ((i & 1) == 0) {
        e_even = e;
      } else {
        ehi = p_max3(ehi, e_even, e);
        elo = p_min3(elo, e_even, e);
      }
    }
    if (bits1) {
      // decisions stay in this thread's lanes: accumulate (llr > 0 -> 1, one bit per int16 half), flush every 16 steps
      acc = acc * 2u + __vimin_s16x2_relu(llr, 0x00010001u);
      acc_n++;
    }
    if (bits2) {
      // decisions belong to natural positions nat[.]
      const u32       dbit = __vimin_s16x2_relu(llr, 0x00010001u);
      const uint32_t  n0 = lds_u16(tq_s + 4u * (unsigned)Lay::kLutWords), n1 = lds_u16(tq_s + 4u * (unsigned)Lay::kLutWords + 2u);
      atomicOr(reinterpret_cast<u32*>(bits_c + ((n0 >> 5) << 2)), __funnelshift_l(0u, dbit & 1u, n0));
      atomicOr(reinterpret_cast<u32*>(bits_c + ((n1 >> 5) << 2)), __funnelshift_l(0u, dbit >> 16, n1));
    }
    // (streamed: the extrinsic plane is read again a whole half-iteration later, by which time 85 MB of them are in flight;
    //  the checkpoints, which come back within the same half-iteration, get the L2 instead)
    stg16_hint(ext + 2u * t0, (uint16_t)(e & 0xffffu), w.pol_first);
    stg16_hint(ext + 2u * t1, (uint16_t)(e >> 16), w.pol_first);
  };
  auto ck_load = [&](int t, u32 (&v)[8]) { // checkpoint of alpha tile t: [half][lane][4 words]
    const uint4* c = reinterpret_cast<const uint4*>(w.sm + Lay::kCkRing + (t & 1) * 256) + lane;
    const uint4  lo = c[0], hi = c[32];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  };
  uint4* const ysp = reinterpret_cast<uint4*>(w.sm + Lay::kYOff) + lane; // beta spill: entry y at ysp[64 y], ysp[64 y + 32]

  const int n_full = W >> 3;
#pragma unroll 1
  for (int t = 0; t < n_full; t++) {
    asm volatile("" : "+r"(flags));
    d2    = DEC < 0 ? (flags & 1) != 0 : DEC == 1;
    bits1 = DEC != 1 && (flags & 2);
    bits2 = DEC != 0 && (flags & 4);
    const u32* tb = acquire(kAlpha);
    u32        bs[4][8];
    // ---- recompute beta_{8t+7} .. beta_{8t+1} from the checkpoint beta_{8t+8}: the upper three go to shared memory, the
    //      lower four stay in registers
    ck_load(t, st);
    if (8 * (t + 1) < W)
      P::normalize_now(st); // the recursion continued from the normalised value; beta[W] itself was never normalised
#pragma unroll
    for (int kk = 7; kk >= 1; kk--) {
      u32 x, y;
      row(tb, kk, x, y);
      bwd_step<PS>(st, x, y, P::add(x, y));
      if (PS::kNeedsFix && kk == 7 && !(8 * (t + 1) < W))
        PS::fix(st); // un-normalised input: beta[W]
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
    // ---- two trips through one 4-step body: steps 8t .. 8t+3 against beta_{8t+1} .. beta_{8t+4} (registers), then steps
    //      8t+4 .. 8t+7 against beta_{8t+5} .. beta_{8t+7} (spilled by this lane) and the checkpoint beta_{8t+8}
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      if (half) {
#pragma unroll
        for (int y = 0; y < 3; y++) {
          const uint4 lo = ysp[64 * y], hi = ysp[64 * y + 32];
          bs[y][0] = lo.x; bs[y][1] = lo.y; bs[y][2] = lo.z; bs[y][3] = lo.w;
          bs[y][4] = hi.x; bs[y][5] = hi.y; bs[y][6] = hi.z; bs[y][7] = hi.w;
        }
        ck_load(t, bs[3]);
      }
      const u32* tp = tb + half * 4 * 32;                               // plane rows 4 half .. 4 half + 3
      const u32* tq = tb - lane + kLutOff + j + half * 4 * T;           // table rows of the same steps
#pragma unroll
      for (int i = 0; i < 4; i++)
        out_step(tp, tq, 8 * t + 4 * half + i, i, bs[i], mon_a, (i | half | t) != 0);
      if (P::kMonitor && t == 0 && half == 0) { // what was tracked so far belongs to the head monitor
        mon_h.hi = p_max(mon_h.hi, mon_a.hi);
        mon_h.lo = p_min(mon_h.lo, mon_a.lo);
        mon_a.hi = 0;
        mon_a.lo = 0;
      }
    }
    if (t & 1)
      flush_bits(8 * t + 8); // (nothing accumulated unless decoder 1 needs its decisions)
  }
  if (W & 7) {
    // partial top tile, guarded: beta_{p+1} of each step is recomputed from the checkpoint beta[W] (at most 6 steps)
    const int  t  = n_full;
    const int  nv = W & 7;
    const u32* tb = acquire(kAlpha);
    d2    = DEC < 0 ? (flags & 1) != 0 : DEC == 1;
    bits1 = DEC != 1 && (flags & 2);
    bits2 = DEC != 0 && (flags & 4);
#pragma unroll 1
    for (int i = 0; i < nv; i++) {
      u32 b[8];
      ck_load(t, b);
#pragma unroll 1
      for (int kk = nv - 1; kk > i; kk--) { // -> beta_{8t+kk}, normalised on the way except the one that is used
        u32 x, y;
        row(tb, kk, x, y);
        bwd_step<PS>(b, x, y, P::add(x, y));
        if (PS::kNeedsFix && kk == nv - 1)
          PS::fix(b); // un-normalised input: beta[W]
        if (kk > i + 1 && (kNP == 1 || (kk & 1) == 0))
          P::normalize_now(b);
      }
      // (t >= 5 here: a lane has at least 40 steps.)  A runtime row index: the step parity that drives the normalisation
      // cadence and the monitor is i's, passed as the compile-time 0 / 1 of the two branches
      const int  r0 = i & ~1;
      const u32* tp = tb + r0 * 32;
      const u32* tq = tb - lane + kLutOff + j + r0 * T;
      if (i & 1)
        out_step(tp, tq, 8 * t + i, 1, b, mon_a, true);
      else
        out_step(tp, tq, 8 * t + i, 0, b, mon_a, true);
    }
  }
  flush_bits(W);

  ge_out = 0;
  if (P::kMonitor) {
    // max |extrinsic| handed to the next half-iteration (its a-priori / systematic input)
    ehi    = p_max(ehi, e_even); // (an odd number of steps leaves the last one pending; counting one twice is harmless)
    elo    = p_min(elo, e_even);
    int ge = max(max(lo16(ehi), hi16(ehi)), max(-lo16(elo), -hi16(elo)));
#pragma unroll
    for (int o = T / 2; o >= 1; o >>= 1)
      ge = max(ge, __shfl_xor_sync(gmask, ge, o, T));
    ge_out = ge;
    const bool bad = !fast16_alpha_ok(mon_a.spread_lo(), mon_b.spread_lo(), g) || !fast16_alpha_ok(mon_a.spread_hi(), mon_b.spread_hi(), g) ||
                     !fast16_alpha_ok(mon_h.spread_lo(), mon_b.spread_lo(), g) || !fast16_alpha_ok(mon_h.spread_hi(), mon_b.spread_hi(), g) ||
                     ((mon_a.ovf | mon_h.ovf) & 0x80008000u) != 0;
    if (__any_sync(gmask, bad && live))
      flagged = flagged || live;
  }
  // the extrinsic values were written through the generic proxy; the next half-iteration's tiles read them through the
  // async proxy
  asm volatile("fence.proxy.async.global;\n" ::: "memory");
  return flagged;
}

template <class P, int N>
__global__ void __launch_bounds__(384, 1) k_map_fused(const FusedArgs a)
{
  constexpr int T = N / 2, G = 32 / T;
  using Lay = FusedLay<T>;
  extern __shared__ __align__(128) u32 smem_f[];
  const int lane = threadIdx.x & 31;
  const int wib  = threadIdx.x >> 5;

  FusedWarp<T> w;
  w.sm        = smem_f + (size_t)wib * a.warp_words; // (one warp per CTA: a finished warp hands its share of the SM back)
  w.sm_s      = (unsigned)__cvta_generic_to_shared(w.sm);
  w.lane      = lane;
  w.j         = lane % T;
  w.gmask     = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane / T * T));
  w.wr_stage  = 0;
  w.rd_stage  = 0;
  w.rd_phase  = 0;
  w.pol_first = l2_policy_evict_first();
  w.pol_last  = l2_policy_evict_last();
  w.pol_ck_st = a.ck_policy ? l2_policy_evict_normal() : w.pol_last;
  w.pol_ck_ld = a.ck_policy ? l2_policy_evict_normal() : w.pol_first;
  w.ck_warp   = a.ck_scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + wib) * a.ck_words;
  w.bits      = w.sm + Lay::kBitsOff + (lane / T) * a.bits_words;
  const int j = w.j;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < Lay::kStages; i++)
      mbar_init(w.sm_s + 4u * (unsigned)(Lay::kBarOff + 2 * i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();

  const int  n_groups = a.mode == 2 ? (int)a.counters[2] : a.n_groups;
  const bool sliced   = a.slice_first > 0 && a.mode != 2;
  for (;;) {
    // (warp-wide reductions instead of shuffles wherever a value is the same for every lane: their results live in the
    //  uniform datapath, so the tile bookkeeping and the operands of the bulk copies need no per-lane code)
    int gi = 0;
    if (!sliced) {
      if (lane == 0)
        gi = (int)atomicAdd(&a.counters[a.ctr_fetch], 1u);
      gi = (int)__reduce_add_sync(0xffffffffu, (unsigned)gi);
      if (gi >= n_groups)
        break;
    } else {
      // tickets 0 .. n_groups-1 are the groups themselves, ticket n_groups + i is the i-th group handed back.  A ticket is
      // only taken against a credit (counters[ctr_avail] + n_groups = entries nobody has claimed yet), so no entry is left
      // behind by a warp that finds nothing and exits; every step is one fetch-add (a compare-and-swap loop on one address
      // with 1776 warps ending their slices together is quadratic: measured 7x slower than no slicing at all)
      if (lane == 0) {
        int* avail = reinterpret_cast<int*>(a.counters + a.ctr_avail);
        gi         = -1;
        for (;;) {
          if (atomicAdd(avail, -1) + n_groups > 0) {
            gi = (int)atomicAdd(&a.counters[a.ctr_fetch], 1u);
            break;
          }
          // No credit: give the decrement back.  Another warp's transient decrement may have hidden a credit from a third
          // one in the meantime (A finds nothing, B publishes an entry, B finds "nothing" because of A) -- whoever gives back
          // LAST sees the true count, so look again before leaving: no entry is ever stranded
          atomicAdd(avail, 1);
          if (*(volatile int*)avail + n_groups <= 0)
            break;
        }
        if (gi >= n_groups) {
          volatile int* slot_q = a.queue + (gi - n_groups);
          int           v;
          // (the entry is being published by the warp that took the matching credit: a few instructions away.  Bounded, so
          //  that a logic error is a trap and an error code instead of a hung GPU)
          for (uint32_t spins = 0; (v = *slot_q) < 0; spins++) {
            __nanosleep(64);
            if (spins > (1u << 22))
              __trap();
          }
          gi = v | 0x40000000;
        }
      }
      gi = (int)__reduce_add_sync(0xffffffffu, (unsigned)gi);
      if (gi < 0)
        break;
      if (gi & 0x40000000) {
        gi &= 0x3fffffff;
        __threadfence(); // acquire: the state and the extrinsic planes the previous owner wrote
        asm volatile("fence.proxy.async.global;\n" ::: "memory");
      }
    }
    const int grp  = a.mode == 2 ? a.parked[gi] : gi;
    const int slot = grp * G + lane / T;
    const int cb   = a.work[slot];
    bool      live = cb >= 0;
    uint32_t  d_W = 0, d_K = 0, d_ps = 0, d_qpp = 0, d_sat = 0, n_iter0 = 0, max_iter = 0, crc_poly = 0, out_off = 0;
    uint32_t  xq[4] = {0, 0, 0, 0};
    uint64_t  d_ws = 0;
    if (live) {
      const CbDev*   dp = a.cbs + cb;
      const CbState* sp = a.state + cb;
      d_W      = dp->W;
      d_K      = dp->K;
      d_ps     = dp->ps;
      d_qpp    = dp->qpp_off;
      d_sat    = dp->sat_end;
      d_ws     = dp->ws_off;
      max_iter = dp->max_iter;
      crc_poly = dp->crc_poly;
      out_off  = dp->out_off;
#pragma unroll
      for (int i = 0; i < 4; i++)
        xq[i] = dp->crc_xq[i];
      n_iter0 = sp->n_iter;
      // (mode 1 comes back to a group it handed back: the blocks it parked meanwhile wait for the exact launch)
      if (sp->done || n_iter0 >= max_iter || (a.mode == 2 && !sp->redo) || (a.mode == 1 && sp->redo))
        live = false;
    }
    const unsigned live_mask = __ballot_sync(0xffffffffu, live);
    if (live_mask == 0)
      continue;
    // geometry of the group: the first slot of a group is never padding and every block of a group has the same size
    const CbDev* d0   = a.cbs + a.work[grp * G];
    const int    W    = d0->W;
    const int    K    = (int)d0->K;
    const int    qoff = (int)d0->qpp_off;
    if (live && (int)d_W != W)
      __trap(); // host planning never mixes sizes in a group
    int16_t*        ws = a.ws + d_ws;
    const size_t    ps = d_ps;
    const int16_t*  tl = a.tails + (size_t)(live ? cb : 0) * 12;
    const uint16_t* q  = a.qpp + qoff;
    int*            gm = a.gmax + (size_t)(live ? cb : 0) * 4;
    // max |LLR| of the three input planes and of the extrinsic values handed to the next half-iteration (Fast16 monitor)
    const int g_syst = P::kMonitor ? gm[0] : 0, g_par0 = P::kMonitor ? gm[1] : 0, g_par1 = P::kMonitor ? gm[2] : 0;
    int       g_ext = P::kMonitor ? gm[3] : 0;
    const CUtensorMap* tmap3 = a.tmaps + 2 * a.winfo[2 * grp];
    const int          blk0  = a.winfo[2 * grp + 1];
    const uint32_t     nwords = (uint32_t)K >> 5, nbytes = (uint32_t)K >> 3;
    bool               parked_any = false;

    // The blocks of a group normally sit at the same half-iteration; blocks parked at different half-iterations (exact
    // launch) are taken in cohorts of equal count, because the constituent decoder is a property of the warp's code path
    bool pending = live, handed_back = false;
    for (;;) {
    const unsigned pmask = __ballot_sync(0xffffffffu, pending);
    if (pmask == 0)
      break;
    int niter = (int)__reduce_min_sync(0xffffffffu, pending ? n_iter0 : 0xffffffffu);
    live      = pending && (int)n_iter0 == niter;
    pending   = pending && !live;
    int       niter_first = niter;
    int       budget      = niter == 0 ? a.slice_first : a.slice_next;
    while (true) {
      // decisions are needed when a CRC check follows this half-iteration or the run ends with it
      const bool want_bits = __any_sync(0xffffffffu, live && (crc_poly != 0 || (uint32_t)niter + 1 >= max_iter));
      if (want_bits) {
        for (uint32_t i = j; i < (uint32_t)a.bits_words; i += T)
          w.bits[i] = 0;
        __syncwarp();
      }
      bool flagged;
      int  ge = 0;
      {
        // one call site, runtime flags: the body must exist once (see fused_half)
        const bool dec2 = niter & 1, apr3 = !dec2 && niter != 0;
        const int  g    = dec2 ? g_ext + g_par1 : (apr3 ? g_ext : 0) + g_syst + g_par0;
        const CUtensorMap* tm = apr3 ? tmap3 : tmap3 + 1;
        if (!P::kMonitor)
          flagged = fused_half<P, N, -1>(w, a, tm, blk0, W, K, q, ws, ps, tl, g, d_sat, live, ge, dec2, apr3, want_bits);
        else if (dec2)
          flagged = fused_half<P, N, 1>(w, a, tm, blk0, W, K, q, ws, ps, tl, g, d_sat, live, ge, true, false, want_bits);
        else
          flagged = fused_half<P, N, 0>(w, a, tm, blk0, W, K, q, ws, ps, tl, g, d_sat, live, ge, false, apr3, want_bits);
      }
      g_ext = ge;
      __syncwarp(); // the decision bits of every lane are in shared memory
      if (P::kMonitor && flagged) {
        // park: the half-iteration is replayed (and the block finished) by the exact-arithmetic launch
        if (j == 0) {
          a.state[cb].n_iter = (uint32_t)niter;
          a.state[cb].redo   = 1;
          if (niter > niter_first)
            atomicAdd(&a.counters[1], (uint32_t)(niter - niter_first));
        }
        live = false;
      }
      parked_any = parked_any || __any_sync(0xffffffffu, P::kMonitor && flagged);
      niter++;
      if (live && j == 0 && P::kMonitor)
        gm[3] = ge;
      const bool need = live && (crc_poly != 0 || (uint32_t)niter >= max_iter);
      uint32_t   crc  = 1;
      if (__any_sync(0xffffffffu, need && crc_poly != 0)) { // (every lane takes part: ghosts compute a value nobody uses)
        const bool     is_a = crc_poly == kCrc24A;
        const uint32_t c    = group_crc24<T>(w.bits, nbytes, a.crc_tab + (is_a ? 0 : 256), is_a ? kCrc24A : kCrc24B, xq, j, w.gmask);
        if (need && crc_poly != 0)
          crc = c;
      }
      const bool ok  = need && crc_poly != 0 && crc == 0; // early stop (sch.c:441-450)
      const bool fin = live && (ok || (uint32_t)niter >= max_iter);
      if (fin) {
        // decided bytes, first bit of the block = MSB of byte 0
        uint8_t* ob = a.cb_out + out_off;
        if (((out_off | nbytes) & 3u) == 0) {
          u32* o32 = reinterpret_cast<u32*>(ob);
          for (uint32_t i = j; i < nwords; i += T)
            o32[i] = __byte_perm(__brev(w.bits[i]), 0, 0x0123);
        } else {
          for (uint32_t i = j; i < nbytes; i += T)
            ob[i] = (uint8_t)(__brev((w.bits[i >> 2] >> (8u * (i & 3u))) & 0xffu) >> 24);
        }
        if (j == 0) {
          CbState* s = a.state + cb;
          s->n_iter  = (uint32_t)niter;
          s->crc     = crc;
          if (ok) {
            s->crc_ok = 1;
            s->done   = 1;
          }
          const uint32_t ran = (uint32_t)(niter - niter_first);
          atomicAdd(&a.counters[1], ran);
          if (a.mode == 2) {
            s->redo = 0;
            s->n_redo += ran;
            atomicAdd(&a.counters[0], ran);
          }
        }
        live = false;
      }
      if (!__any_sync(0xffffffffu, live))
        break;
      if (sliced && --budget <= 0) {
        // end of the slice: keep the group if nobody is waiting, else hand it back and take the oldest waiting one
        int waiting = 0;
        if (lane == 0)
          waiting = *(volatile int*)(a.counters + a.ctr_avail) + n_groups > 0;
        if (__reduce_add_sync(0xffffffffu, (unsigned)waiting) == 0) {
          budget = a.slice_next;
          continue;
        }
        if (live && j == 0) {
          a.state[cb].n_iter = (uint32_t)niter;
          atomicAdd(&a.counters[1], (uint32_t)(niter - niter_first));
        }
        handed_back = true;
        pending     = false;
        break;
      }
    }
    }
    if (a.mode == 1 && parked_any && lane == 0) {
      // (a sliced group can park blocks in several of its slices: one list entry per group)
      if (!sliced || atomicExch(&a.queue[a.queue_cap + grp], 0) == -1) {
        const uint32_t at = atomicAdd(&a.counters[2], 1u);
        a.parked[at]      = grp;
      }
    }
    if (handed_back) {
      __threadfence(); // release: this lane's extrinsic values and state
      __syncwarp();
      if (lane == 0) {
        const uint32_t at = atomicAdd(&a.counters[a.ctr_tail], 1u);
        *(volatile int*)(a.queue + at) = grp;
        __threadfence();
        atomicAdd(reinterpret_cast<int*>(a.counters + a.ctr_avail), 1); // the credit follows the entry
      }
    }
  }
}

#endif // __CUDACC__

} // namespace b200
