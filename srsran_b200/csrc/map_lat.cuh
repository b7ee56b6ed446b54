// k_map_lat -- the latency-shaped MAP kernel for batches that leave most of the GPU empty (one subframe or a few).
//
// k_map_f16 gives a group of G code blocks to ONE warp; alone on its SM sub-partition that warp is bound by its
// dependent-instruction chain (about 238 cycles per trellis step for 147 instructions, profiles/README.md), not by
// issue slots.  The same integers come out faster when the work of a group is spread over the four warps of a CTA:
//
//   phase 1   warp 0 runs the backward pass (warm-up, lane hand-over, tail trellis, main pass) and stores EVERY beta
//             vector; warp 1 runs, at the same time, the forward warm-up and the alpha recursion and stores every alpha
//             vector.  Both passes only depend on the inputs (turbodecoder_win.h:551-681 / 684-760).
//   phase 2   with alpha_p and beta_{p+1} known for every step, the a-posteriori LLRs and the half-iteration glue
//             (win.h:769-813, iter.h:104-128) of different steps are independent: the four warps take the 8-step tiles
//             round robin.
//
// Arithmetic (Fast16 under the range monitor, or the saturating int8 policy Sat8), normalisation points, tracking points
// and the soundness conditions are those of k_map_f16 (map_core.cuh): the alpha / beta vectors are simply stored instead of being consumed in place.
// Scratch per CTA: beta [0..W] then alpha [0..W-1], 1 KB per vector ([2 halves][32 lanes][4 words]), L2-resident.
#pragma once
#include "map_f16.cuh"

namespace b200 {

#if defined(__CUDACC__)

template <int T>
struct LatLay {
  static constexpr int kPlaneWords  = 8 * 32;                    // one plane of a tile: 8 box rows of 32 words
  static constexpr int kBoxWords    = 3 * kPlaneWords;           // up to three input planes
  static constexpr int kSmallStages = 8;                         // phase 1: box-only stages, deep prefetch for a lone warp
  static constexpr int kLutOff      = kBoxWords;                 // phase 2 stage: box | QPP rows [8][T] | alpha 8 x 256 | beta 8 x 256
  static constexpr int kAlOff       = kLutOff + 8 * T;
  static constexpr int kBeOff       = kAlOff + 8 * 256;
  static constexpr int kBigStage    = (kBeOff + 8 * 256 + 31) / 32 * 32;
  static constexpr int kBigStages   = 2;
  static constexpr int kDataWords   = kBigStages * kBigStage > kSmallStages * kBoxWords ? kBigStages * kBigStage : kSmallStages * kBoxWords;
  static constexpr int kBarOff      = kDataWords;                // kSmallStages mbarriers (64 bit each)
  static constexpr int kWarpWords   = (kBarOff + 2 * kSmallStages + 31) / 32 * 32;
};

// MODE: 0 = DEC1 without a-priori input, 1 = DEC1 with a-priori input, 2 = DEC2.  One CTA of four warps per group of G slots.
template <class P, int N, int MODE>
__global__ void __launch_bounds__(128, 1) k_map_lat(const MapArgs a)
{
  constexpr int  T = N / 2, G = 32 / T;
  constexpr int  kNP = P::kNormPeriod;
  constexpr bool kDec2 = MODE == 2, kApr = MODE == 1;
  using Lay = LatLay<T>;
  extern __shared__ __align__(128) u32 smem_f[];
  __shared__ u32 s_mon[32][6]; // alpha pass: mon_a.hi, mon_a.lo, mon_h.hi, mon_h.lo; [4] beta verdict, [5] unused
  __shared__ u32 s_out[4][32][3]; // phase 2, per warp: LLR-subtraction overflow bits, max / min extrinsic

  if (a.iter > 0 && a.counters[4 + a.iter - 1] == 0)
    return; // nothing left to decode in this batch
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;
  const int slot = blockIdx.x * G + lane / T;
  const int j    = lane % T;
  const int cb   = slot < a.n_slots ? a.work[slot] : -1;
  bool      live = cb >= 0;
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
  const unsigned live_mask = __ballot_sync(0xffffffffu, live); // the same in all four warps: a uniform exit
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
      __trap();
  }
  const unsigned gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane / T * T));

  int16_t*        ws = a.ws + d_ws;
  const size_t    ps = d_ps;
  const int16_t*  tl = a.tails + (size_t)(live ? cb : 0) * 12;
  const uint16_t* q  = a.qpp + qoff;

  u32*           sm   = smem_f + role * Lay::kWarpWords;
  const unsigned sm_s = (unsigned)__cvta_generic_to_shared(sm);
  const u32*     my   = sm + lane;
  const CUtensorMap* tmap   = a.tmaps + 2 * a.winfo[2 * blockIdx.x] + (kApr ? 0 : 1);
  const int          blk0   = a.winfo[2 * blockIdx.x + 1];
  constexpr int      plane0 = kDec2 ? kPlApp2 : kPlSyst;
  constexpr unsigned kBoxBytes = (kApr ? 3u : 2u) * 8u * 128u;
  const u32*         lut  = (const u32*)(kDec2 ? q : q + K);
  u32* const         be_g = a.ck_scratch + (size_t)blockIdx.x * (size_t)a.ck_slots * 512; // beta of step p at + 256 p
  u32* const         al_g = be_g + (size_t)a.ck_slots * 256;                               // alpha before step p at + 256 p

  const int nT  = (W + 7) >> 3;
  const int a0  = (W - kWinOverlap) >> 3;
  const int nAW = nT - a0;
  const uint64_t pol_first = l2_policy_evict_first();
  auto bar_of = [&](int stage) -> unsigned { return sm_s + 4u * (unsigned)(Lay::kBarOff + 2 * stage); };
  unsigned rd_phase = 0; // bit s = parity the consumer waits for on barrier s (carried over from phase 1 to phase 2)
  auto row = [&](const u32* tb, int i, u32& x, u32& y) {
    const u32 vin = tb[i * 32];
    y             = tb[Lay::kPlaneWords + i * 32];
    x             = kApr ? P::add(tb[2 * Lay::kPlaneWords + i * 32], vin) : vin;
  };
  auto vec_store = [&](u32* base, int p, const u32 (&v)[8]) {
    if (live) {
      uint4* g = reinterpret_cast<uint4*>(base + (size_t)p * 256) + lane;
      g[0]  = make_uint4(v[0], v[1], v[2], v[3]);
      g[32] = make_uint4(v[4], v[5], v[6], v[7]);
    }
  };

  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < Lay::kSmallStages; i++)
      mbar_init(bar_of(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();

  RangeMon mon_b, mon_a, mon_h;
  mon_b.reset();
  mon_a.reset();
  mon_h.reset();
  int g = 0;
  int* gm = a.gmax + (size_t)(live ? cb : 0) * 4;

  // =============================================================== phase 1: the two recursions, side by side
  if (role < 2) {
    // tile sequence of this warp: role 0 beta warm-up (tiles 4..0) + beta main (top..0); role 1 alpha warm-up (a0..top) + main (0..top)
    const int n_seq = role == 0 ? 5 + nT : nAW + nT;
    int       wr_idx = 0, wr_stage = 0, rd_stage = 0;
    auto issue = [&]() {
      __syncwarp();
      if (lane == 0 && wr_idx < n_seq) {
        const int t = role == 0 ? (wr_idx < 5 ? 4 - wr_idx : nT - 1 - (wr_idx - 5)) : (wr_idx < nAW ? a0 + wr_idx : wr_idx - nAW);
        const unsigned bar = bar_of(wr_stage);
        mbar_expect_tx(bar, kBoxBytes);
        tma_tile4_hint(sm_s + 4u * (unsigned)(wr_stage * Lay::kBoxWords), tmap, 0, blk0, 8 * t, plane0, bar, pol_first);
      }
      wr_idx++;
      wr_stage = wr_stage + 1 == Lay::kSmallStages ? 0 : wr_stage + 1;
    };
    auto acquire = [&]() -> const u32* {
      issue();
      mbar_wait(bar_of(rd_stage), (rd_phase >> rd_stage) & 1u);
      rd_phase ^= 1u << rd_stage;
      const u32* tb = my + rd_stage * Lay::kBoxWords;
      rd_stage      = rd_stage + 1 == Lay::kSmallStages ? 0 : rd_stage + 1;
      return tb;
    };
#pragma unroll
    for (int i = 0; i < Lay::kSmallStages - 1; i++)
      issue();
    u32 st[8];
#pragma unroll
    for (int s = 0; s < 8; s++)
      st[s] = splat16(-P::kInf);

    if (role == 0) {
      // ---- backward warm-up: steps 39..0 of the lane's own sub-block (win.h:622-630)
      for (int t = 4; t >= 0; t--) {
        const u32* tb = acquire();
#pragma unroll
        for (int i = 7; i >= 0; i--) {
          u32 x, y;
          row(tb, i, x, y);
          bwd_step<P>(st, x, y, P::add(x, y));
          if (P::kMonitor && (i & 1) == 0 && (t < 4 || i < 6))
            mon_b.track(st);
          if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
            P::normalize_now(st);
        }
      }
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
      if (P::kMonitor)
        mon_b.track(st);
      vec_store(be_g, W, st);
      // ---- backward main pass: beta[p] stored before normalisation, as the LLR consumes it
      int t = nT - 1;
      if (W & 7) {
        const u32* tb = acquire();
#pragma unroll 1
        for (int i = (W & 7) - 1; i >= 0; i--) {
          u32 x, y;
          row(tb, i, x, y);
          bwd_step<P>(st, x, y, P::add(x, y));
          vec_store(be_g, 8 * t + i, st);
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
          vec_store(be_g, 8 * t + i, st);
          if (P::kMonitor && (i & 1) == 0)
            mon_b.track(st);
          if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
            P::normalize_now(st);
        }
      }
    } else {
      // ---- forward warm-up: steps W-40..W-1 of the lane's own sub-block (win.h:747-756)
      int kk = 0;
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
      if (P::kMonitor)
        mon_h.track(st);
      // ---- alpha recursion: alpha before step p stored; tracking and normalisation points of k_map_f16's output pass
      const int n_full = W >> 3;
      for (int t = 0; t < n_full; t++) {
        const u32* tb = acquire();
#pragma unroll
        for (int i = 0; i < 8; i++) {
          vec_store(al_g, 8 * t + i, st);
          u32 x, y;
          row(tb, i, x, y);
          fwd_step<P>(st, x, y, P::add(x, y));
          if (P::kMonitor && (i & 1) == 0)
            mon_a.track(st);
          if ((kNP == 1 || (i & 1) == 0) && (i != 0 || t != 0))
            P::normalize_now(st);
          if (P::kMonitor && t == 0 && i == 3) { // what was tracked so far belongs to the head monitor
            mon_h.hi = p_max(mon_h.hi, mon_a.hi);
            mon_h.lo = p_min(mon_h.lo, mon_a.lo);
            mon_a.hi = 0;
            mon_a.lo = 0;
          }
        }
      }
      if (W & 7) {
        const u32* tb = acquire();
#pragma unroll 1
        for (int i = 0; i < (W & 7); i++) {
          vec_store(al_g, 8 * n_full + i, st);
          u32 x, y;
          row(tb, i, x, y);
          fwd_step<P>(st, x, y, P::add(x, y));
          if (P::kMonitor && (i & 1) == 0)
            mon_a.track(st);
          if (kNP == 1 || (i & 1) == 0)
            P::normalize_now(st);
        }
      }
    }
    // the vectors were written through the generic proxy and come back through the async proxy (bulk copies of phase 2)
    asm volatile("fence.proxy.async.global;\n" ::: "memory");
    __threadfence_block();
    if (role == 0) {
      bool bad = false;
      if (P::kMonitor) {
        g   = kDec2 ? gm[3] + gm[2] : (kApr ? gm[3] : 0) + gm[0] + gm[1];
        bad = !fast16_beta_ok(mon_b.spread_lo(), g) || !fast16_beta_ok(mon_b.spread_hi(), g);
      }
      s_mon[lane][4] = __any_sync(gmask, bad && live) ? 1u : 0u;
    } else {
      s_mon[lane][0] = mon_a.hi;
      s_mon[lane][1] = mon_a.lo;
      s_mon[lane][2] = mon_h.hi;
      s_mon[lane][3] = mon_h.lo;
    }
  }
  __syncthreads();
  const bool beta_bad = s_mon[lane][4] != 0; // the whole code block is replayed with the exact policy (mode 2 launch)
  if (beta_bad)
    live = false;

  // =============================================================== phase 2: LLRs + glue, tiles round robin over the warps
  u32* const     post   = (u32*)(ws + kPlPost * ps);
  int16_t* const post16 = ws + kPlPost * ps;
  int16_t* const ext    = kDec2 ? ws + kPlApr * ps : ws + kPlApp2 * ps;
  const bool     write_post = (a.mode & kMapSkipPost) == 0;
  u32            ehi = 0, elo = 0;
  RangeMon       mon_l;
  mon_l.reset();
  {
    int wr_t = role, wr_stage = 0, rd_stage = 0;
    auto issue = [&]() {
      __syncwarp();
      if (lane == 0 && wr_t < nT) {
        const int      t   = wr_t;
        const unsigned bar = bar_of(wr_stage);
        const unsigned dst = sm_s + 4u * (unsigned)(wr_stage * Lay::kBigStage);
        const int      r1  = (8 * t + 8) < W ? 8 : W - 8 * t;
        const unsigned lut_bytes = (unsigned)r1 * T * 4u, vec_bytes = (unsigned)r1 * 1024u;
        mbar_expect_tx(bar, kBoxBytes + lut_bytes + 2u * vec_bytes);
        tma_tile4_hint(dst, tmap, 0, blk0, 8 * t, plane0, bar, pol_first);
        bulk_g2s(dst + 4u * (unsigned)Lay::kLutOff, lut + (size_t)8 * t * T, lut_bytes, bar);
        bulk_g2s_hint(dst + 4u * (unsigned)Lay::kAlOff, al_g + (size_t)(8 * t) * 256, vec_bytes, bar, pol_first);
        bulk_g2s_hint(dst + 4u * (unsigned)Lay::kBeOff, be_g + (size_t)(8 * t + 1) * 256, vec_bytes, bar, pol_first);
      }
      wr_t += 4;
      wr_stage ^= 1;
    };
    auto vec_load = [&](const u32* tb, int off, int i, u32 (&v)[8]) {
      const uint4* c = reinterpret_cast<const uint4*>(tb - lane + off + i * 256) + lane;
      const uint4  lo = c[0], hi = c[32];
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
      v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    };
    issue();
    for (int t = role; t < nT; t += 4) {
      issue();
      mbar_wait(bar_of(rd_stage), (rd_phase >> rd_stage) & 1u);
      rd_phase ^= 1u << rd_stage;
      const u32* tb = my + rd_stage * Lay::kBigStage;
      rd_stage ^= 1;
      const int r1 = (8 * t + 8) <= W ? 8 : W - 8 * t;
#pragma unroll 2
      for (int i = 0; i < r1; i++) {
        u32 al[8], b[8], x, y;
        vec_load(tb, Lay::kAlOff, i, al);
        vec_load(tb, Lay::kBeOff, i, b);
        row(tb, i, x, y);
        const u32 xy  = P::add(x, y);
        u32 llr;
        if (P::kMonitor) // wrapping arithmetic under the range monitor: factored form
          llr = llr_factored<P>(al, b, x, y, xy, mon_l);
        else // saturating arithmetic: the operation order of the reference is part of the result (the state update is discarded)
          llr = fwd_step_llr<P>(al, b, x, y, xy, mon_l);
        const uint16_t* r16 = reinterpret_cast<const uint16_t*>(tb + (Lay::kLutOff + i * T + j - lane));
        const uint32_t  t0 = r16[0], t1 = r16[1];
        u32 e;
        if (!kDec2) {
          const uint32_t w2 = 2u * (uint32_t)((8 * t + i) * T + j);
          e = kApr ? P::glue_sub(llr, tb[2 * Lay::kPlaneWords + i * 32], w2 < d_sat, w2 + 1 < d_sat) : llr;
          if (live && write_post)
            post[(8 * t + i) * T + j] = llr;
          if (live) {
            ext[t0] = (int16_t)lo16(e);
            ext[t1] = (int16_t)hi16(e);
          }
        } else {
          e = P::glue_sub(llr, x, t0 < d_sat, t1 < d_sat);
          if (live) {
            ext[t0] = (int16_t)lo16(e);
            ext[t1] = (int16_t)hi16(e);
          }
          if (live && write_post) {
            post16[t0] = (int16_t)lo16(llr);
            post16[t1] = (int16_t)hi16(llr);
          }
        }
        ehi = p_max(ehi, e);
        elo = p_min(elo, e);
      }
    }
  }
  s_out[role][lane][0] = mon_l.ovf;
  s_out[role][lane][1] = ehi;
  s_out[role][lane][2] = elo;
  __syncthreads();
  if (role == 0 && P::kMonitor) {
    u32 ovf = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      ovf |= s_out[r][lane][0];
      ehi = p_max(ehi, s_out[r][lane][1]);
      elo = p_min(elo, s_out[r][lane][2]);
    }
    mon_a.hi = s_mon[lane][0];
    mon_a.lo = s_mon[lane][1];
    mon_h.hi = s_mon[lane][2];
    mon_h.lo = s_mon[lane][3];
    int ge = max(max(lo16(ehi), hi16(ehi)), max(-lo16(elo), -hi16(elo)));
#pragma unroll
    for (int o = T / 2; o >= 1; o >>= 1)
      ge = max(ge, __shfl_xor_sync(gmask, ge, o, T));
    const bool bad = !fast16_alpha_ok(mon_a.spread_lo(), mon_b.spread_lo(), g) || !fast16_alpha_ok(mon_a.spread_hi(), mon_b.spread_hi(), g) ||
                     !fast16_alpha_ok(mon_h.spread_lo(), mon_b.spread_lo(), g) || !fast16_alpha_ok(mon_h.spread_hi(), mon_b.spread_hi(), g) ||
                     (ovf & 0x80008000u) != 0;
    const bool was_live = cb >= 0 && (live || beta_bad); // blocks that entered this launch live
    if (__any_sync(gmask, (bad && live) || (beta_bad && was_live))) {
      if (j == 0 && was_live)
        a.state[cb].redo = 1;
      return;
    }
    if (j == 0 && live)
      gm[3] = ge;
  }
}

#endif // __CUDACC__

} // namespace b200
