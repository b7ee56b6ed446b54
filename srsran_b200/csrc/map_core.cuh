// Windowed max-log-MAP for ONE thread = two adjacent sub-block lanes (int16x2) of one code block.
//
// Replaces tdec_win*_beta / tdec_win*_alpha (reference include/srslte/phy/fec/turbodecoder_win.h:551-681,
// 684-832) with a B200-shaped schedule that produces the same integers:
//
//   * the reference stores all 8*K beta metrics of a code block (98 KB at K=6144) and reads them back in the
//     alpha pass.  Here the backward pass keeps only one 8-state CHECKPOINT every L steps (shared memory) and
//     the forward pass RECOMPUTES each L-step beta segment from its checkpoint into registers just before it
//     consumes it.  The recursion is deterministic, so the recomputed values are the stored ones.
//   * segments are aligned to the END of the sub-block (top = W - m*L), so register arrays are indexed
//     statically and only the first segment of the forward pass can be partial.
//
// The code is __host__ __device__: tests/host_emul runs exactly this schedule on the CPU against the oracle.
#pragma once
#include "arith.cuh"

namespace b200 {

constexpr int kWinOverlap = 40; // win_overlap_len (turbodecoder_win.h:54)

// backward (beta) trellis step, turbodecoder_win.h:643-664
template <class P>
B200_HD void bwd_step(u32 (&o)[8], u32 x, u32 y, u32 xy)
{
  const u32 n0 = P::addmax(o[4], xy, o[0]);
  const u32 n1 = P::addmax(o[0], xy, o[4]);
  const u32 n2 = P::addmax2(o[5], y, o[1], x);
  const u32 n3 = P::addmax2(o[5], x, o[1], y);
  const u32 n4 = P::addmax2(o[6], x, o[2], y);
  const u32 n5 = P::addmax2(o[6], y, o[2], x);
  const u32 n6 = P::addmax(o[3], xy, o[7]);
  const u32 n7 = P::addmax(o[7], xy, o[3]);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}

// forward (alpha) trellis step without output (warm-up), turbodecoder_win.h:769-785, 820-822
template <class P>
B200_HD void fwd_step(u32 (&o)[8], u32 x, u32 y, u32 xy)
{
  const u32 n0 = P::addmax(o[1], xy, o[0]);
  const u32 n1 = P::addmax2(o[3], y, o[2], x);
  const u32 n2 = P::addmax2(o[4], y, o[5], x);
  const u32 n3 = P::addmax(o[6], xy, o[7]);
  const u32 n4 = P::addmax(o[0], xy, o[1]);
  const u32 n5 = P::addmax2(o[2], y, o[3], x);
  const u32 n6 = P::addmax2(o[5], y, o[4], x);
  const u32 n7 = P::addmax(o[7], xy, o[6]);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}

// forward step with a-posteriori LLR against beta_{k+1} (turbodecoder_win.h:769-822)
struct RangeMon;
template <class P, class Mon>
B200_HD u32 fwd_step_llr(u32 (&o)[8], const u32 (&b)[8], u32 x, u32 y, u32 xy, Mon& mon)
{
  // branch candidates: mb = hypothesis 0, nw = hypothesis 1
  const u32 mb0 = o[0], mb1 = P::cand(o[3], y), mb2 = P::cand(o[4], y), mb3 = o[7];
  const u32 mb4 = o[1], mb5 = P::cand(o[2], y), mb6 = P::cand(o[5], y), mb7 = o[6];
  const u32 nw0 = P::cand(o[1], xy), nw1 = P::cand(o[2], x), nw2 = P::cand(o[5], x), nw3 = P::cand(o[6], xy);
  const u32 nw4 = P::cand(o[0], xy), nw5 = P::cand(o[3], x), nw6 = P::cand(o[4], x), nw7 = P::cand(o[7], xy);
  u32 m1 = P::sum0(b[0], nw0), m0 = P::sum0(b[0], mb0);
  m1 = P::summax(b[1], nw1, m1); m0 = P::summax(b[1], mb1, m0);
  m1 = P::summax(b[2], nw2, m1); m0 = P::summax(b[2], mb2, m0);
  m1 = P::summax(b[3], nw3, m1); m0 = P::summax(b[3], mb3, m0);
  m1 = P::summax(b[4], nw4, m1); m0 = P::summax(b[4], mb4, m0);
  m1 = P::summax(b[5], nw5, m1); m0 = P::summax(b[5], mb5, m0);
  m1 = P::summax(b[6], nw6, m1); m0 = P::summax(b[6], mb6, m0);
  m1 = P::summax(b[7], nw7, m1); m0 = P::summax(b[7], mb7, m0);
  m1 = P::sumfin(m1);
  m0 = P::sumfin(m0);
  o[0] = P::max(mb0, nw0); o[1] = P::max(mb1, nw1); o[2] = P::max(mb2, nw2); o[3] = P::max(mb3, nw3);
  o[4] = P::max(mb4, nw4); o[5] = P::max(mb5, nw5); o[6] = P::max(mb6, nw6); o[7] = P::max(mb7, nw7);
  const u32 d = P::sub(m1, m0);
  if (P::kMonitor)
    mon.track_sub(m1, m0, d);
  return P::out(d);
}

// 3-step tail termination for the last lane (tdec_win*_beta_trellis, turbodecoder_win.h:500-548); scalar.
// tin/tpar: the 3 tail LLRs of this constituent code (k = K, K+1, K+2)
template <class P>
B200_HD void tail_trellis(const int16_t* tin, const int16_t* tpar, int32_t (&o)[8])
{
  o[0] = 0;
  for (int i = 1; i < 8; i++)
    o[i] = -P::kInf;
  for (int k = 2; k >= 0; k--) {
    const int32_t x = tin[k], y = tpar[k], xy = P::tail_add(x, y);
    int32_t       mb[8], nw[8];
    mb[0] = P::tail_add(o[4], xy); mb[1] = o[4];                  mb[2] = P::tail_add(o[5], y);  mb[3] = P::tail_add(o[5], x);
    mb[4] = P::tail_add(o[6], x);  mb[5] = P::tail_add(o[6], y);  mb[6] = o[7];                  mb[7] = P::tail_add(o[7], xy);
    nw[0] = o[0];                  nw[1] = P::tail_add(o[0], xy); nw[2] = P::tail_add(o[1], x);  nw[3] = P::tail_add(o[1], y);
    nw[4] = P::tail_add(o[2], y);  nw[5] = P::tail_add(o[2], x);  nw[6] = P::tail_add(o[3], xy); nw[7] = o[3];
    for (int i = 0; i < 8; i++)
      o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
  }
}

// Range monitor of the Fast16 policy (per thread = per pair of lanes, each int16 half tracked on its own).
//
// Tracked: max / min of the 8 path metrics BEFORE normalisation at every even step of each pass (the steps at
// which the reference normalises), of the states handed over between passes, and a sticky overflow bit of the
// final LLR subtraction.  With H = max(hi, 0), Lw = min(lo, 0), Sp = H - Lw and g >= every |branch metric| of
// the call, exact (unbounded) arithmetic satisfies: normalisation differences lie within +-Sp (both operands are
// tracked values); metrics at the untracked odd steps within +-(Sp + g); every candidate alpha/beta + gamma
// within +-(Sp + 2g); every LLR operand beta_i + candidate_i within +-(Sp_beta + Sp_alpha + 3g).  If those
// bounds stay inside int16 no saturating operation of the reference saturates and no wrapping operation here
// wraps, i.e. both compute the same integers.  The first two warm-up steps (all metrics equal to -INF, spread 0)
// are exempt from tracking and covered by g <= kMonGmax.  See DESIGN.md "Fast path soundness".
struct RangeMon {
  u32 hi, lo; // packed running max / min
  u32 ovf;    // sticky sign bits of LLR-subtraction overflows
  B200_HD void reset()
  {
    hi  = 0;
    lo  = 0;
    ovf = 0;
  }
  B200_HD void track(const u32 (&o)[8])
  {
    hi = p_max3(hi, o[0], o[1]);
    lo = p_min3(lo, o[0], o[1]);
    hi = p_max3(hi, o[2], o[3]);
    lo = p_min3(lo, o[2], o[3]);
    hi = p_max3(hi, o[4], o[5]);
    lo = p_min3(lo, o[4], o[5]);
    hi = p_max3(hi, o[6], o[7]);
    lo = p_min3(lo, o[6], o[7]);
  }
  // d = a - b (wrapping): overflow iff the operands differ in sign and the result's sign differs from a's
  B200_HD void track_sub(u32 a, u32 b, u32 d) { ovf |= (a ^ b) & (a ^ d); }
  B200_HD int32_t spread_lo() const { return lo16(hi) - lo16(lo); } // hi >= 0 >= lo by construction
  B200_HD int32_t spread_hi() const { return hi16(hi) - hi16(lo); }
};
constexpr int kMonGmax = 4500; // exemption bound for the un-tracked first warm-up steps: INF + 4g and 7g fit int16

// g: bound on every |branch metric| of this MAP call.  sp_*: RangeMon spreads of the beta / alpha passes of ONE lane.
B200_HD bool fast16_beta_ok(int32_t sp_b, int32_t g) { return g <= kMonGmax && sp_b + 2 * g <= 32767; }
B200_HD bool fast16_alpha_ok(int32_t sp_a, int32_t sp_b, int32_t g)
{
  return g <= kMonGmax && sp_a + 2 * g <= 32767 && sp_a + sp_b + 3 * g <= 32767;
}

// ---------------------------------------------------------------------------------------------- row sources
// A MAP call reads its three input rows (systematic-or-extrinsic, a-priori, parity) in CHUNKS of L consecutive
// trellis steps following one global schedule (see MapWin::chunk_rows).  A row source serves those chunks:
//   prefetch(c, p0, lo, hi)  start fetching rows max(p0,lo) <= p < min(p0+L,hi) of chunk c (may be a no-op)
//   wait(c)                  make chunk c readable
//   get(c, i, p, ...)        raw words of row p = p0 + i of chunk c
// DirectSrc reads global memory at the point of use (host emulation, reference for the staged source).
struct DirectSrc {
  static constexpr int kAhead = 1;
  const u32 *in, *apr, *par;
  int        T, j;
  u32*       ck;  // checkpoint store of this thread: 8 words per slot
  const u32* lut; // QPP table words handed to the epilogue (may be nullptr)
  B200_HD u32 get_lut(int, int, int p) const { return lut ? lut[p * T + j] : 0u; }
  B200_HD void prefetch(int, int, int, int, int) {}
  B200_HD void prefetch_none() {}
  B200_HD void wait(int) {}
  B200_HD void drain() {}
  B200_HD u32 get_apr(int, int, int p) const { return apr ? apr[p * T + j] : 0u; }
  B200_HD void get(int, int, int p, u32& vin, u32& vapr, u32& vpar) const
  {
    vin  = in[p * T + j];
    vpar = par[p * T + j];
    vapr = apr ? apr[p * T + j] : 0u;
  }
  B200_HD void ck_put(int slot, const u32 (&st)[8])
  {
    for (int s = 0; s < 8; s++)
      ck[slot * 8 + s] = st[s];
  }
  // checkpoint `slot`, which was requested together with chunk c
  B200_HD void ck_get(int, int slot, u32 (&st)[8]) const
  {
    for (int s = 0; s < 8; s++)
      st[s] = ck[slot * 8 + s];
  }
};

// One thread's view of a MAP call.  Rows of the lane-layout arrays are T = N/2 words; this thread owns word j.
template <class P, int L, class Src>
struct MapWin {
  Src      src;
  int      W;   // steps per lane (K / N)
  RangeMon mon_b, mon_a; // Fast16 only

  static constexpr int kWarmChunks = (kWinOverlap + L - 1) / L;

  // Global chunk schedule: beta warm-up (descending), beta main (descending), alpha warm-up (ascending), alpha
  // main (ascending).  Chunks are top-aligned so only the lowest chunk of a pass can be partial.
  B200_HD int n_chunks() const { return 2 * kWarmChunks + 2 * ((W + L - 1) / L); }
  // ck_slot: checkpoint the chunk's consumer needs (alpha main: beta[top] of its segment), else -1
  B200_HD void chunk_rows(int c, int& p0, int& lo, int& hi, int& ck_slot) const
  {
    const int S = (W + L - 1) / L;
    ck_slot     = -1;
    if (c < kWarmChunks) {
      p0 = kWinOverlap - (c + 1) * L;
      lo = 0;
      hi = kWinOverlap;
    } else if (c < kWarmChunks + S) {
      p0 = W - (c - kWarmChunks + 1) * L;
      lo = 0;
      hi = W;
    } else if (c < 2 * kWarmChunks + S) {
      p0 = W - kWinOverlap + (c - kWarmChunks - S) * L;
      lo = W - kWinOverlap;
      hi = W;
    } else {
      const int m = S - 1 - (c - 2 * kWarmChunks - S);
      p0      = W - (m + 1) * L;
      lo      = 0;
      hi      = W;
      ck_slot = m;
    }
  }
  B200_HD void prefetch_chunk(int c)
  {
    if (c < n_chunks()) {
      int p0, lo, hi, slot;
      chunk_rows(c, p0, lo, hi, slot);
      src.prefetch(c, p0, lo, hi, slot);
    } else {
      src.prefetch_none();
    }
  }
  // Chunk protocol: begin_chunk(c) starts the next chunk's transfer and makes chunk c readable; row(c, i, p, x, y)
  // then yields x (systematic + a-priori) and y (parity) of row p = p0 + i at the point of use (the staged source
  // reads shared memory, so nothing has to be parked in registers for a whole chunk).
  //
  // kFast variants below: INTERIOR chunk with even p0 -- all L rows valid, every k > 0 and the parity of k equals
  // the parity of i, so the unrolled body carries no per-step guards, compares or branches.  Everything else (the
  // lowest chunk of a pass, odd W, the warm-up passes) runs the guarded variant; both compute the same values.
  B200_HD void begin_chunk(int c)
  {
    prefetch_chunk(c + Src::kAhead); // keep kAhead chunks of memory traffic in flight behind the arithmetic
    src.wait(c);
  }
  B200_HD void row(int c, int i, int p, u32& x, u32& y) const
  {
    u32 vin, vapr;
    src.get(c, i, p, vin, vapr, y); // vapr == 0 when the call has no a-priori input
    x = P::add(vapr, vin);
  }

  B200_HD void begin()
  {
    mon_b.reset();
    mon_a.reset();
    for (int c = 0; c < Src::kAhead; c++)
      prefetch_chunk(c);
  }
  B200_HD bool interior(int p0, int min_p) const { return (L % 2 == 0) && p0 >= min_p && (p0 & 1) == 0 && p0 > 0; }

  // beta pass 0: steps 39..0 of the lane's own sub-block from the all-"unknown" state (win.h:622-630)
  B200_HD void beta_warm(u32 (&st)[8])
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
      st[i] = splat16(-P::kInf);
    for (int c = 0; c < kWarmChunks; c++) {
      const int p0 = kWinOverlap - (c + 1) * L;
      begin_chunk(c);
#pragma unroll
      for (int i = L - 1; i >= 0; i--) {
        const int k = p0 + i;
        if (k >= 0) {
          u32 x, y;
          row(c, i, k, x, y);
          bwd_step<P>(st, x, y, P::add(x, y));
          if (P::kMonitor && (k & 1) == 0 && k < kWinOverlap - 2)
            mon_b.track(st);
          P::normalize((uint32_t)k, st);
        }
      }
    }
  }

  template <bool kFast>
  B200_HD void beta_chunk(int m, u32 (&st)[8])
  {
    const int p0 = W - (m + 1) * L; // may be negative for the last (lowest) segment
    const int c  = kWarmChunks + m;
    begin_chunk(c);
#pragma unroll
    for (int i = L - 1; i >= 0; i--) {
      const int k = p0 + i;
      if (kFast || k >= 0) {
        u32 x, y;
        row(c, i, k, x, y);
        bwd_step<P>(st, x, y, P::add(x, y));
        if (i == 0 && (kFast || k > 0))
          src.ck_put(m + 1, st); // beta[k] before normalisation (win.h:666-678)
        if (kFast) {
          if (P::kMonitor && (i & 1) == 0)
            mon_b.track(st);
          if (P::kNormPeriod == 1 || (i & 1) == 0)
            P::normalize_now(st);
        } else {
          if (P::kMonitor && (k & 1) == 0)
            mon_b.track(st);
          P::normalize((uint32_t)k, st);
        }
      }
    }
  }

  // beta pass 1: st = initial state at step W (neighbour estimate or tail); stores one checkpoint per segment
  B200_HD void beta_main(u32 (&st)[8])
  {
    const int S = (W + L - 1) / L;
    if (P::kMonitor)
      mon_b.track(st); // the state handed over (neighbour's estimate or tail)
    src.ck_put(0, st); // slot 0 = beta[W]
    for (int m = 0; m < S; m++) {
      if (interior(W - (m + 1) * L, 0))
        beta_chunk<true>(m, st);
      else
        beta_chunk<false>(m, st);
    }
  }

  // alpha pass 0: steps W-40..W-1 of the lane's own sub-block (win.h:747-756)
  B200_HD void alpha_warm(u32 (&st)[8])
  {
    const int S = (W + L - 1) / L;
#pragma unroll
    for (int i = 0; i < 8; i++)
      st[i] = splat16(-P::kInf);
    for (int q = 0; q < kWarmChunks; q++) {
      const int p0 = W - kWinOverlap + q * L;
      const int c  = kWarmChunks + S + q;
      begin_chunk(c);
#pragma unroll
      for (int i = 0; i < L; i++) {
        const int k = q * L + i; // loop counter of the pass
        if (k < kWinOverlap) {
          u32 x, y;
          row(c, i, p0 + i, x, y);
          fwd_step<P>(st, x, y, P::add(x, y));
          if (P::kMonitor && (k & 1) == 0 && k > 2)
            mon_a.track(st);
          P::normalize((uint32_t)k, st);
        }
      }
    }
  }

  // alpha pass 1 with output.  epi(p, llr, x, apr, lut) is called once per step in ascending p.
  //
  // Fast16: the first kExactHead steps run with the exact saturating policy.  Lane 0 starts from the known state
  // [0, -INF x 7]; after 3 steps every state is reachable from state 0 and the -INF remnants are gone, so the
  // range monitor (which starts with the state entering step kExactHead) is not inflated by INF.
  static constexpr int kExactHead = 4;

  template <bool kFast, class Epi>
  B200_HD void alpha_segment(int m, u32 (&a)[8], Epi& epi)
  {
    const int S   = (W + L - 1) / L;
    const int top = W - m * L; // this segment covers steps [top-L, top) (clipped at 0)
    const int p0  = top - L;
    const int c   = 2 * kWarmChunks + S + (S - 1 - m);
    begin_chunk(c);
    // ---- recompute beta[p+1] for the segment from the checkpoint beta[top]
    u32 bs[L][8];
    u32 st[8];
    src.ck_get(c, m, st);
#pragma unroll
    for (int s = 0; s < 8; s++)
      bs[L - 1][s] = st[s];
    if (m > 0)
      P::normalize((uint32_t)top, st); // the recursion continued from the normalised beta[top]
#pragma unroll
    for (int i = L - 2; i >= 0; i--) {
      const int k = p0 + i + 1;
      if (kFast || k >= 1) {
        u32 x, y;
        row(c, i + 1, k, x, y);
        bwd_step<P>(st, x, y, P::add(x, y));
#pragma unroll
        for (int s = 0; s < 8; s++)
          bs[i][s] = st[s];
        if (kFast) {
          if (P::kNormPeriod == 1 || ((i + 1) & 1) == 0)
            P::normalize_now(st);
        } else {
          P::normalize((uint32_t)k, st);
        }
      }
    }
    // ---- forward recursion + LLR
#pragma unroll
    for (int i = 0; i < L; i++) {
      const int p = p0 + i;
      if (kFast || p >= 0) {
        u32 llr, x, y;
        row(c, i, p, x, y);
        if (!kFast && P::kMonitor && p < kExactHead) {
          RangeMon unused;
          llr = fwd_step_llr<Sat16>(a, bs[i], x, y, P::add(x, y), unused);
          Sat16::normalize((uint32_t)p, a);
          if (p == kExactHead - 1)
            mon_a.track(a); // state entering the monitored region
        } else {
          llr = fwd_step_llr<P>(a, bs[i], x, y, P::add(x, y), mon_a);
          if (kFast) {
            if (P::kMonitor && (i & 1) == 0)
              mon_a.track(a);
            if (P::kNormPeriod == 1 || (i & 1) == 0)
              P::normalize_now(a);
          } else {
            if (P::kMonitor && (p & 1) == 0)
              mon_a.track(a);
            P::normalize((uint32_t)p, a);
          }
        }
        epi(p, llr, x, src.get_apr(c, i, p), src.get_lut(c, i, p));
      }
    }
  }

  template <class Epi>
  B200_HD void alpha_main(u32 (&a)[8], Epi& epi)
  {
    const int S = (W + L - 1) / L;
    for (int m = S - 1; m >= 0; m--) {
      if (interior(W - (m + 1) * L, P::kMonitor ? kExactHead : 0))
        alpha_segment<true>(m, a, epi);
      else
        alpha_segment<false>(m, a, epi);
    }
  }
};

// lane exchange helpers on packed pairs: lane d <- lane d+1 (beta) / lane d <- lane d-1 (alpha)
// own = (lane 2j | lane 2j+1 << 16); next = word of thread j+1; prev = word of thread j-1
B200_HD u32 shift_down_lanes(u32 own, u32 next) { return (own >> 16) | (next << 16); }
B200_HD u32 shift_up_lanes(u32 prev, u32 own) { return (prev >> 16) | (own << 16); }

} // namespace b200
