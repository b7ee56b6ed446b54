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
  const u32 n2 = P::addmax(o[5], y, P::add(o[1], x));
  const u32 n3 = P::addmax(o[5], x, P::add(o[1], y));
  const u32 n4 = P::addmax(o[6], x, P::add(o[2], y));
  const u32 n5 = P::addmax(o[6], y, P::add(o[2], x));
  const u32 n6 = P::addmax(o[3], xy, o[7]);
  const u32 n7 = P::addmax(o[7], xy, o[3]);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}

// forward (alpha) trellis step without output (warm-up), turbodecoder_win.h:769-785, 820-822
template <class P>
B200_HD void fwd_step(u32 (&o)[8], u32 x, u32 y, u32 xy)
{
  const u32 n0 = P::addmax(o[1], xy, o[0]);
  const u32 n1 = P::addmax(o[3], y, P::add(o[2], x));
  const u32 n2 = P::addmax(o[4], y, P::add(o[5], x));
  const u32 n3 = P::addmax(o[6], xy, o[7]);
  const u32 n4 = P::addmax(o[0], xy, o[1]);
  const u32 n5 = P::addmax(o[2], y, P::add(o[3], x));
  const u32 n6 = P::addmax(o[5], y, P::add(o[4], x));
  const u32 n7 = P::addmax(o[7], xy, o[6]);
  o[0] = n0; o[1] = n1; o[2] = n2; o[3] = n3; o[4] = n4; o[5] = n5; o[6] = n6; o[7] = n7;
}

// forward step with a-posteriori LLR against beta_{k+1} (turbodecoder_win.h:769-822)
template <class P>
B200_HD u32 fwd_step_llr(u32 (&o)[8], const u32 (&b)[8], u32 x, u32 y, u32 xy)
{
  // branch candidates: mb = hypothesis 0, nw = hypothesis 1
  const u32 mb0 = o[0], mb1 = P::add(o[3], y), mb2 = P::add(o[4], y), mb3 = o[7];
  const u32 mb4 = o[1], mb5 = P::add(o[2], y), mb6 = P::add(o[5], y), mb7 = o[6];
  const u32 nw0 = P::add(o[1], xy), nw1 = P::add(o[2], x), nw2 = P::add(o[5], x), nw3 = P::add(o[6], xy);
  const u32 nw4 = P::add(o[0], xy), nw5 = P::add(o[3], x), nw6 = P::add(o[4], x), nw7 = P::add(o[7], xy);
  u32 m1 = P::add(b[0], nw0), m0 = P::add(b[0], mb0);
  m1 = P::addmax(b[1], nw1, m1); m0 = P::addmax(b[1], mb1, m0);
  m1 = P::addmax(b[2], nw2, m1); m0 = P::addmax(b[2], mb2, m0);
  m1 = P::addmax(b[3], nw3, m1); m0 = P::addmax(b[3], mb3, m0);
  m1 = P::addmax(b[4], nw4, m1); m0 = P::addmax(b[4], mb4, m0);
  m1 = P::addmax(b[5], nw5, m1); m0 = P::addmax(b[5], mb5, m0);
  m1 = P::addmax(b[6], nw6, m1); m0 = P::addmax(b[6], mb6, m0);
  m1 = P::addmax(b[7], nw7, m1); m0 = P::addmax(b[7], mb7, m0);
  o[0] = P::max(mb0, nw0); o[1] = P::max(mb1, nw1); o[2] = P::max(mb2, nw2); o[3] = P::max(mb3, nw3);
  o[4] = P::max(mb4, nw4); o[5] = P::max(mb5, nw5); o[6] = P::max(mb6, nw6); o[7] = P::max(mb7, nw7);
  return P::out(P::sub(m1, m0));
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

// One thread's view of a MAP call.  Rows of the lane-layout arrays are T = N/2 words; this thread owns word j.
template <class P, int L>
struct MapWin {
  const u32* in;  // systematic (DEC1) or interleaved extrinsic (DEC2)
  const u32* apr; // a-priori, may be nullptr
  const u32* par; // parity
  int        T;   // words per row
  int        W;   // steps per lane (K / N)
  int        j;   // this thread's word in the row
  u32*       ck;  // checkpoint store for this thread: element (slot, state) at ck[(slot*8+state)*cks]
  int        cks;

  B200_HD void load_xy(int p, u32& x, u32& y) const
  {
    x = in[p * T + j];
    y = par[p * T + j];
    if (apr)
      x = P::add(apr[p * T + j], x);
  }

  // beta pass 0: steps 39..0 of the lane's own sub-block from the all-"unknown" state (win.h:622-630)
  B200_HD void beta_warm(u32 (&st)[8]) const
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
      st[i] = splat16(-P::kInf);
    constexpr int C = 8;
    for (int base = kWinOverlap - C; base >= 0; base -= C) {
      u32 xs[C], ys[C];
#pragma unroll
      for (int i = 0; i < C; i++)
        load_xy(base + i, xs[i], ys[i]);
#pragma unroll
      for (int i = C - 1; i >= 0; i--) {
        bwd_step<P>(st, xs[i], ys[i], P::add(xs[i], ys[i]));
        P::normalize((uint32_t)(base + i), st);
      }
    }
  }

  // beta pass 1: st = initial state at step W (neighbour estimate or tail); stores one checkpoint per segment
  B200_HD void beta_main(u32 (&st)[8]) const
  {
    const int S = (W + L - 1) / L;
#pragma unroll
    for (int i = 0; i < 8; i++)
      ck[i * cks] = st[i]; // slot 0 = beta[W]
    for (int m = 0; m < S; m++) {
      const int p0 = W - (m + 1) * L; // may be negative for the last (lowest) segment
      u32       xs[L], ys[L];
#pragma unroll
      for (int i = 0; i < L; i++)
        if (p0 + i >= 0)
          load_xy(p0 + i, xs[i], ys[i]);
#pragma unroll
      for (int i = L - 1; i >= 0; i--) {
        const int k = p0 + i;
        if (k >= 0) {
          bwd_step<P>(st, xs[i], ys[i], P::add(xs[i], ys[i]));
          if (i == 0 && k > 0) {
#pragma unroll
            for (int s = 0; s < 8; s++)
              ck[((m + 1) * 8 + s) * cks] = st[s]; // beta[k] before normalisation (win.h:666-678)
          }
          P::normalize((uint32_t)k, st);
        }
      }
    }
  }

  // alpha pass 0: steps W-40..W-1 of the lane's own sub-block (win.h:747-756)
  B200_HD void alpha_warm(u32 (&st)[8]) const
  {
#pragma unroll
    for (int i = 0; i < 8; i++)
      st[i] = splat16(-P::kInf);
    constexpr int C = 8;
    for (int k0 = 0; k0 < kWinOverlap; k0 += C) {
      u32 xs[C], ys[C];
#pragma unroll
      for (int i = 0; i < C; i++)
        load_xy(W - kWinOverlap + k0 + i, xs[i], ys[i]);
#pragma unroll
      for (int i = 0; i < C; i++) {
        fwd_step<P>(st, xs[i], ys[i], P::add(xs[i], ys[i]));
        P::normalize((uint32_t)(k0 + i), st);
      }
    }
  }

  // alpha pass 1 with output.  epi(p, llr, x) is called once per step in ascending p.
  template <class Epi>
  B200_HD void alpha_main(u32 (&a)[8], Epi& epi) const
  {
    const int S = (W + L - 1) / L;
    for (int m = S - 1; m >= 0; m--) {
      const int top = W - m * L; // this segment covers steps [top-L, top) (clipped at 0)
      const int p0  = top - L;
      u32       xs[L], ys[L];
#pragma unroll
      for (int i = 0; i < L; i++)
        if (p0 + i >= 0)
          load_xy(p0 + i, xs[i], ys[i]);
      // ---- recompute beta[p+1] for the segment from the checkpoint beta[top]
      u32 bs[L][8];
      u32 st[8];
#pragma unroll
      for (int s = 0; s < 8; s++) {
        st[s]        = ck[(m * 8 + s) * cks];
        bs[L - 1][s] = st[s];
      }
      if (m > 0)
        P::normalize((uint32_t)top, st); // the recursion continued from the normalised beta[top]
#pragma unroll
      for (int i = L - 2; i >= 0; i--) {
        const int k = p0 + i + 1;
        if (k >= 1) {
          bwd_step<P>(st, xs[i + 1], ys[i + 1], P::add(xs[i + 1], ys[i + 1]));
#pragma unroll
          for (int s = 0; s < 8; s++)
            bs[i][s] = st[s];
          P::normalize((uint32_t)k, st);
        }
      }
      // ---- forward recursion + LLR
#pragma unroll
      for (int i = 0; i < L; i++) {
        const int p = p0 + i;
        if (p >= 0) {
          const u32 llr = fwd_step_llr<P>(a, bs[i], xs[i], ys[i], P::add(xs[i], ys[i]));
          P::normalize((uint32_t)p, a);
          epi(p, llr, xs[i]);
        }
      }
    }
  }
};

// lane exchange helpers on packed pairs: lane d <- lane d+1 (beta) / lane d <- lane d-1 (alpha)
// own = (lane 2j | lane 2j+1 << 16); next = word of thread j+1; prev = word of thread j-1
B200_HD u32 shift_down_lanes(u32 own, u32 next) { return (own >> 16) | (next << 16); }
B200_HD u32 shift_up_lanes(u32 prev, u32 own) { return (prev >> 16) | (own << 16); }

} // namespace b200
