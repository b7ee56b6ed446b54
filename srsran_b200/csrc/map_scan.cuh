// k_scan_mat + k_scan_out -- the int16 windowed max-log-MAP of ONE half-iteration, parallel in TIME, for batches of a subframe
// or two that leave the GPU empty (per-TTI latency).
//
// The reference's recursions (include/srslte/phy/fec/turbodecoder_win.h:551-681 backward, 684-832 forward) are serial over
// the W = K / N trellis steps of a sub-block; a lone warp needs ~120 cycles per step (map_lat.cuh), so one half-iteration
// costs 26 us of recursion whatever the machine.  But as long as no saturating operation of the reference saturates -- the
// regime the Fast16 range monitor certifies; everything else is replayed by the exact kernel -- the recursion is
// LINEAR in the (max, +) algebra: the normalisation (subtract metric 0 every second step) only shifts all eight metrics of a
// step by a common amount, and state_{k+1} = A_k (x) state_k with an 8 x 8 (max, +) matrix A_k made of the step's branch
// metrics.  (max, +) products are associative, so the sub-block is cut into chunks of 96 steps:
//
//   k_scan_mat   one thread per (lane, chunk, basis state j): runs the UN-normalised recursion of its chunk from the unit
//                vector e_j in int32 (no overflow: |metric| <= 424 steps x 2 g) and stores the 8-vector it has reached at
//                the sample points the output kernel needs -- i.e. column j of the transfer matrix from the chunk boundary
//                to that step -- and at the chunk's far end.  Forward and backward recursions, all chunks, all lanes and
//                all basis states run at the same time: 96 serial steps instead of 424.  Beside them one warp per group
//                runs each 40-step warm-up pass (packed int16, the reference's schedule) and the lane hand-over.  The CTA
//                of a group that finishes last chains the chunk-end matrices: exact state at every chunk boundary.
//   k_scan_out   one warp per (group, 8-step tile): exact state at its sample point = matrix (x) boundary state; the
//                reference's int16 state there is that vector minus its element 0 (sample points are steps right after a
//                normalisation); from there the tile runs the reference's own packed int16 arithmetic (map_f16.cuh /
//                map_core.cuh: same step functions, normalisation points, range-monitor tracking points) for its <= 11
//                backward and 9 forward steps, the a-posteriori LLRs and the half-iteration glue
//                (turbodecoder_iter.h:104-128) exactly as k_map_lat's second phase does.  The warp of a group that
//                finishes last folds the range-monitor records of all tiles and warm-ups and gives the verdict.
//
// Same integers as k_map_lat / k_map_f16 whenever the verdict is "sound" (then neither the reference saturates nor anything
// here wraps, and exact arithmetic is what both compute); a flagged block is replayed by the exact Sat16 kernel launched
// right after, as with the other Fast16 kernels.
//
// Two ways to run it: k_scan_mat + k_scan_out, one pair of launches per half-iteration (transport blocks: a CRC decision
// follows every half-iteration), or k_scan_fused, ONE cooperative launch for all half-iterations of a run_all batch: the
// CTAs of a group meet at a barrier between the phases (a counter in global memory; the cooperative launch guarantees that
// every CTA is resident), the descriptors are read once, and a block the monitor flags is parked for the exact kernel
// (k_map_fused<Sat16>, mode 2) that runs after it.
#pragma once
#include "map_f16.cuh"

namespace b200 {

// one int32 per word: the step functions of map_core.cuh instantiate with it unchanged
struct Exact32 {
  B200_HD static u32 add(u32 a, u32 b) { return (u32)((int32_t)a + (int32_t)b); }
  B200_HD static u32 max(u32 a, u32 b) { return (u32)((int32_t)a > (int32_t)b ? (int32_t)a : (int32_t)b); }
  B200_HD static u32 addmax(u32 a, u32 b, u32 c) { return max(add(a, b), c); }
  B200_HD static u32 addmax2(u32 a, u32 b, u32 c, u32 d) { return max(add(a, b), add(c, d)); }
};

constexpr int     kScanChunk     = 96;         // steps per chunk (12 tiles): at most 4 chunks, 3 chain steps per direction
constexpr int32_t kScanNegInf    = -(1 << 28); // "unreachable" in the unit vectors: never wins a maximum after three steps
constexpr int     kScanMinW      = 192;        // shorter sub-blocks: k_map_lat's serial recursion is as fast (measured)
constexpr int     kScanMaxW      = 384;        // longest sub-block of the AUTO dispatch (K = 6144 over 16 lanes); manual decoder types can exceed it
constexpr int     kScanMaxGroups = 8;          // the scratch is sized for this many groups (two subframes of 13 blocks)
constexpr int     kScanThreads   = 256;        // CTA size of k_scan_mat / k_scan_fused
constexpr int     kScanTileWarps = 4;          // warps of a k_scan_fused CTA that take output tiles (staging memory for each)
constexpr int     kScanAcc       = 18;         // words per thread of a group's range-monitor accumulator

// scratch of one group, in int32 words: per tile and direction a transfer matrix of each of the group's 64 lanes, stored
// [state][half of the basis][lane][4 basis states] (the output kernel's 128-bit loads of a warp are then contiguous); per
// chunk and direction the chunk-end matrix [lane][state][basis] and the exact boundary state [lane][state]
struct ScanLay {
  int n_tiles, n_chunks;
  B200_HD size_t mat(int t, int dir) const { return ((size_t)t * 2 + dir) * 4096; }
  B200_HD static size_t mat_elem(int l, int s, int k) { return ((size_t)(s * 2 + (k >> 2)) * 64 + l) * 4 + (k & 3); }
  B200_HD size_t end(int c, int dir) const { return (size_t)n_tiles * 8192 + ((size_t)c * 2 + dir) * 4096; }
  B200_HD size_t vec(int c, int dir) const { return (size_t)n_tiles * 8192 + (size_t)n_chunks * 8192 + ((size_t)c * 2 + dir) * 512; }
  B200_HD size_t words() const { return (size_t)n_tiles * 8192 + (size_t)n_chunks * 8192 + (size_t)n_chunks * 1024; }
};

#if defined(__CUDACC__)

struct ScanArgs {
  MapArgs   m;        // work list, descriptors, workspace (as for k_map_lat); m.iter = half-iteration index within the batch
  int32_t*  scratch;  // n_groups x group_words
  size_t    group_words;
  ScanLay   lay;      // for the largest W of the class
  uint32_t* arrive;   // per group 4 counters, monotonic over the batch: k_scan_mat CTAs / k_scan_out warps that have finished;
                      // k_scan_fused: arrivals and releases of its two barriers
  int32_t*  acc;      // [group][kScanAcc][32]: range-monitor accumulator of the half-iteration in flight; all zero between launches
  int       launch_no; // how many times the kernel pair ran before in this batch
  // k_scan_fused only
  int       n_half;   // half-iterations to run (every live block of the class starts at the same count and has no CRC)
  int*      parked;   // groups holding parked blocks, for k_map_fused<Sat16> mode 2
  uint32_t* n_parked; // its length (FusedArgs::counters[2])
};

// what a thread knows about the code block it works for: block lane / T of the group in the packed kernels
struct ScanSlot {
  int      cb;
  bool     live;
  int      W, K, qoff;
  uint64_t d_ws;
  size_t   ps;
};
// warp-collective; false when no block of the group is being decoded.  skip_redo: blocks the monitor parked are not live
template <int N>
__device__ __forceinline__ bool scan_load_slot(const MapArgs& a, int grp, int lane, bool skip_redo, ScanSlot& g, bool same_count = true)
{
  constexpr int T = N / 2, G = 32 / T;
  const int     slot = grp * G + lane / T;
  g.cb   = slot < a.n_slots ? a.work[slot] : -1;
  g.live = g.cb >= 0;
  uint32_t d_W = 0, d_K = 0, d_ps = 0, d_qpp = 0, n_iter0 = 0;
  g.d_ws = 0;
  if (g.live) {
    const CbDev*   dp = a.cbs + g.cb;
    const CbState* sp = a.state + g.cb;
    d_W     = dp->W;
    d_K     = dp->K;
    d_ps    = dp->ps;
    d_qpp   = dp->qpp_off;
    g.d_ws  = dp->ws_off;
    n_iter0 = sp->n_iter;
    if (sp->done || n_iter0 >= dp->max_iter || (skip_redo && sp->redo))
      g.live = false;
  }
  const unsigned live_mask = __ballot_sync(0xffffffffu, g.live);
  if (live_mask == 0)
    return false;
  const int leader = __ffs(live_mask) - 1;
  g.W    = __shfl_sync(0xffffffffu, (int)d_W, leader);
  g.K    = __shfl_sync(0xffffffffu, (int)d_K, leader);
  g.qoff = __shfl_sync(0xffffffffu, (int)d_qpp, leader);
  const int niter = __shfl_sync(0xffffffffu, (int)n_iter0, leader);
  g.ps   = d_ps;
  if (g.live && ((int)d_W != g.W || (same_count && (int)n_iter0 != niter)))
    __trap();
  return true;
}

// Range-monitor contribution of one warp (a tile or a warm-up pass) to its group's accumulator: packed int16x2 maxima /
// minima go in as two int32 each (fire-and-forget reductions in L2).  Word w of thread `lane` at acc[w * 32 + lane]:
// 0-3 alpha max (lo, hi half), min; 4-7 head; 8-11 beta; 12 LLR-subtraction overflow bits; 13-16 extrinsic max, min; 17 bad start
__device__ __forceinline__ void scan_accumulate(int32_t* acc, int lane, const RangeMon* mon_a, const RangeMon* mon_h, const RangeMon* mon_b, u32 ovf, u32 ehi,
                                                u32 elo, u32 bad)
{
  int32_t* q = acc + lane;
  auto mx = [&](int w, u32 v) {
    if (v) {
      atomicMax(q + 32 * w, (int32_t)lo16(v));
      atomicMax(q + 32 * (w + 1), (int32_t)hi16(v));
    }
  };
  auto mn = [&](int w, u32 v) {
    if (v) {
      atomicMin(q + 32 * w, (int32_t)lo16(v));
      atomicMin(q + 32 * (w + 1), (int32_t)hi16(v));
    }
  };
  if (mon_a) { mx(0, mon_a->hi); mn(2, mon_a->lo); }
  if (mon_h) { mx(4, mon_h->hi); mn(6, mon_h->lo); }
  if (mon_b) { mx(8, mon_b->hi); mn(10, mon_b->lo); }
  if (ovf)
    atomicOr(reinterpret_cast<unsigned*>(q + 32 * 12), ovf);
  mx(13, ehi);
  mn(15, elo);
  if (bad)
    atomicOr(reinterpret_cast<unsigned*>(q + 32 * 17), bad);
}

// ---------------------------------------------------------------------------------------------------- transfer matrices
// Rows of one chunk for 32 lanes of a group, staged in shared memory by the whole CTA with ONE round trip to L2 (every load
// independent): sm[plane][step][32 lanes] int16.  A global round trip costs about a microsecond here; the recursion that
// follows must not pay one every eight steps.
struct ScanStageDesc { // what thread tid knows about the lane pair 32 h + 2 (tid % 16) it stages
  bool       live;
  const u32* pin; // word of the pair in row 0 of the first input plane (decoder 1's)
  size_t     psw; // plane stride in words
};
constexpr int kScanStageWords = 3 * kScanChunk * 16;
template <int N>
__device__ __forceinline__ void scan_stage_desc(const MapArgs& a, int grp, int tid, ScanStageDesc (&sd)[2])
{
  constexpr int G = 64 / N;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int l    = 32 * h + 2 * (tid & 15);
    const int slot = grp * G + l / N;
    const int cb   = slot < a.n_slots ? a.work[slot] : -1;
    sd[h].live = cb >= 0;
    sd[h].pin  = nullptr;
    sd[h].psw  = 0;
    if (cb >= 0) {
      const CbDev* dp = a.cbs + cb;
      sd[h].psw = dp->ps / 2;
      sd[h].pin = reinterpret_cast<const u32*>(a.ws + dp->ws_off + (size_t)kPlSyst * dp->ps) + (l % N) / 2;
    }
  }
}
template <int N>
__device__ __forceinline__ void scan_stage_rows(u32* sm, const ScanStageDesc& d, int tid, int p_lo, int p_hi, int mode)
{
  const int    n_planes = mode == 1 ? 3 : 2;
  const size_t shift    = mode == 2 ? (size_t)(kPlApp2 - kPlSyst) * d.psw : 0;
  const int    w        = tid & 15;
  // (plane, step) pairs round robin over the 16 thread groups of the CTA; every load is issued before the first store
  constexpr int kPer = 3 * kScanChunk * 16 / kScanThreads;
  const int     n    = p_hi - p_lo, total = n_planes * n;
  u32           v[kPer];
#pragma unroll
  for (int k = 0; k < kPer; k++) {
    const int q  = (tid >> 4) + (kScanThreads / 16) * k;
    const int pl = q >= 2 * n ? 2 : (q >= n ? 1 : 0), i = q - pl * n;
    v[k]         = (d.live && q < total) ? d.pin[shift + (size_t)pl * d.psw + (size_t)(p_lo + i) * (N / 2)] : 0u;
  }
#pragma unroll
  for (int k = 0; k < kPer; k++) {
    const int q  = (tid >> 4) + (kScanThreads / 16) * k;
    const int pl = q >= 2 * n ? 2 : (q >= n ? 1 : 0), i = q - pl * n;
    if (q < total)
      sm[(pl * kScanChunk + i) * 16 + w] = v[k];
  }
}

// One thread: column jb of the transfer matrices of lane l (lane lq = l % 32 of the staged rows) over chunk c in direction dir.
template <int N>
__device__ __forceinline__ void scan_mat_columns(const ScanLay& lay, int32_t* sc, const u32* sm, int dir, int c, int l, int jb, int W, int mode)
{
  const bool kApr = mode == 1;
  const int  p_lo = c * kScanChunk, p_hi = min(p_lo + kScanChunk, W);
  if (p_lo >= W)
    return;
  u32 e[8];
#pragma unroll
  for (int s = 0; s < 8; s++)
    e[s] = (u32)(s == jb ? 0 : kScanNegInf);
  const int16_t* rows = reinterpret_cast<const int16_t*>(sm) + (l & 31);
  auto rowx = [&](int p, int32_t& x, int32_t& y) {
    const int16_t* r = rows + (p - p_lo) * 32;
    x = r[0];
    y = r[kScanChunk * 32];
    if (kApr)
      x += r[2 * kScanChunk * 32];
  };
  auto put_tile = [&](int32_t* dst) { // the output kernel's layout
#pragma unroll
    for (int s = 0; s < 8; s++)
      dst[ScanLay::mat_elem(l, s, jb)] = (int32_t)e[s];
  };
  auto put_end = [&](int32_t* dst) { // [l][state][basis]
#pragma unroll
    for (int s = 0; s < 8; s++)
      dst[(size_t)l * 64 + s * 8 + jb] = (int32_t)e[s];
  };
  // eight rows at a time into registers, the next eight fetched before these are consumed (chunks start at multiples of 8)
  int32_t xa[8], ya[8], xb[8], yb[8];
  if (!dir) {
    // forward: after step p the vector is the state before step p + 1; sample points 8 t - 1 (tile t >= 1)
    auto load = [&](int p0, int32_t (&xs)[8], int32_t (&ys)[8]) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (p0 + i < p_hi)
          rowx(p0 + i, xs[i], ys[i]);
    };
    load(p_lo, xa, ya);
#pragma unroll 1
    for (int p0 = p_lo; p0 < p_hi; p0 += 16) {
      load(p0 + 8, xb, yb);
      // steps p0 .. p0 + 6, the sample (state before step p0 + 7, tile (p0 >> 3) + 1), step p0 + 7
#pragma unroll
      for (int i = 0; i < 7; i++)
        if (p0 + i < p_hi)
          fwd_step<Exact32>(e, (u32)xa[i], (u32)ya[i], (u32)(xa[i] + ya[i]));
      if (p0 + 7 < W && (p0 >> 3) + 1 < lay.n_tiles && p0 + 7 <= p_hi)
        put_tile(sc + lay.mat((p0 >> 3) + 1, 0));
      if (p0 + 7 < p_hi)
        fwd_step<Exact32>(e, (u32)xa[7], (u32)ya[7], (u32)(xa[7] + ya[7]));
      if (p0 + 8 < p_hi) {
        load(p0 + 16, xa, ya);
#pragma unroll
        for (int i = 0; i < 7; i++)
          if (p0 + 8 + i < p_hi)
            fwd_step<Exact32>(e, (u32)xb[i], (u32)yb[i], (u32)(xb[i] + yb[i]));
        if (p0 + 15 < W && (p0 >> 3) + 2 < lay.n_tiles && p0 + 15 <= p_hi)
          put_tile(sc + lay.mat((p0 >> 3) + 2, 0));
        if (p0 + 15 < p_hi)
          fwd_step<Exact32>(e, (u32)xb[7], (u32)yb[7], (u32)(xb[7] + yb[7]));
      }
    }
    if (p_hi < W)
      put_end(sc + lay.end(c, 0));
  } else {
    // backward: after step p the vector is beta_p; sample points p = 8 t + 10 (p % 8 == 2) for tile t >= 0.
    // Batches of eight steps p0 + 7 .. p0 (p0 a multiple of 8): the sample follows step p0 + 2
    auto load = [&](int p0, int32_t (&xs)[8], int32_t (&ys)[8]) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (p0 + i < p_hi && p0 >= p_lo)
          rowx(p0 + i, xs[i], ys[i]);
    };
    auto run = [&](int p0, const int32_t (&xs)[8], const int32_t (&ys)[8]) {
#pragma unroll
      for (int i = 7; i >= 0; i--)
        if (p0 + i < p_hi) {
          bwd_step<Exact32>(e, (u32)xs[i], (u32)ys[i], (u32)(xs[i] + ys[i]));
          if (i == 2 && p0 + 2 >= 10)
            put_tile(sc + lay.mat((p0 + 2 - 10) >> 3, 1));
        }
    };
    const int p_top = (p_hi - 1) & ~7;
    load(p_top, xa, ya);
#pragma unroll 1
    for (int p0 = p_top; p0 >= p_lo; p0 -= 16) {
      load(p0 - 8, xb, yb);
      run(p0, xa, ya);
      if (p0 - 8 >= p_lo) {
        load(p0 - 16, xa, ya);
        run(p0 - 8, xb, yb);
      }
    }
    if (p_lo > 0)
      put_end(sc + lay.end(c, 1));
  }
}

// One warp: a 40-step warm-up pass in the reference's packed int16 arithmetic (as k_map_lat's first phase), the lane
// hand-over, the boundary state of the chain and the pass's range-monitor contribution.  mode: 0 = decoder 1 without
// a-priori input, 1 = decoder 1, 2 = decoder 2 (a runtime value and rolled loops: this code runs once per launch, what it
// costs is the instruction fetch)
template <int N>
__device__ __noinline__ void scan_warmup(const MapArgs& a, const ScanLay& lay, int32_t* sc, int32_t* acc, bool fwd, const ScanSlot& g, int lane, int mode)
{
  constexpr int  T = N / 2;
  const bool     kDec2 = mode == 2, kApr = mode == 1;
  const int      plane0 = kDec2 ? kPlApp2 : kPlSyst;
  using P = Fast16;
  const int       j     = lane % T;
  const unsigned  gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane / T * T));
  const int       W = g.W;
  const u32*      pin = reinterpret_cast<const u32*>(a.ws + g.d_ws + (size_t)plane0 * g.ps) + j;
  const size_t    psw = g.ps / 2;
  const int16_t*  tl  = a.tails + (size_t)(g.live ? g.cb : 0) * 12;
  const int p_first = fwd ? W - kWinOverlap : 0;
  // rows of 8 steps at a time, the next eight in flight while these are consumed; backward: row index 39 - i
  u32 xa[8], ya[8], xb[8], yb[8];
  auto load = [&](int i0, u32 (&xs)[8], u32 (&ys)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int  k  = i0 + i;
      const bool ok = g.live && k < kWinOverlap;
      const u32* r  = pin + (size_t)(p_first + (fwd ? k : kWinOverlap - 1 - k)) * T;
      const u32  vin = ok ? r[0] : 0u;
      ys[i] = ok ? r[psw] : 0u;
      xs[i] = (kApr && ok) ? P::add(r[2 * psw], vin) : vin;
    }
  };
  RangeMon mon, mon_h;
  mon.reset();
  mon_h.reset();
  u32 st[8];
#pragma unroll
  for (int s = 0; s < 8; s++)
    st[s] = splat16(-P::kInf);
  auto run = [&](int i0, const u32 (&xs)[8], const u32 (&ys)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = i0 + i; // loop counter of the pass
      if (fwd) {
        // forward warm-up: steps W-40..W-1 (win.h:747-756); normalisation and tracking follow the loop counter
        fwd_step<P>(st, xs[i], ys[i], P::add(xs[i], ys[i]));
        if ((i & 1) == 0 && k > 2)
          mon.track(st);
        if ((i & 1) == 0 && k != 0)
          P::normalize_now(st);
      } else {
        // backward warm-up: steps 39..0 of the lane's own sub-block (win.h:622-630); p = 39 - k has the parity of i + 1
        const int p = kWinOverlap - 1 - k;
        bwd_step<P>(st, xs[i], ys[i], P::add(xs[i], ys[i]));
        if ((i & 1) == 1 && p < 38)
          mon.track(st);
        if ((i & 1) == 1 && p != 0)
          P::normalize_now(st);
      }
    }
  };
  load(0, xa, ya);
#pragma unroll 1
  for (int i0 = 0; i0 < kWinOverlap; i0 += 16) {
    load(i0 + 8, xb, yb);
    run(i0, xa, ya);
    if (i0 + 8 < kWinOverlap) {
      load(i0 + 16, xa, ya);
      run(i0 + 8, xb, yb);
    }
  }
  if (!fwd) {
    // the lane below takes the estimate; tail trellis for the last lane (win.h:580-612, 500-548)
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
    mon.track(st);
  } else {
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
    mon_h.track(st);
  }
  // boundary state of the chain: forward chunk 0 / the top backward chunk, one int32 per lane and state
  int32_t* v = sc + lay.vec(fwd ? 0 : (W - 1) / kScanChunk, fwd ? 0 : 1) + (size_t)((lane / T) * N + 2 * j) * 8;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    v[s]     = lo16(st[s]);
    v[8 + s] = hi16(st[s]);
  }
  if (fwd)
    scan_accumulate(acc, lane, &mon, &mon_h, nullptr, 0, 0, 0, 0);
  else
    scan_accumulate(acc, lane, nullptr, nullptr, &mon, 0, 0, 0, 0);
}

// One thread (state s of lanes lA and lB of the group; the eight threads of a lane are consecutive lanes of a warp and the warp
// takes one path per lane): exact state at every chunk boundary from the chunk-end matrices.  Called after every matrix and
// both warm-up passes of the group are in global memory (written by other CTAs: L2 loads, all of them issued before the first use).
__device__ __forceinline__ void scan_chain2(const ScanLay& lay, int32_t* sc, int lA, bool liveA, int lB, bool liveB, int s, int lane, int W)
{
  constexpr int kMaxC = (kScanMaxW + kScanChunk - 1) / kScanChunk;
  const int     nc    = (W + kScanChunk - 1) / kScanChunk;
  int32_t mf[2][kMaxC - 1][8], mb[2][kMaxC - 1][8], vf[2], vb[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int  l  = h ? lB : lA;
    const bool lv = h ? liveB : liveA;
#pragma unroll
    for (int c = 0; c < kMaxC - 1; c++)
      if (lv && c + 1 < nc) {
        const int4* pf = reinterpret_cast<const int4*>(sc + lay.end(c, 0) + (size_t)l * 64 + s * 8);
        const int4* pb = reinterpret_cast<const int4*>(sc + lay.end(c + 1, 1) + (size_t)l * 64 + s * 8);
        const int4  f0 = __ldcg(pf), f1 = __ldcg(pf + 1), b0 = __ldcg(pb), b1 = __ldcg(pb + 1);
        mf[h][c][0] = f0.x; mf[h][c][1] = f0.y; mf[h][c][2] = f0.z; mf[h][c][3] = f0.w; mf[h][c][4] = f1.x; mf[h][c][5] = f1.y; mf[h][c][6] = f1.z; mf[h][c][7] = f1.w;
        mb[h][c][0] = b0.x; mb[h][c][1] = b0.y; mb[h][c][2] = b0.z; mb[h][c][3] = b0.w; mb[h][c][4] = b1.x; mb[h][c][5] = b1.y; mb[h][c][6] = b1.z; mb[h][c][7] = b1.w;
      }
    vf[h] = lv ? __ldcg(sc + lay.vec(0, 0) + (size_t)l * 8 + s) : 0;
    vb[h] = lv ? __ldcg(sc + lay.vec(nc - 1, 1) + (size_t)l * 8 + s) : 0;
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int  l  = h ? lB : lA;
    const bool lv = h ? liveB : liveA;
    if (!lv)
      continue;
    // forward: state before step 96 (c + 1) = end matrix of chunk c (x) state before step 96 c
#pragma unroll
    for (int c = 0; c < kMaxC - 1; c++)
      if (c + 1 < nc) {
        int32_t nv = INT32_MIN;
#pragma unroll
        for (int k = 0; k < 8; k++)
          nv = max(nv, mf[h][c][k] + __shfl_sync(0xffffffffu, vf[h], (lane & ~7) + k));
        vf[h] = nv;
        sc[lay.vec(c + 1, 0) + (size_t)l * 8 + s] = nv;
      }
    // backward: beta at step 96 c = end matrix of chunk c (x) beta at the chunk's top boundary
#pragma unroll
    for (int c = kMaxC - 1; c >= 1; c--)
      if (c < nc) {
        int32_t nv = INT32_MIN;
#pragma unroll
        for (int k = 0; k < 8; k++)
          nv = max(nv, mb[h][c - 1][k] + __shfl_sync(0xffffffffu, vb[h], (lane & ~7) + k));
        vb[h] = nv;
        sc[lay.vec(c - 1, 1) + (size_t)l * 8 + s] = nv;
      }
  }
}

// ------------------------------------------------------------------------------------------------------------ output tiles
// Staging of a tile's transfer matrices and boundary states: cp.async (16 bytes per thread and copy, no registers held) into
// the warp's shared memory, so that EVERYTHING a tile reads from L2 travels in one round trip (a dependent global round trip
// costs about a microsecond when the GPU is this empty).  Slot k of thread `lane` at stg[k * 32 + lane] (conflict-free).
// slots: 0-31 forward matrix (lane h: 16 h + 2 s + half), 32-63 backward matrix, 64-67 forward state (2 h + half), 68-71 backward
constexpr int kScanTileSlots = 72;
__device__ __forceinline__ void scan_cp16(int4* dst, const void* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void scan_fetch_state(int4* stg, int lane, int slot0, int mslot0, const int32_t* mat, const int32_t* vec, int l0)
{
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int4* vp = reinterpret_cast<const int4*>(vec + (size_t)(l0 + h) * 8);
    scan_cp16(stg + (slot0 + 2 * h) * 32 + lane, vp);
    scan_cp16(stg + (slot0 + 2 * h + 1) * 32 + lane, vp + 1);
    if (mat) {
      const int4* mp = reinterpret_cast<const int4*>(mat) + (l0 + h); // [state][half of the basis][lane]
#pragma unroll
      for (int q = 0; q < 16; q++)
        scan_cp16(stg + (mslot0 + 16 * h + q) * 32 + lane, mp + q * 64);
    }
  }
}
// state at a sample point = transfer matrix (x) exact boundary state, minus element 0 (the reference's normalisation), for
// the thread's two lanes; sets bad when a value leaves int16 (the reference would have saturated before).  has_mat false:
// the boundary state as it is (never normalised)
__device__ __forceinline__ void scan_fold_state(const int4* stg, int lane, int slot0, int mslot0, bool has_mat, u32 (&st)[8], u32& bad)
{
  int32_t r[2][8];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int4 va = stg[(slot0 + 2 * h) * 32 + lane], vb = stg[(slot0 + 2 * h + 1) * 32 + lane];
    const int32_t v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
    if (has_mat) {
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const int4 ma = stg[(mslot0 + 16 * h + 2 * s) * 32 + lane], mb = stg[(mslot0 + 16 * h + 2 * s + 1) * 32 + lane];
        int32_t e = max(max(ma.x + v[0], ma.y + v[1]), max(ma.z + v[2], ma.w + v[3]));
        e         = max(e, max(max(mb.x + v[4], mb.y + v[5]), max(mb.z + v[6], mb.w + v[7])));
        r[h][s]   = e;
      }
#pragma unroll
      for (int s = 7; s >= 0; s--) {
        r[h][s] -= r[h][0];
        if (r[h][s] > 32767 || r[h][s] < -32768)
          bad = 1;
      }
    } else {
#pragma unroll
      for (int s = 0; s < 8; s++)
        r[h][s] = v[s];
    }
  }
#pragma unroll
  for (int s = 0; s < 8; s++)
    st[s] = ((u32)r[0][s] & 0xffffu) | ((u32)r[1][s] << 16);
}
// the forward start state of tile t only depends on where the warp sits (not on the descriptors): its copies can start first
__device__ __forceinline__ void scan_fetch_alpha(const ScanLay& lay, const int32_t* sc, int4* stg, int lane, int t, int l0)
{
  scan_fetch_state(stg, lane, 64, 0, t == 0 ? nullptr : sc + lay.mat(t, 0), sc + lay.vec(t == 0 ? 0 : (8 * t - 1) / kScanChunk, 0), l0);
}

// One warp: tile t (steps 8t .. 8t + 7) of the group.  al: forward state at the sample point (scan_alpha_start).
// Thread = two adjacent sub-block lanes of one code block (int16x2), as in every other MAP kernel of the library; rows come
// straight from global memory (eleven rows, all loads independent).
// stg: the warp's staging memory, scan_fetch_alpha() already issued
template <int N>
__device__ __forceinline__ void scan_tile(const MapArgs& a, const ScanLay& lay, int32_t* sc, int32_t* acc, int t, const ScanSlot& g, int lane, int4* stg,
                                          bool write_post, int mode)
{
  constexpr int  T = N / 2;
  const bool     kDec2 = mode == 2, kApr = mode == 1;
  const int      plane0 = kDec2 ? kPlApp2 : kPlSyst;
  using P = Fast16;
  const int      j    = lane % T;
  const int      l0   = (lane / T) * N + 2 * j;
  const int      W    = g.W;
  const bool     live = g.live;
  int16_t*       ws   = a.ws + g.d_ws;
  const size_t   ps   = g.ps;
  const int      r1   = (8 * t + 8) <= W ? 8 : W - 8 * t; // steps of this tile
  const int      m_b  = 8 * t + 10 < W ? 8 * t + 10 : W;  // the backward recursion of this tile starts from beta at m_b
  const bool     from_top = m_b == W;
  // ---- everything the tile reads, requested at once: the backward matrix and state (copies), rows 8t-1 .. 8t+9 and QPP rows
  scan_fetch_state(stg, lane, 68, 32, from_top ? nullptr : sc + lay.mat(t, 1), sc + lay.vec((from_top ? W - 1 : m_b) / kScanChunk, 1), l0);
  const u32*   pin = reinterpret_cast<const u32*>(ws + (size_t)plane0 * ps) + j;
  const size_t psw = ps / 2;
  u32 xr[11], yr[11], ar[8], qr[8];
#pragma unroll
  for (int i = 0; i < 11; i++) {
    const int  p  = 8 * t - 1 + i;
    const bool ok = live && p >= 0 && p < W;
    const u32* r  = pin + (size_t)p * T;
    const u32  vin = ok ? r[0] : 0u;
    yr[i] = ok ? r[psw] : 0u;
    const u32 ap = (kApr && ok) ? r[2 * psw] : 0u; // (zero without a-priori input: x = input, e = llr - 0)
    if (i >= 1 && i <= 8)
      ar[i - 1] = ap;
    xr[i] = P::add(ap, vin);
  }
  {
    const uint16_t* q   = a.qpp + g.qoff;
    const u32*      lut = (const u32*)(kDec2 ? q : q + g.K) + (size_t)8 * t * T + j;
#pragma unroll
    for (int i = 0; i < 8; i++)
      qr[i] = i < r1 ? __ldg(lut + i * T) : 0u;
  }
  // ---- start states
  u32 al[8], bt[8], bad_start = 0;
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncwarp();
  scan_fold_state(stg, lane, 64, 0, t != 0, al, bad_start);
  scan_fold_state(stg, lane, 68, 32, !from_top, bt, bad_start);
  __syncwarp(); // (the staging memory may be refilled for the warp's next tile)

  RangeMon mon_a, mon_h, mon_b, mon_l;
  mon_a.reset();
  mon_h.reset();
  mon_b.reset();
  mon_l.reset();
  // ---- backward: beta_p for p = m_b - 1 .. 8t + 1 (tile 0: .. 0); beta_{8t+1} .. beta_{8t+r1} are what the LLRs consume.
  //      Tracking / normalisation points of the serial pass (even p; never p = 0); every even p is tracked by exactly one
  //      tile: p in [8t+1, 8t+8] (and p = 0 by tile 0; the start state by the top tile)
  u32 bs[8][8]; // bs[i] = beta_{8t+1+i} before normalisation
  const bool top_tile = from_top && 8 * t + r1 == W;
  if (top_tile)
    mon_b.track(bt);
  const int p_stop = t == 0 ? 0 : 8 * t + 1;
#pragma unroll
  for (int k = 9; k >= 0; k--) { // p = 8t + k: static indices, the guards are uniform
    const int p = 8 * t + k;
    if (k >= 1 && k <= 8 && top_tile && p == W) {
#pragma unroll
      for (int s = 0; s < 8; s++)
        bs[k - 1][s] = bt[s]; // beta[W]: consumed as it is
    }
    if (p < m_b && p >= p_stop) {
      const u32 x = xr[k + 1], y = yr[k + 1];
      bwd_step<P>(bt, x, y, P::add(x, y));
      if (k >= 1 && k <= 8) {
#pragma unroll
        for (int s = 0; s < 8; s++)
          bs[k - 1][s] = bt[s];
      }
      if ((k & 1) == 0 && k <= 8)
        mon_b.track(bt);
      if ((k & 1) == 0 && p != 0)
        P::normalize_now(bt);
    }
  }
  // ---- forward: the step before the tile (8t - 1, odd: neither tracked nor normalised), then the tile
  if (t > 0)
    fwd_step<P>(al, xr[0], yr[0], P::add(xr[0], yr[0]));
  else
    mon_h.track(al); // the start state belongs to the head monitor
  u32* const     post   = (u32*)(ws + kPlPost * ps);
  int16_t* const post16 = ws + kPlPost * ps;
  int16_t* const ext    = kDec2 ? ws + kPlApr * ps : ws + kPlApp2 * ps;
  u32            ehi = 0, elo = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i < r1) {
      const int p = 8 * t + i;
      const u32 x = xr[i + 1], y = yr[i + 1];
      const u32 xy  = P::add(x, y);
      const u32 llr = llr_factored<P>(al, bs[i], x, y, xy, mon_l);
      if (i < 7) { // (the state after the tile's last step is the next tile's business)
        fwd_step<P>(al, x, y, xy);
        if ((i & 1) == 0) {
          if (t == 0 && i < 4)
            mon_h.track(al);
          else
            mon_a.track(al);
        }
        if ((i & 1) == 0 && p != 0)
          P::normalize_now(al);
      }
      const uint32_t t0 = qr[i] & 0xffffu, t1 = qr[i] >> 16;
      u32 e;
      if (!kDec2) {
        e = P::glue_sub(llr, ar[i], false, false);
        if (live && write_post)
          post[p * T + j] = llr;
        if (live) {
          ext[t0] = (int16_t)lo16(e);
          ext[t1] = (int16_t)hi16(e);
        }
      } else {
        e = P::glue_sub(llr, x, false, false);
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
  // (tile 0's first tracked values and the start state belong to the head monitor)
  scan_accumulate(acc, lane, &mon_a, t == 0 ? &mon_h : nullptr, &mon_b, mon_l.ovf, ehi, elo, bad_start);
}

// One warp, once every contribution of the group is in: the verdict of the half-iteration (k_map_lat's epilogue).  Returns
// true for the threads of flagged blocks.  The accumulator is left zeroed for the next half-iteration.
template <int N>
__device__ __noinline__ bool scan_verdict(const MapArgs& a, int32_t* acc, const ScanSlot& g, int lane, int mode)
{
  constexpr int  T = N / 2;
  const bool     kDec2 = mode == 2, kApr = mode == 1;
  const int      j = lane % T;
  const unsigned gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane / T * T));
  int32_t* q = acc + lane;
  int32_t  w[kScanAcc];
#pragma unroll
  for (int i = 0; i < kScanAcc; i++)
    w[i] = __ldcg(q + 32 * i);
#pragma unroll
  for (int i = 0; i < kScanAcc; i++)
    q[32 * i] = 0;
  auto pk = [](int32_t lo, int32_t hi) -> u32 { return ((u32)lo & 0xffffu) | ((u32)hi << 16); };
  RangeMon mon_a, mon_h, mon_b;
  mon_a.hi = pk(w[0], w[1]);  mon_a.lo = pk(w[2], w[3]);
  mon_h.hi = pk(w[4], w[5]);  mon_h.lo = pk(w[6], w[7]);
  mon_b.hi = pk(w[8], w[9]);  mon_b.lo = pk(w[10], w[11]);
  const u32 ovf = (u32)w[12], ehi = pk(w[13], w[14]), elo = pk(w[15], w[16]), bad_start = (u32)w[17];
  const bool live = g.live;
  int*       gm   = a.gmax + (size_t)(live ? g.cb : 0) * 4;
  const int  gg   = kDec2 ? gm[3] + gm[2] : (kApr ? gm[3] : 0) + gm[0] + gm[1];
  int ge = max(max(lo16(ehi), hi16(ehi)), max(-lo16(elo), -hi16(elo)));
#pragma unroll
  for (int o = T / 2; o >= 1; o >>= 1)
    ge = max(ge, __shfl_xor_sync(gmask, ge, o, T));
  const bool bad = !fast16_beta_ok(mon_b.spread_lo(), gg) || !fast16_beta_ok(mon_b.spread_hi(), gg) ||
                   !fast16_alpha_ok(mon_a.spread_lo(), mon_b.spread_lo(), gg) || !fast16_alpha_ok(mon_a.spread_hi(), mon_b.spread_hi(), gg) ||
                   !fast16_alpha_ok(mon_h.spread_lo(), mon_b.spread_lo(), gg) || !fast16_alpha_ok(mon_h.spread_hi(), mon_b.spread_hi(), gg) ||
                   (ovf & 0x80008000u) != 0 || bad_start != 0;
  const bool flagged = __any_sync(gmask, bad && live) && live;
  if (flagged) {
    if (j == 0)
      a.state[g.cb].redo = 1;
  } else if (j == 0 && live) {
    gm[3] = ge;
  }
  return flagged;
}

// ------------------------------------------------------------------------------------------------------------ k_scan_mat
// grid: n_groups x (4 n_chunks + 2) CTAs of 256 threads (a subframe's four groups then cover the 148 SMs).  Item i of a
// group: i < 2 n_chunks forward chunk i / 2, lanes 32 (i % 2) .. +31; i < 4 n_chunks the same for the backward chunks; then
// the forward warm-up and the backward warm-up (one warp each).
// Thread of a matrix item: lane l = 32 (i % 2) + tid / 8 of the group's 64 lanes, basis state tid % 8.
struct ScanLaneDesc { // what a thread of a matrix item / of the chain knows about lane 32 h + tid / 8
  bool           live;
  int            W, cb;
  const int16_t* pin;
  size_t         ps;
};
template <int N>
__device__ __forceinline__ void scan_lane_desc(const MapArgs& a, int grp, int tid, bool skip_redo, ScanLaneDesc (&ld)[2], int mode)
{
  constexpr int G = 64 / N;
  const int plane0 = mode == 2 ? kPlApp2 : kPlSyst;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int l    = 32 * h + (tid >> 3);
    const int slot = grp * G + l / N;
    const int cb   = slot < a.n_slots ? a.work[slot] : -1;
    ld[h].live = cb >= 0;
    ld[h].cb   = cb;
    ld[h].W    = 0;
    ld[h].ps   = 0;
    ld[h].pin  = nullptr;
    if (cb >= 0) {
      const CbDev*   dp = a.cbs + cb;
      const CbState* sp = a.state + cb;
      if (sp->done || sp->n_iter >= dp->max_iter || (skip_redo && sp->redo))
        ld[h].live = false;
      ld[h].W   = (int)dp->W;
      ld[h].ps  = dp->ps;
      ld[h].pin = a.ws + dp->ws_off + (size_t)plane0 * dp->ps + l % N;
    }
  }
}
// item of a group: matrix columns or a warm-up pass
// (every thread of the CTA calls it: the matrix items stage their rows through shared memory with a CTA barrier)
template <int N>
__device__ __forceinline__ void scan_item(const ScanArgs& sa, int grp, int item, int tid, const ScanLaneDesc (&ld)[2], const ScanStageDesc (&sd)[2], u32* sm,
                                          const ScanSlot& g, bool slot_ok, int mode, int W)
{
  int32_t* const sc = sa.scratch + (size_t)grp * sa.group_words;
  if (item < 4 * sa.lay.n_chunks) {
    const int dir = item >= 2 * sa.lay.n_chunks;
    const int c   = (dir ? item - 2 * sa.lay.n_chunks : item) >> 1;
    const int h   = item & 1;
    const ScanLaneDesc& d = h ? ld[1] : ld[0];
    const int p_lo = c * kScanChunk, p_hi = min(p_lo + kScanChunk, W);
    __syncthreads(); // (the previous item's rows are no longer read)
    if (p_lo < p_hi)
      scan_stage_rows<N>(sm, h ? sd[1] : sd[0], tid, p_lo, p_hi, mode);
    __syncthreads();
    if (d.live)
      scan_mat_columns<N>(sa.lay, sc, sm, dir, c, 32 * h + (tid >> 3), tid & 7, d.W, mode);
  } else if ((tid >> 5) == 0 && slot_ok) {
    scan_warmup<N>(sa.m, sa.lay, sc, sa.acc + (size_t)grp * kScanAcc * 32, item == 4 * sa.lay.n_chunks, g, tid & 31, mode);
  }
}

template <int N, int MODE>
__global__ void __launch_bounds__(kScanThreads, 1) k_scan_mat(const ScanArgs sa)
{
  const MapArgs& a = sa.m;
  __shared__ int s_last;
  __shared__ u32 s_rows[kScanStageWords];
  const int items = 4 * sa.lay.n_chunks + 2;
  const int grp   = blockIdx.x / items;
  const int item  = blockIdx.x - grp * items;
  const int tid   = threadIdx.x, lane = tid & 31;
  int32_t* const sc = sa.scratch + (size_t)grp * sa.group_words;

  const bool group_live = !(a.iter > 0 && a.counters[4 + a.iter - 1] == 0); // else nothing is left to decode in this batch
  ScanLaneDesc  ld[2];
  ScanStageDesc sd[2];
  ScanSlot      g;
  bool          slot_ok = false;
  if (group_live) {
    // (descriptors up front: three dependent global loads that would otherwise sit on the critical path of the chain)
    scan_lane_desc<N>(a, grp, tid, false, ld, MODE);
    scan_stage_desc<N>(a, grp, tid, sd);
    const int W = (int)a.cbs[a.work[grp * (64 / N)]].W; // (the first slot of a group is never padding; one size per group)
    if (item >= 4 * sa.lay.n_chunks && tid < 32)
      slot_ok = scan_load_slot<N>(a, grp, lane, false, g);
    scan_item<N>(sa, grp, item, tid, ld, sd, s_rows, g, slot_ok, MODE, W);
  }
  // ---- the CTA of the group that arrives last chains the chunk-end matrices
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const uint32_t old = atomicAdd(&sa.arrive[4 * grp], 1u);
    s_last             = old == (uint32_t)(sa.launch_no + 1) * (uint32_t)items - 1u;
  }
  __syncthreads();
  if (!s_last || !group_live)
    return;
  __threadfence();
  // (the eight threads of a lane -- and the four lanes of a warp, which belong to one code block -- agree on liveness)
  scan_chain2(sa.lay, sc, tid >> 3, ld[0].live, 32 + (tid >> 3), ld[1].live, tid & 7, lane, ld[0].live ? ld[0].W : ld[1].W);
}

// ------------------------------------------------------------------------------------------------------------ k_scan_out
// grid: n_groups x n_tiles warps (one per CTA: the tiles of a subframe spread over the whole GPU)
template <int N, int MODE>
__global__ void __launch_bounds__(32, 1) k_scan_out(const ScanArgs sa)
{
  constexpr int T = N / 2;
  const MapArgs& a = sa.m;
  const int lane = threadIdx.x;
  const int grp  = blockIdx.x / sa.lay.n_tiles;
  const int t    = blockIdx.x - grp * sa.lay.n_tiles;
  int32_t* const sc  = sa.scratch + (size_t)grp * sa.group_words;
  int32_t* const acc = sa.acc + (size_t)grp * kScanAcc * 32;

  extern __shared__ __align__(16) int4 s_stg[];
  // the copies of the forward start state run while the descriptor loads are in flight
  scan_fetch_alpha(sa.lay, sc, s_stg, lane, t, (lane / T) * N + 2 * (lane % T));

  ScanSlot g;
  bool group_live = !(a.iter > 0 && a.counters[4 + a.iter - 1] == 0);
  if (group_live)
    group_live = scan_load_slot<N>(a, grp, lane, false, g);
  if (group_live && t < ((g.W + 7) >> 3))
    scan_tile<N>(a, sa.lay, sc, acc, t, g, lane, s_stg, (a.mode & kMapSkipPost) == 0, MODE);
  else
    asm volatile("cp.async.wait_all;\n" ::: "memory");

  // ---- the warp of the group that arrives last gives the verdict
  __threadfence();
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    const uint32_t old = atomicAdd(&sa.arrive[4 * grp + 1], 1u);
    last               = old == (uint32_t)(sa.launch_no + 1) * (uint32_t)sa.lay.n_tiles - 1u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last || !group_live)
    return;
  __threadfence();
  scan_verdict<N>(a, acc, g, lane, MODE);
}

// ------------------------------------------------------------------------------------------------------------ k_scan_fused
// Cooperative launch: gridDim.x / n_groups CTAs of 256 threads per group, all half-iterations of a run_all batch.
// Barrier of a group's CTAs: every thread's writes are fenced, one thread per CTA arrives on a monotonic counter; the CTA
// that arrives last does the serial bit between the phases (chain / verdict) and then releases the others.
template <int N>
__global__ void __launch_bounds__(kScanThreads, 1) k_scan_fused(const ScanArgs sa)
{
  constexpr int T = N / 2;
  const MapArgs& a = sa.m;
  __shared__ int s_last;
  const int n_groups = (a.n_slots + 64 / N - 1) / (64 / N);
  const int cpg = gridDim.x / n_groups; // CTAs per group
  const int grp = blockIdx.x / cpg, cig = blockIdx.x - grp * cpg;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  if (grp >= n_groups)
    return;
  int32_t* const  sc  = sa.scratch + (size_t)grp * sa.group_words;
  int32_t* const  acc = sa.acc + (size_t)grp * kScanAcc * 32;
  uint32_t* const bar = sa.arrive + 4 * grp;
  const int       items = 4 * sa.lay.n_chunks + 2;

  // the last arriver runs `serial` (all its threads), then releases the group
  uint32_t epochs[2] = {0, 0};
  auto barrier = [&](int which, auto serial) {
    const uint32_t epoch = ++epochs[which];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const uint32_t old = atomicAdd(bar + 2 * which, 1u);
      s_last             = old == epoch * (uint32_t)cpg - 1u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      serial();
      __threadfence();
      __syncthreads();
      if (tid == 0)
        atomicAdd(bar + 2 * which + 1, 1u);
    } else if (tid == 0) {
      // (every CTA of the grid is resident -- cooperative launch -- so the release always comes; bounded all the same: a
      //  logic error must be a trap and an error code, not a hung GPU)
      for (uint32_t spins = 0; *(volatile uint32_t*)(bar + 2 * which + 1) < epoch; spins++) {
        __nanosleep(32);
        if (spins > (1u << 24))
          __trap();
      }
    }
    __syncthreads();
    __threadfence();
  };

  // descriptors, once; liveness is refreshed every half-iteration (blocks the monitor flags are parked: state.redo)
  __shared__ u32 s_rows[kScanStageWords];
  extern __shared__ __align__(16) int4 s_stg[]; // kScanTileWarps x kScanTileSlots x 32
  ScanLaneDesc  ld0[2], ld[2];
  ScanStageDesc sd[2];
  ScanSlot      g0, g;
  scan_stage_desc<N>(a, grp, tid, sd);

  scan_lane_desc<N>(a, grp, tid, false, ld0, 0);
  if (!scan_load_slot<N>(a, grp, lane, false, g0))
    return; // nothing to decode in this group (uniform over its CTAs)
  const int niter0 = __reduce_max_sync(0xffffffffu, g0.live ? (int)a.state[g0.cb].n_iter : 0);
  for (int it = 0; it < sa.n_half; it++) {
    ld[0] = ld0[0];
    ld[1] = ld0[1];
    g     = g0;
    if (it > 0) {
      if (ld[0].live && a.state[ld[0].cb].redo)
        ld[0].live = false;
      if (ld[1].live && a.state[ld[1].cb].redo)
        ld[1].live = false;
      if (g.live && a.state[g.cb].redo)
        g.live = false;
    }
    const bool slot_ok = __any_sync(0xffffffffu, g.live);
    if (!slot_ok)
      break; // (uniform over the group: every CTA reads the same state, written before the last barrier)
    // MODE of the half-iteration: every live block of the class sits at the same count (run_all semantics)
    const int niter = niter0 + it;
    const int mode  = (niter & 1) ? 2 : (niter ? 1 : 0);
    const int plane_shift = mode == 2 ? kPlApp2 - kPlSyst : 0;
#pragma unroll
    for (int h = 0; h < 2; h++)
      if (ld[h].pin)
        ld[h].pin += (size_t)plane_shift * ld[h].ps;
    const bool write_post = it + 1 == sa.n_half; // the a-posteriori plane is only read by the decision after the last half-iteration

    // ---- phase A: transfer matrices and warm-up passes
    for (int item = cig; item < items; item += cpg)
      scan_item<N>(sa, grp, item, tid, ld, sd, s_rows, g, slot_ok, mode, g0.W);
    barrier(0, [&]() { scan_chain2(sa.lay, sc, tid >> 3, ld[0].live, 32 + (tid >> 3), ld[1].live, tid & 7, lane, g0.W); });
    // ---- phase B: output tiles, one warp each
    const int nT = (g.W + 7) >> 3;
    if (wib < kScanTileWarps)
      for (int t = cig + wib * cpg; t < nT; t += cpg * kScanTileWarps) { // (spread over the CTAs first)
        int4* stg = s_stg + (size_t)wib * kScanTileSlots * 32;
        scan_fetch_alpha(sa.lay, sc, stg, lane, t, (lane / T) * N + 2 * (lane % T));
        scan_tile<N>(a, sa.lay, sc, acc, t, g, lane, stg, write_post, mode);
      }
    barrier(1, [&]() {
      if (wib == 0) {
        const bool f = scan_verdict<N>(a, acc, g, lane, mode);
        // a flagged block is parked at this half-iteration: its inputs are intact, the exact kernel takes over from here
        if (f && lane % T == 0)
          a.state[g.cb].n_iter = (uint32_t)niter;
      }
    });
  }
  // ---- end of the run: the blocks still live have run all their half-iterations but the decision (k_decide_crc takes the
  //      a-posteriori plane from here: it counts the half-iteration it decides after); groups with parked blocks go on the list
  if (cig == 0 && wib == 0) {
    ScanSlot ge;
    // every block that was being decoded when the kernel started, parked or not
    const bool ok = scan_load_slot<N>(a, grp, lane, false, ge, false);
    if (ok) {
      bool has_parked = false;
      if (ge.live && lane % T == 0) {
        CbState* s = a.state + ge.cb;
        if (s->redo)
          has_parked = true;
        else
          s->n_iter = a.cbs[ge.cb].max_iter - 1;
        // half-iterations this kernel ran for the block (MapArgs::counters[1]; the decision kernel counts the last one, the
        // exact kernel its own)
        const uint32_t ran = s->n_iter - (uint32_t)niter0;
        if (ran)
          atomicAdd(const_cast<uint32_t*>(a.counters) + 1, ran);
      }
      if (__any_sync(0xffffffffu, has_parked) && lane == 0) {
        const uint32_t at = atomicAdd(sa.n_parked, 1u);
        sa.parked[at]     = grp;
      }
    }
  }
}

#endif // __CUDACC__

} // namespace b200
