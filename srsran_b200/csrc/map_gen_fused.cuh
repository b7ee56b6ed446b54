// k_gen_fused: the generic decoder (tdec_gen_dec, turbodecoder_gen.c:58-236) with the whole per-code-block loop of
// sch.c:420-450 in ONE launch -- every half-iteration, the hard decision (gen.c:260-276), srslte_crc_checksum_byte
// (crc.c:143-157) and the early stop -- for the short blocks the reference's dispatch sends there (K <= 400 under
// SRSLTE_TDEC_AUTO, turbodecoder.c:381-393).  Same integers as k_map_gen (wrapping int16, full-length recursions,
// normalisation every fourth step), different shape:
//
//   * k_map_gen is one thread per pair of blocks and one launch (plus one decision launch) per half-iteration: a mixed
//     batch brings a few dozen warps, each a serial chain of 2K steps that reads its rows from global memory and writes
//     8 beta words per step to a global scratch -- 0.4 ms per half-iteration whatever the batch size (45 % of the
//     device time of the all-sizes workload c3).
//   * here ONE CTA owns a pair of equal-K blocks (one per int16x2 half) from its first half-iteration to its last, with
//     every plane of the pair in shared memory.  A half-iteration is four phases:
//       0. all threads: the decoder's input rows x = systematic + a-priori (or the interleaved extrinsic), y = parity;
//       1. TWO serial chains side by side: lane 0 of warp 0 runs the backward recursion and stores beta_k, lane 0 of
//          warp 1 runs the forward recursion WITHOUT outputs and stores alpha_{k-1} (the recursions do not depend on each
//          other; only the LLR needs both);
//       2. all threads: LLR_k from (alpha_{k-1}, beta_k, x, y) for every k in parallel, the glue of iter.h:107-127
//          (a-posteriori, extrinsic through the QPP table) as 16-bit stores into the shared planes;
//       3. decisions: one thread per output byte, CRC by lane 0 of the half's warp (K/8 <= 64 bytes), early stop.
//     The chains never wait for memory and a half-iteration of K = 400 is ~400 dependent steps instead of ~800.
//
// The planes (a-priori, interleaved extrinsic, a-posteriori) are written back at the end, so single-block sessions
// (srslte_tdec_iteration) and the debug plane read-back see what k_map_gen would have left.
#pragma once

namespace b200 {

struct GenFusedArgs {
  const int*      work; // pairs: work[2p], work[2p+1] (second may be -1)
  int             n_pairs;
  const CbDev*    cbs;
  CbState*        state;
  int16_t*        ws;
  const int16_t*  tails;
  const uint16_t* qpp;
  uint8_t*        cb_out;
  const uint32_t* crc_tab;
  uint32_t*       counters; // [1]: half-iterations run
  int             n_half;   // half-iterations this launch may run per block
  int             kp;       // row length of the shared arrays (>= K_max + 4)
};
constexpr int kGenFusedThreads = 128; // warps 0 and 1 hold the two chains; all four share the parallel phases
constexpr int kGenFusedMaxK    = 512;
// shared memory: alpha[8][kp], beta[8][kp], x, y, 6 planes (u32 = the pair), the QPP tables fwd | rev (uint16), CRC tables,
// decided bytes of both halves
constexpr int kGenFusedRows = 16 + 2 + 6 + 1;
inline size_t gen_fused_smem(int kp) { return (size_t)kp * kGenFusedRows * 4 + 2 * 256 * 4 + 2 * (kGenFusedMaxK / 8); }

__global__ void __launch_bounds__(kGenFusedThreads) k_gen_fused(const GenFusedArgs a)
{
  extern __shared__ __align__(16) u32 gsm[];
  const int kp    = a.kp;
  u32*      sA    = gsm;          // alpha_{k-1} at [s * kp + k], k = 1..K
  u32*      sB    = sA + 8 * kp;  // beta_k      at [s * kp + k], k = 1..K
  u32*      sX    = sB + 8 * kp;  // k = 0..K+2
  u32*      sY    = sX + kp;
  u32*      sPl   = sY + kp;      // planes in workspace order: kPlSyst, kPlPar0, kPlApr, kPlApp2, kPlPar1, kPlPost
  const uint16_t* sQ = reinterpret_cast<const uint16_t*>(sPl + 6 * kp); // fwd[K] | rev[K] (natural order: one table per K)
  uint32_t(*s_tab)[256] = reinterpret_cast<uint32_t(*)[256]>(sPl + 7 * kp);
  uint8_t* sBytes = reinterpret_cast<uint8_t*>(s_tab + 2); // [2][kGenFusedMaxK / 8]
  __shared__ uint32_t s_res[2][2];                          // per half: crc, stop flag

  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int cb0 = a.work[2 * blockIdx.x], cb1r = a.work[2 * blockIdx.x + 1];
  if (cb0 < 0)
    return;
  const int cb[2] = {cb0, cb1r < 0 ? cb0 : cb1r};
  const uint32_t K = a.cbs[cb0].K;
  int16_t*       w[2];
  uint32_t       ps[2], maxit[2], poly[2], n_iter[2], crc_last[2] = {1, 1};
  bool           act[2], act_start[2], decided[2] = {false, false}, passed[2] = {false, false};
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const CbDev*  d = &a.cbs[cb[h]];
    const CbState s = a.state[cb[h]];
    w[h]         = a.ws + d->ws_off;
    ps[h]        = d->ps;
    maxit[h]     = d->max_iter;
    poly[h]      = d->crc_poly;
    n_iter[h]    = s.n_iter;
    act[h]       = (h == 0 || cb1r >= 0) && !(s.done || s.n_iter >= d->max_iter);
    act_start[h] = act[h];
  }
  if (!act[0] && !act[1])
    return;
  // the pair's planes: element k of a shared row holds block 0 in the low half, block 1 in the high half; the three tail
  // values of each constituent code sit behind the planes they continue (systematic / parity 0, interleaved / parity 1)
  {
    const bool fresh = n_iter[0] == 0 && n_iter[1] == 0; // nothing but the inputs has been written yet
    const bool al32  = ((a.cbs[cb[0]].ws_off | a.cbs[cb[1]].ws_off | ps[0] | ps[1]) & 1u) == 0;
#pragma unroll
    for (int pl = 0; pl < 6; pl++) {
      if (fresh && pl != kPlSyst && pl != kPlPar0 && pl != kPlPar1)
        continue;
      if (al32) {
        const u32* g0 = reinterpret_cast<const u32*>(w[0] + pl * (size_t)ps[0]);
        const u32* g1 = reinterpret_cast<const u32*>(w[1] + pl * (size_t)ps[1]);
#pragma unroll 2
        for (uint32_t k2 = t; k2 < K / 2; k2 += kGenFusedThreads) {
          const u32 v0 = g0[k2], v1 = g1[k2];
          sPl[pl * kp + 2 * k2]     = __byte_perm(v0, v1, 0x5410);
          sPl[pl * kp + 2 * k2 + 1] = __byte_perm(v0, v1, 0x7632);
        }
      } else {
        for (uint32_t k = t; k < K; k += kGenFusedThreads)
          sPl[pl * kp + k] = pack16(w[0][pl * (size_t)ps[0] + k], w[1][pl * (size_t)ps[1] + k]);
      }
    }
    if (t < 12) {
      const int     c2 = t / 6, r = t % 6; // constituent code, x tails 0..2 then y tails 3..5
      const int     pl = c2 == 0 ? (r < 3 ? kPlSyst : kPlPar0) : (r < 3 ? kPlApp2 : kPlPar1);
      sPl[pl * kp + K + r % 3] = pack16(a.tails[(size_t)cb[0] * 12 + t], a.tails[(size_t)cb[1] * 12 + t]);
    }
    const u32* gq = reinterpret_cast<const u32*>(a.qpp + a.cbs[cb0].qpp_off);
    u32*       sq = sPl + 6 * kp;
    if ((a.cbs[cb0].qpp_off & 1u) == 0) {
      for (uint32_t i = t; i < K; i += kGenFusedThreads)
        sq[i] = gq[i];
    } else {
      const uint16_t* g16 = a.qpp + a.cbs[cb0].qpp_off;
      uint16_t*       s16 = reinterpret_cast<uint16_t*>(sq);
      for (uint32_t i = t; i < 2 * K; i += kGenFusedThreads)
        s16[i] = g16[i];
    }
  }
  if (poly[0] != 0 || poly[1] != 0)
    load_crc_tables(s_tab, a.crc_tab);
  int16_t*        sPl16 = reinterpret_cast<int16_t*>(sPl);
  uint32_t        ran   = 0;
  __syncthreads();

  for (int it = 0; it < a.n_half; it++) {
    const bool dec2[2]  = {(n_iter[0] & 1u) != 0, (n_iter[1] & 1u) != 0};
    const bool first[2] = {n_iter[0] == 0, n_iter[1] == 0};
    // ---- phase 0: input rows of this half-iteration (gen.c:238-258 + the tail bits), per half
    for (uint32_t k = t; k < K + 3; k += kGenFusedThreads) {
      int32_t x[2], y[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int32_t sy = sPl16[2 * ((dec2[h] ? kPlApp2 : kPlSyst) * kp + k) + h];
        const int32_t ap = (dec2[h] || first[h] || k >= K) ? 0 : sPl16[2 * (kPlApr * kp + k) + h];
        x[h]             = sy + ap; // (wraps in pack16, like the int16 addition of the reference)
        y[h]             = sPl16[2 * ((dec2[h] ? kPlPar1 : kPlPar0) * kp + k) + h];
      }
      sX[k] = pack16(x[0], x[1]);
      sY[k] = pack16(y[0], y[1]);
    }
    __syncthreads();

    // ---- phase 1: the two recursions, one lane each, in different warps
    if (t == 0) { // map_gen_beta, gen.c:71-111: k = K+2 .. 0
      u32 o[8];
      o[0] = 0;
#pragma unroll
      for (int i = 1; i < 8; i++)
        o[i] = splat16(-Wrap16::kInf);
#pragma unroll
      for (int u = 0; u < 3; u++) { // the termination steps k = K+2, K+1, K (beta_K is stored, nothing is normalised)
        const uint32_t k = K + 2 - u;
        const u32      x = sX[k], y = sY[k];
        bwd_step<Wrap16>(o, x, y, p_add_wrap(x, y));
        if (u == 2) {
#pragma unroll
          for (int s = 0; s < 8; s++)
            sB[s * kp + k] = o[s];
        }
      }
      u32 xs[4], ys[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        xs[u] = sX[K - 1 - u];
        ys[u] = sY[K - 1 - u];
      }
      for (int k0 = (int)K - 1; k0 >= 3; k0 -= 4) { // k0 = 3 (mod 4): the chunk ends on a normalisation step
        u32 xn[4], yn[4];
        if (k0 >= 7) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            xn[u] = sX[k0 - 4 - u];
            yn[u] = sY[k0 - 4 - u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int k = k0 - u;
          bwd_step<Wrap16>(o, xs[u], ys[u], p_add_wrap(xs[u], ys[u]));
          if (k >= 1) {
#pragma unroll
            for (int s = 0; s < 8; s++)
              sB[s * kp + k] = o[s];
          }
          if (u == 3) {
#pragma unroll
            for (int i = 1; i < 8; i++)
              o[i] = p_sub_wrap(o[i], o[0]);
            o[0] = 0;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xs[u] = xn[u];
          ys[u] = yn[u];
        }
      }
    } else if (t == 32) { // the forward recursion of map_gen_alpha, gen.c:135-197, k = 1 .. K, outputs left to phase 2
      u32 o[8];
      o[0] = 0;
#pragma unroll
      for (int i = 1; i < 8; i++)
        o[i] = splat16(-Wrap16::kInf);
      u32 xs[4], ys[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        xs[u] = sX[u];
        ys[u] = sY[u];
      }
      for (uint32_t k0 = 1; k0 <= K; k0 += 4) { // k0 = 1 (mod 4): the chunk ends on a normalisation step
        u32 xn[4], yn[4];
        if (k0 + 4 <= K) {
#pragma unroll
          for (int u = 0; u < 4; u++) {
            xn[u] = sX[k0 + 3 + u];
            yn[u] = sY[k0 + 3 + u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t k = k0 + u;
#pragma unroll
          for (int s = 0; s < 8; s++)
            sA[s * kp + k] = o[s];
          fwd_step<Wrap16>(o, xs[u], ys[u], p_add_wrap(xs[u], ys[u]));
          if (u == 3) {
#pragma unroll
            for (int i = 1; i < 8; i++)
              o[i] = p_sub_wrap(o[i], o[0]);
            o[0] = 0;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          xs[u] = xn[u];
          ys[u] = yn[u];
        }
      }
    }
    __syncthreads();

    // ---- phase 2: a-posteriori LLRs and the glue, every trellis step in parallel
    for (uint32_t k = 1 + t; k <= K; k += kGenFusedThreads) {
      u32 o[8], b[8];
#pragma unroll
      for (int s = 0; s < 8; s++) {
        o[s] = sA[s * kp + k];
        b[s] = sB[s * kp + k];
      }
      const u32      x = sX[k - 1], y = sY[k - 1];
      RangeMon       nomon;
      const u32      llr = fwd_step_llr<Wrap16>(o, b, x, y, p_add_wrap(x, y), nomon);
      const uint32_t i   = k - 1;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (!act[h])
          continue;
        const int32_t l = h ? hi16(llr) : lo16(llr);
        if (!dec2[h]) { // app2[rev[i]] = ext1[i] - app1[i]
          const int32_t  ap = first[h] ? 0 : sPl16[2 * (kPlApr * kp + i) + h];
          const uint32_t f  = sQ[K + i];
          sPl16[2 * (kPlPost * kp + i) + h] = (int16_t)l;
          sPl16[2 * (kPlApp2 * kp + f) + h] = (int16_t)(l - ap);
        } else { // app1[fwd[i]] = ext2[i] - app2[i] (decoder 2: x is its own input)
          const int32_t  xi = h ? hi16(x) : lo16(x);
          const uint32_t f  = sQ[i];
          sPl16[2 * (kPlPost * kp + f) + h] = (int16_t)l;
          sPl16[2 * (kPlApr * kp + f) + h]  = (int16_t)(l - xi);
        }
      }
    }
    __syncthreads();

    // ---- phase 3: hard decisions (first bit in time is the MSB), CRC, early stop
    bool need[2];
#pragma unroll
    for (int h = 0; h < 2; h++)
      need[h] = act[h] && (poly[h] != 0 || n_iter[h] + 1 >= maxit[h]);
    if ((uint32_t)t < K / 8) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (!need[h])
          continue;
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 8; j++)
          v = 2 * v + (sPl16[2 * (kPlPost * kp + 8 * t + j) + h] > 0 ? 1u : 0u);
        sBytes[h * (kGenFusedMaxK / 8) + t] = (uint8_t)v;
      }
    }
    __syncthreads();
    if (lane == 0 && wid < 2 && (wid ? act[1] : act[0])) {
      const int      h  = wid;
      const uint32_t pg = wid ? poly[1] : poly[0];
      uint32_t       crc = 1;
      if ((wid ? need[1] : need[0]) && pg != 0) {
        const uint32_t* tab = s_tab[pg == kCrc24A ? 0 : 1];
        const uint8_t*  by  = sBytes + h * (kGenFusedMaxK / 8);
        crc                 = 0;
        for (uint32_t bi = 0; bi < K / 8; bi++)
          crc = ((crc << 8) ^ tab[((crc >> 16) & 0xffu) ^ by[bi]]) & 0xffffffu;
      }
      s_res[h][0] = crc;
      s_res[h][1] = (pg != 0 && crc == 0) ? 1u : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (!act[h])
        continue;
      n_iter[h]++;
      ran++;
      decided[h]  = decided[h] || need[h];
      crc_last[h] = s_res[h][0];
      if (s_res[h][1]) // early stop (sch.c:441-450)
        passed[h] = true;
      act[h] = !(passed[h] || n_iter[h] >= maxit[h]);
    }
    if (!act[0] && !act[1])
      break;
    // (s_res is rewritten after the next three barriers at the earliest)
  }

  // ---- results: state, decided bytes, and the planes a later half-iteration (or a debug read-back) starts from
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (!act_start[h])
      continue;
    if (t == 0) {
      CbState* s = &a.state[cb[h]];
      s->n_iter  = n_iter[h];
      s->crc     = crc_last[h];
      if (passed[h]) {
        s->crc_ok = 1;
        s->done   = 1;
      }
    }
    if (decided[h]) {
      uint8_t* out = a.cb_out + a.cbs[cb[h]].out_off;
      for (uint32_t i = t; i < K / 8; i += kGenFusedThreads)
        out[i] = sBytes[h * (kGenFusedMaxK / 8) + i];
    }
    const bool al32 = ((a.cbs[cb[h]].ws_off | ps[h]) & 1u) == 0;
#pragma unroll
    for (int sel = 0; sel < 3; sel++) {
      const int pl = sel == 0 ? kPlApr : sel == 1 ? kPlApp2 : kPlPost;
      if (al32) {
        u32* g = reinterpret_cast<u32*>(w[h] + pl * (size_t)ps[h]);
        for (uint32_t k2 = t; k2 < K / 2; k2 += kGenFusedThreads)
          g[k2] = __byte_perm(sPl[pl * kp + 2 * k2], sPl[pl * kp + 2 * k2 + 1], h ? 0x7632 : 0x5410);
      } else {
        for (uint32_t k = t; k < K; k += kGenFusedThreads)
          w[h][pl * (size_t)ps[h] + k] = sPl16[2 * (pl * kp + k) + h];
      }
    }
  }
  if (t == 0 && ran)
    atomicAdd(&a.counters[1], ran);
}

} // namespace b200
