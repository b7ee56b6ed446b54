// Device kernels of the B200 LTE turbo-decode engine (sm_100a).  Included by engine.cu only.
//
//   k_dematch_prepare  srslte_rm_turbo_rx_lut[_8bit] + input extraction  (rm_turbo.c:397-493, win.h:880-923)
//   k_dematch          srslte_rm_turbo_rx_lut[_8bit] for the host-pointer compatibility symbol
//   k_prepare          extract_input / extract_input_tail_sb  (win.h:880-923, iter.h:59-69, gen.c:238-258)
//   k_map_f16          tdec_win*_dec + half-iteration glue, int16 (Fast16) and int8 (Sat8)  -> map_f16.cuh
//   k_map_lat          the same integers as k_map_f16 for single-subframe batches: beta and alpha passes on two warps at once,
//                      LLRs + glue spread over four  -> map_lat.cuh
//   k_map_win          the same in exact saturating int16 (replay of blocks the Fast16 range monitor flags; fast16 off)
//   k_map_gen          tdec_gen_dec + glue                    (turbodecoder_gen.c:58-236)
//   k_decide_crc       tdec_*_decision_byte + srslte_crc_checksum_byte + early stop (win.h:925-993, crc.c:143-157,
//                      sch.c:420-450)
//   k_tb_finish        TB assembly, CRC24A, HARQ bookkeeping  (sch.c:462-486, 546-552)
//   k_demod_descramble srslte_demod_soft_demodulate_{s,b} + srslte_scrambling_{s,sb}_offset (demod_soft.c:896-945)
//   k_ulsch_deinterleave  ulsch_deinterleave + ACK/RI/CQI LLR extraction of srslte_ulsch_decode (sch.c:992-1019, 1021-1180)
//   k_enc_tb_crc, k_enc_cb  encode_tb_off: CRC, turbo code, rate matching (sch.c:235-349, turbocoder.c, rm_turbo.c:349-395)
#pragma once
#include <cuda.h> // CUtensorMap (type only: the encoder is reached through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include "map_core.cuh"
#include "ulsch_core.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------ descriptors
struct CbDev {
  uint32_t K;        // code block size
  uint16_t W;        // steps per lane (K / N); 0 for the generic decoder
  uint8_t  N;        // sub-block lanes (0 = generic decoder)
  uint8_t  bits;     // arithmetic of the decoder: 16 or 8
  uint32_t ps;       // plane stride in int16 elements
  uint32_t qpp_off;  // offset (uint16 elements) of fwd[K] | rev[K] in the QPP pool
  uint64_t ws_off;   // offset (int16 elements) of plane 0 in the workspace
  uint32_t max_iter; // half-iteration limit
  uint32_t crc_poly; // early-stop CRC polynomial (24 bit), 0 = no CRC (run_all semantics)
  uint32_t out_off;  // byte offset of this CB's K/8 decided bytes in the CB output pool
  uint32_t sat_end;  // 8-bit glue: elements below this index use the saturating subtract
  uint32_t crc_xp[5]; // CRC chunk-combination constants for K/8 bytes under crc_poly (warp_crc24)
  uint32_t crc_xq[4]; // the same for chunks of ceil(K/8 / (N/2)) bytes: the block's own lanes compute the CRC (group_crc24)
  // input description for k_prepare / k_dematch
  void*       in_ptr; // decoder input: HARQ soft buffer of this CB, or the caller's LLRs (device memory)
  const void* e_ptr;  // rate-matched e-bits of this CB (device memory), dematch only
  uint8_t  in_bits;  // 16 or 8
  uint8_t  in_sb;    // 1: lane layout with K+32 plane stride; 0: standard 3k+s
  uint8_t  dematch;  // 1: run k_dematch into in_ptr first
  uint8_t  skip;     // CB already decoded in a previous HARQ transmission
  uint32_t E;        // number of e-bits of this CB
  uint32_t rm_off;   // offset of the base table (uint16[3K+12]) in the rm pool
  uint32_t rm_start; // rank of the first transmitted entry for this rv
  uint32_t tb;       // owning transport block
  uint32_t cb_in_tb;
  uint32_t fresh;    // soft buffer must be cleared before combining (new transmission): 1 = one-shot, nothing kept;
                     // 2 = caller-owned HARQ buffer that was reset since its last use (cleared, combined, written back)
};

struct CbState {
  uint32_t n_iter;
  uint32_t done;   // early stop hit (CRC passed)
  uint32_t crc_ok;
  uint32_t crc;
  uint32_t redo;   // Fast16 range monitor could not rule out a saturation: replay this half-iteration exactly
  uint32_t n_redo; // statistics: half-iterations replayed
};

struct TbDev {
  uint32_t tbs;
  uint32_t C;
  uint32_t first_cb; // index of CB 0 of this TB in the CB arrays
  uint32_t rlen_bytes[2]; // payload bytes per CB for (K1, K2) type CBs
  uint32_t C1;
  uint8_t* data;       // TB output bytes (device memory), >= tbs/8 + 6
  uint8_t* hdata;      // HARQ saved-CB area (C x 768 bytes) or nullptr
  uint8_t* hcrc;       // HARQ cb_crc flags (C bytes) or nullptr
  uint32_t crc_xp[5];  // CRC24A chunk-combination constants for tbs/8 bytes
};

struct TbResult {
  int32_t  ret;      // 0 ok, -1 CRC error
  uint32_t sum_iter; // total half-iterations over the CBs processed in this call
  uint32_t cb_crc;   // bit i = CB i passed its CRC
  uint32_t par_rx;
};

// Planes of a code block's workspace (plane stride d.ps int16 each).  The order keeps the inputs of each constituent
// decoder adjacent so one TMA box covers them: DEC1 reads {syst, par0, apr}, DEC2 reads {app2, par1}.
constexpr int kPlSyst = 0, kPlPar0 = 1, kPlApr = 2, kPlApp2 = 3, kPlPar1 = 4, kPlPost = 5;
constexpr uint32_t kSbPadDev = 32;
constexpr uint32_t kMaxKDev  = 6144; // plane padding of the lane layout (rm_turbo.c:272)

__constant__ uint32_t c_crc_tab[2][256]; // [0] = CRC24A, [1] = CRC24B byte tables (crc.c:30-46)

constexpr uint32_t kCrc24A = 0x1864CFB;
constexpr uint32_t kCrc24B = 0x1800063;

// ------------------------------------------------------------------------------------------ de-rate-matching
// One CTA per code block.  Gather form of  out[tab[i % L]] += in[i]  (wrapping): thread i < L sums its
// repetitions i, i+L, ... and does a single read-modify-write, so no atomics are needed.
template <typename T>
__global__ void __launch_bounds__(256) k_dematch(const CbDev* __restrict__ cbs, const int* __restrict__ list,
                                                 const uint16_t* __restrict__ rm_pool)
{
  const CbDev d = cbs[list[blockIdx.x]];
  const uint32_t L = 3 * d.K + 12;
  T*             out = (T*)d.in_ptr;
  if (d.fresh) {
    const uint32_t n = d.in_sb ? 3 * (d.K + kSbPadDev) + 12 : L;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
      out[i] = 0;
    __syncthreads();
  }
  const T*        in  = (const T*)d.e_ptr;
  const uint16_t* tab = rm_pool + d.rm_off;
  for (uint32_t i = threadIdx.x; i < L && i < d.E; i += blockDim.x) {
    uint32_t acc = 0;
    for (uint32_t r = i; r < d.E; r += L)
      acc += (uint32_t)(int32_t)in[r];
    uint32_t pos = i + d.rm_start;
    if (pos >= L)
      pos -= L;
    const uint16_t o = tab[pos];
    out[o]           = (T)(uint32_t)((uint32_t)(int32_t)out[o] + acc);
  }
}

// ------------------------------------------------------------------------------------------ input extraction
// One CTA per code block: decoder input -> int16 planes syst | par0 | par1 in the decoder's own layout + 12 tail
// values {syst[3], par0[3], app2[3], par1[3]}.
__global__ void __launch_bounds__(256) k_prepare(const CbDev* __restrict__ cbs, const int* __restrict__ list, int16_t* __restrict__ ws,
                                                 int16_t* __restrict__ tails, CbState* __restrict__ state, int* __restrict__ gmax, int reset_state)
{
  __shared__ int s_g[3];
  if (threadIdx.x < 3)
    s_g[threadIdx.x] = 0;
  __syncthreads();
  int g0 = 0, g1 = 0, g2 = 0;
  const int   cb = list[blockIdx.x];
  const CbDev d  = cbs[cb];
  const int16_t* in16 = (const int16_t*)d.in_ptr;
  const int8_t*  in8  = (const int8_t*)d.in_ptr;
  auto get = [&](uint32_t i) -> int32_t {
    int32_t v = d.in_bits == 16 ? (int32_t)in16[i] : (int32_t)in8[i];
    // 16-bit LLRs into an 8-bit decoder are truncated like convert_16_to_8 (turbodecoder.c:451-456)
    if (d.bits == 8 && d.in_bits == 16)
      v = (int32_t)(int8_t)(uint8_t)v;
    return v;
  };
  int16_t* p0 = ws + d.ws_off;
  const uint32_t K = d.K, N = d.N, W = d.W;
  if (threadIdx.x == 0 && reset_state) {
    state[cb].n_iter = 0;
    state[cb].done   = 0;
    state[cb].crc_ok = 0;
    state[cb].crc    = 0;
    state[cb].redo   = 0;
    state[cb].n_redo = 0;
  }
  auto amax = [](int& g, int32_t v) { v = v < 0 ? -v : v; g = v > g ? v : g; };
  if (d.in_sb) {
    // planes are K+32 apart (rm_turbo.c:263-277); the decoder layout equals the input layout
    for (uint32_t i = threadIdx.x; i < K; i += blockDim.x) {
      const int32_t a = get(i), b = get(K + kSbPadDev + i), c = get(2 * (K + kSbPadDev) + i);
      p0[i]            = (int16_t)a;
      p0[d.ps + i]     = (int16_t)b;
      p0[kPlPar1 * (size_t)d.ps + i] = (int16_t)c;
      amax(g0, a); amax(g1, b); amax(g2, c);
    }
  } else {
    // standard order 3n+s: read the block contiguously into shared memory, then emit the three planes in the
    // decoder's layout with consecutive threads writing consecutive elements (both sides coalesced)
    extern __shared__ __align__(16) int16_t s_in[];
    const uint32_t n_in = 3 * K + 12;
    const bool     vec  = d.in_bits == 16 && d.bits == 16 && (((uintptr_t)in16) & 7u) == 0;
    if (vec && N && (W & 3u) == 0 && (blockDim.x % (N / 2)) == 0) {
      // Fast transposition (the common case: int16 LLRs into an int16 windowed decoder).  The block is staged with one
      // word of padding per decoder lane, so the N/2 threads that build the words (lane 2j, lane 2j+1) of one step read
      // different banks; every thread then emits whole 32-bit words of all three planes with consecutive threads
      // writing consecutive words.
      const uint32_t LW = 3 * W, hw = N / 2, hsh = 31 - __clz(hw), nth = blockDim.x; // LW: int16 per lane region
      {
        const uint2* src = reinterpret_cast<const uint2*>(in16);
        u32*         dst = reinterpret_cast<u32*>(s_in);
        for (uint32_t i = threadIdx.x; i < 3 * K / 4; i += nth) {
          const uint2    v  = __ldg(src + i);
          const uint32_t e  = 4 * i, dl = e / LW; // 8-byte pieces never straddle two lane regions (LW % 4 == 0)
          dst[e / 2 + dl]     = v.x;
          dst[e / 2 + dl + 1] = v.y;
        }
      }
      __syncthreads();
      const uint32_t jj = threadIdx.x & (hw - 1);
      const int16_t* la = s_in + (2 * jj) * (LW + 2);
      const int16_t* lb = s_in + (2 * jj + 1) * (LW + 2);
      u32*           q0 = reinterpret_cast<u32*>(p0);
      u32*           q1 = reinterpret_cast<u32*>(p0 + kPlPar0 * (size_t)d.ps);
      u32*           q2 = reinterpret_cast<u32*>(p0 + kPlPar1 * (size_t)d.ps);
      u32            h0 = 0, l0 = 0, h1 = 0, l1 = 0, h2 = 0, l2 = 0;
      for (uint32_t p = threadIdx.x >> hsh; p < W; p += nth >> hsh) {
        const uint32_t w  = p * hw + jj;
        const u32      v0 = (u32)(uint16_t)la[3 * p] | ((u32)(uint16_t)lb[3 * p] << 16);
        const u32      v1 = (u32)(uint16_t)la[3 * p + 1] | ((u32)(uint16_t)lb[3 * p + 1] << 16);
        const u32      v2 = (u32)(uint16_t)la[3 * p + 2] | ((u32)(uint16_t)lb[3 * p + 2] << 16);
        q0[w] = v0;
        q1[w] = v1;
        q2[w] = v2;
        h0 = p_max(h0, v0); l0 = p_min(l0, v0);
        h1 = p_max(h1, v1); l1 = p_min(l1, v1);
        h2 = p_max(h2, v2); l2 = p_min(l2, v2);
      }
      g0 = max(max(lo16(h0), hi16(h0)), max(-lo16(l0), -hi16(l0)));
      g1 = max(max(lo16(h1), hi16(h1)), max(-lo16(l1), -hi16(l1)));
      g2 = max(max(lo16(h2), hi16(h2)), max(-lo16(l2), -hi16(l2)));
    } else {
    if (vec) {
      // 8-byte vector loads (n_in is a multiple of 4)
      const uint2* src = reinterpret_cast<const uint2*>(in16);
      uint2*       dst = reinterpret_cast<uint2*>(s_in);
      for (uint32_t i = threadIdx.x; i < n_in / 4; i += blockDim.x)
        dst[i] = __ldg(src + i);
    } else {
      for (uint32_t i = threadIdx.x; i < n_in; i += blockDim.x)
        s_in[i] = (int16_t)get(i);
    }
    __syncthreads();
    if (N) {
      // transpose natural order (n = lane * W + step) into the lane layout (step * N + lane) one plane at a time
      // through a padded shared-memory plane: consecutive threads read consecutive n and write rows that are an odd
      // number of words apart (no bank conflicts either way); the plane then leaves with consecutive 32-bit stores
      int16_t*       s_out = s_in + (n_in + 7) / 8 * 8;
      const uint32_t rs    = N + 2, hw = N / 2; // padded row length (int16), words per row
      const uint32_t hsh = 31 - __clz(hw); // hw is a power of two and divides the block size
      const uint32_t nth = blockDim.x;
      for (int sp = 0; sp < 3; sp++) {
        // natural order -> padded lane layout; (lane, step) of n = idx and both offsets advance without divisions
        {
          uint32_t dl = threadIdx.x / W, p = threadIdx.x - dl * W;
          uint32_t oo = p * rs + dl;
          const int16_t* si = s_in + 3 * threadIdx.x + sp;
          for (uint32_t idx = threadIdx.x; idx < K; idx += nth) {
            s_out[oo] = *si;
            si += 3 * nth;
            p += nth;
            oo += nth * rs;
            while (p >= W) {
              p -= W;
              oo += 1 - W * rs;
            }
          }
        }
        __syncthreads();
        // padded plane -> global, one word (two lanes of one step) per thread and trip; max |LLR| on packed words
        u32*       q  = reinterpret_cast<u32*>(p0 + (sp == 0 ? kPlSyst : sp == 1 ? kPlPar0 : kPlPar1) * (size_t)d.ps);
        const u32* so = reinterpret_cast<const u32*>(s_out) + (threadIdx.x >> hsh) * (rs / 2) + (threadIdx.x & (hw - 1));
        const uint32_t so_step = (nth >> hsh) * (rs / 2);
        u32            vhi = 0, vlo = 0;
        for (uint32_t w = threadIdx.x; w < W * hw; w += nth) {
          const u32 v = *so;
          q[w]        = v;
          vhi         = p_max(vhi, v);
          vlo         = p_min(vlo, v);
          so += so_step;
        }
        const int g = max(max(lo16(vhi), hi16(vhi)), max(-lo16(vlo), -hi16(vlo)));
        __syncthreads();
        if (sp == 0)
          g0 = g;
        else if (sp == 1)
          g1 = g;
        else
          g2 = g;
      }
    } else {
      for (uint32_t j = threadIdx.x; j < K; j += blockDim.x) {
        const int32_t a = s_in[3 * j], b = s_in[3 * j + 1], c = s_in[3 * j + 2];
        p0[j]            = (int16_t)a;
        p0[d.ps + j]     = (int16_t)b;
        p0[kPlPar1 * (size_t)d.ps + j] = (int16_t)c;
        amax(g0, a); amax(g1, b); amax(g2, c);
      }
    }
    } // (general path)
  }
  if (threadIdx.x < 12) { // (the shared-memory copy is not used here: generic inputs may not have filled it)
    const uint32_t tb = d.in_sb ? 3 * (K + kSbPadDev) : 3 * K;
    // tail order on the wire: x_K z_K x_K+1 z_K+1 x_K+2 z_K+2 | x'_K z'_K x'_K+1 z'_K+1 x'_K+2 z'_K+2
    const uint32_t t = threadIdx.x, grp = t / 3, i = t % 3;
    const uint32_t src = (grp == 0) ? 2 * i : (grp == 1) ? 2 * i + 1 : (grp == 2) ? 6 + 2 * i : 6 + 2 * i + 1;
    tails[(size_t)cb * 12 + t] = (int16_t)get(tb + src);
  }
  // per-plane max |LLR| (bounds the branch metrics for the Fast16 range monitor)
  atomicMax(&s_g[0], g0);
  atomicMax(&s_g[1], g1);
  atomicMax(&s_g[2], g2);
  __syncthreads();
  if (threadIdx.x < 3)
    gmax[(size_t)cb * 4 + threadIdx.x] = s_g[threadIdx.x];
  if (threadIdx.x == 3)
    gmax[(size_t)cb * 4 + 3] = 0;
}

// ------------------------------------------------------------------------------------------ grouping by difficulty
// A warp of the persistent MAP kernel holds G code blocks of one size and runs until the slowest of them stops (CRC early stop,
// sch.c:441-450); blocks that stopped ride along as ghosts.  Which blocks share a warp is free -- any blocks of equal size in
// the same decoder class -- so the blocks of a size are ordered by how hard they look before their planes are written:
//   k_cb_stat   one CTA per code block: 1 - (sum |e|)^2 / (E sum e^2) over its e-bits, the normalised spread of the LLR
//               magnitudes (0 for a noiseless block, growing with the noise), as a fixed-point integer;
//   k_regroup   one CTA per size: sorts the block ids of the size's slots by that number and rewrites the work list and the
//               blocks' workspace offsets (slot i of a size sits at ws_base + i * 6 * ps: the tensor maps address blocks by slot).
// Only the grouping changes; every block is decoded exactly as before.
template <typename T>
__global__ void __launch_bounds__(128) k_cb_stat(const CbDev* __restrict__ cbs, const int* __restrict__ list, int32_t* __restrict__ stat)
{
  const int      cb = list[blockIdx.x];
  const CbDev    d  = cbs[cb];
  const T*       in = (const T*)d.e_ptr;
  unsigned long long sa = 0, sq = 0;
  auto               acc = [&](int32_t v) {
    sa += (unsigned)(v < 0 ? -v : v);
    sq += (unsigned)(v * v);
  };
  // 16-byte loads over the aligned body (one 2-byte load per thread and trip was a chain of 144 round trips for K = 6144:
  // 71 us for 1760 blocks), scalar head and tail
  constexpr uint32_t kPer = 16 / sizeof(T);
  const uint32_t     mis  = (uint32_t)((16u - ((uintptr_t)in & 15u)) & 15u) / (uint32_t)sizeof(T);
  const uint32_t     head = mis < d.E ? mis : d.E;
  const uint32_t     nvec = (d.E - head) / kPer;
  for (uint32_t i = threadIdx.x; i < head; i += blockDim.x)
    acc((int32_t)in[i]);
  const uint4* vp = reinterpret_cast<const uint4*>(in + head);
#pragma unroll 4
  for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
    const uint4 w    = __ldg(vp + i);
    const u32   ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (sizeof(T) == 2) {
        acc(lo16(ww[k]));
        acc(hi16(ww[k]));
      } else {
#pragma unroll
        for (int b8 = 0; b8 < 4; b8++)
          acc((int32_t)(int8_t)(uint8_t)(ww[k] >> (8 * b8)));
      }
    }
  }
  for (uint32_t i = head + nvec * kPer + threadIdx.x; i < d.E; i += blockDim.x)
    acc((int32_t)in[i]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    sa += __shfl_down_sync(0xffffffffu, sa, o);
    sq += __shfl_down_sync(0xffffffffu, sq, o);
  }
  __shared__ unsigned long long s_a[4], s_q[4];
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = sa;
    s_q[threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double a = (double)(s_a[0] + s_a[1] + s_a[2] + s_a[3]), q = (double)(s_q[0] + s_q[1] + s_q[2] + s_q[3]);
    const double r = q > 0 ? (a * a) / ((double)d.E * q) : 1.0;
    stat[cb]       = (int32_t)((1.0 - r) * 1.0e8);
  }
}

constexpr int kRegroupMax = 1024; // blocks of one size the sort handles (larger sizes keep their order)
// kg: per size {offset of its slots in the lists, number of blocks, plane stride, workspace base low, high}
__global__ void __launch_bounds__(256) k_regroup(const int* __restrict__ kg, int* __restrict__ lists, CbDev* __restrict__ cbs, const int32_t* __restrict__ stat)
{
  __shared__ int32_t s_key[kRegroupMax];
  __shared__ int     s_id[kRegroupMax];
  const int*     g   = kg + 5 * blockIdx.x;
  const int      off = g[0], n = g[1];
  const uint32_t ps  = (uint32_t)g[2];
  const uint64_t base = (uint64_t)(uint32_t)g[3] | ((uint64_t)(uint32_t)g[4] << 32);
  int            m   = 1;
  while (m < n)
    m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const int id = i < n ? lists[off + i] : 0x7fffffff;
    s_id[i]      = id;
    s_key[i]     = i < n ? stat[id] : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const int32_t ka = s_key[i], kb = s_key[l];
          const int     ia = s_id[i], ib = s_id[l];
          const bool    gt = ka > kb || (ka == kb && ia > ib); // (ties by block id: the order is a function of the inputs only)
          if (gt == up) {
            s_key[i] = kb; s_key[l] = ka;
            s_id[i]  = ib; s_id[l]  = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int id       = s_id[i];
    lists[off + i]     = id;
    cbs[id].ws_off     = base + (uint64_t)i * 6ull * ps;
  }
}

// ------------------------------------------------------------------------------------------ dematch + extraction
// Transport-block path: rate de-matching (HARQ combining) and input extraction in ONE pass per code block.  The
// soft buffer of the block is assembled in shared memory (zero-filled for a new transmission, loaded for a
// retransmission), the e-bits are accumulated into it (same gather form as k_dematch), and the result leaves the
// SM as the decoder's three int16 planes + tails + per-plane max|LLR|; it is written back to the HARQ soft buffer
// only when the caller keeps one (d.fresh != 1 means a caller-owned buffer).
// Shared-memory index of soft-buffer element i.  The sub-block interleaver hands consecutive e-bits to positions 32 trellis
// steps apart, i.e. a power-of-two stride in the lane layout: un-swizzled, the 32 lanes of a warp hit one or two banks
// (10-16-way conflicts on both the load and the store of the accumulation, measured on the tables).  XOR-ing the bank bits
// with higher index bits spreads them (2-2.5-way); element pairs (int16) / quads (int8) stay together in their word.
template <typename T>
__device__ __forceinline__ uint32_t sb_swz(uint32_t i)
{
  constexpr int  kShift = sizeof(T) == 2 ? 1 : 2;
  const uint32_t h = i >> (5 + kShift); // (< 2^10 for every soft buffer: 18 600 elements)
  return i ^ (((h ^ (h >> 5)) & 31u) << kShift);
}
template <typename T>
__global__ void __launch_bounds__(256) k_dematch_prepare(const CbDev* __restrict__ cbs, const int* __restrict__ list,
                                                         const uint16_t* __restrict__ rm_pool, int16_t* __restrict__ ws,
                                                         int16_t* __restrict__ tails, CbState* __restrict__ state, int* __restrict__ gmax)
{
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_g[3];
  T*             sb = reinterpret_cast<T*>(s_raw);
  const int      cb = list[blockIdx.x];
  const CbDev    d  = cbs[cb];
  const uint32_t K = d.K, N = d.N, W = d.W, L = 3 * K + 12;
  const uint32_t n_sb = d.in_sb ? 3 * (K + kSbPadDev) + 12 : L;
  T*             gsb  = (T*)d.in_ptr;
  if (threadIdx.x < 3)
    s_g[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    state[cb].n_iter = 0;
    state[cb].done   = 0;
    state[cb].crc_ok = 0;
    state[cb].crc    = 0;
    state[cb].redo   = 0;
    state[cb].n_redo = 0;
  }
  if (d.fresh) {
    uint4*         z  = reinterpret_cast<uint4*>(s_raw);
    const uint32_t nz = ((n_sb * sizeof(T) + 127) / 128) * 8; // (whole 128-byte lines: the swizzle permutes within them)
    for (uint32_t i = threadIdx.x; i < nz; i += blockDim.x)
      z[i] = make_uint4(0, 0, 0, 0);
  } else {
    for (uint32_t i = threadIdx.x; i < n_sb; i += blockDim.x)
      sb[sb_swz<T>(i)] = gsb[i];
  }
  __syncthreads();
  const T*        in  = (const T*)d.e_ptr;
  const uint16_t* tab = rm_pool + d.rm_off;
  const uint32_t  n_i = d.E < L ? d.E : L;
  // eight independent elements per thread and trip; all their global loads are issued before the first use
  constexpr int kU = 8;
  for (uint32_t i0 = threadIdx.x; i0 < n_i; i0 += kU * blockDim.x) {
    int32_t  v[kU];
    uint16_t o[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) {
      const uint32_t i = i0 + u * blockDim.x;
      v[u]             = i < n_i ? (int32_t)in[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < kU; u++) {
      const uint32_t i   = i0 + u * blockDim.x;
      uint32_t       pos = i + d.rm_start;
      if (pos >= L)
        pos -= L;
      o[u] = i < n_i ? tab[pos] : (uint16_t)0;
    }
#pragma unroll
    for (int u = 0; u < kU; u++) {
      const uint32_t i = i0 + u * blockDim.x;
      if (i < n_i) {
        uint32_t acc = (uint32_t)v[u];
        for (uint32_t r = i + L; r < d.E; r += L) // repetitions (E > 3K+12): same soft bit received again
          acc += (uint32_t)(int32_t)in[r];
        const uint32_t at = sb_swz<T>(o[u]);
        sb[at] = (T)(uint32_t)((uint32_t)(int32_t)sb[at] + acc); // wraps in the width of the soft buffer
      }
    }
  }
  __syncthreads();
  if (d.fresh != 1) {
    for (uint32_t i = threadIdx.x; i < n_sb; i += blockDim.x)
      gsb[i] = sb[sb_swz<T>(i)];
  }
  // decoder planes (int16 containers), tails, max |LLR| per plane
  int16_t* p0 = ws + d.ws_off;
  int      g0 = 0, g1 = 0, g2 = 0;
  auto     amax = [](int& g, int32_t v) { v = v < 0 ? -v : v; g = v > g ? v : g; };
  if (d.in_sb) {
    u32* q0 = reinterpret_cast<u32*>(p0);
    u32* q1 = reinterpret_cast<u32*>(p0 + d.ps);
    u32* q2 = reinterpret_cast<u32*>(p0 + kPlPar1 * (size_t)d.ps);
    if (sizeof(T) == 2) {
      // int16 soft buffer: the planes are already pairs of adjacent lanes, copy them as 32-bit words
      const u32*     sw = reinterpret_cast<const u32*>(sb);
      const uint32_t o1 = K + kSbPadDev, o2 = 2 * (K + kSbPadDev);
      u32            h0 = 0, l0 = 0, h1 = 0, l1 = 0, h2 = 0, l2 = 0;
      for (uint32_t h = threadIdx.x; h < K / 2; h += blockDim.x) {
        const u32 a = sw[sb_swz<T>(2 * h) >> 1], b = sw[sb_swz<T>(o1 + 2 * h) >> 1], c = sw[sb_swz<T>(o2 + 2 * h) >> 1];
        q0[h] = a;
        q1[h] = b;
        q2[h] = c;
        h0 = p_max(h0, a); l0 = p_min(l0, a);
        h1 = p_max(h1, b); l1 = p_min(l1, b);
        h2 = p_max(h2, c); l2 = p_min(l2, c);
      }
      g0 = max(max(lo16(h0), hi16(h0)), max(-lo16(l0), -hi16(l0)));
      g1 = max(max(lo16(h1), hi16(h1)), max(-lo16(l1), -hi16(l1)));
      g2 = max(max(lo16(h2), hi16(h2)), max(-lo16(l2), -hi16(l2)));
    } else {
      for (uint32_t h = threadIdx.x; h < K / 2; h += blockDim.x) {
        const uint32_t j  = 2 * h;
        // (j is even: the two elements of a pair sit side by side under the swizzle as well)
        const uint32_t ja = sb_swz<T>(j), jb = sb_swz<T>(K + kSbPadDev + j), jc = sb_swz<T>(2 * (K + kSbPadDev) + j);
        const int32_t  a0 = sb[ja], a1 = sb[ja + 1], b0 = sb[jb], b1 = sb[jb + 1];
        const int32_t  c0 = sb[jc], c1 = sb[jc + 1];
        q0[h] = pack16(a0, a1);
        q1[h] = pack16(b0, b1);
        q2[h] = pack16(c0, c1);
        amax(g0, a0); amax(g1, b0); amax(g2, c0);
        amax(g0, a1); amax(g1, b1); amax(g2, c1);
      }
    }
  } else {
    for (uint32_t n = threadIdx.x; n < K; n += blockDim.x) {
      const uint32_t j = N ? (n % W) * N + n / W : n;
      const int32_t  a = sb[sb_swz<T>(3 * n)], b = sb[sb_swz<T>(3 * n + 1)], c = sb[sb_swz<T>(3 * n + 2)];
      p0[j]            = (int16_t)a;
      p0[d.ps + j]     = (int16_t)b;
      p0[kPlPar1 * (size_t)d.ps + j] = (int16_t)c;
      amax(g0, a); amax(g1, b); amax(g2, c);
    }
  }
  if (threadIdx.x < 12) {
    const uint32_t tb = d.in_sb ? 3 * (K + kSbPadDev) : 3 * K;
    const uint32_t t = threadIdx.x, grp = t / 3, i = t % 3;
    const uint32_t src = (grp == 0) ? 2 * i : (grp == 1) ? 2 * i + 1 : (grp == 2) ? 6 + 2 * i : 6 + 2 * i + 1;
    tails[(size_t)cb * 12 + t] = (int16_t)sb[sb_swz<T>(tb + src)];
  }
  atomicMax(&s_g[0], g0);
  atomicMax(&s_g[1], g1);
  atomicMax(&s_g[2], g2);
  __syncthreads();
  if (threadIdx.x < 3)
    gmax[(size_t)cb * 4 + threadIdx.x] = s_g[threadIdx.x];
  if (threadIdx.x == 3)
    gmax[(size_t)cb * 4 + 3] = 0;
}

// ------------------------------------------------------------------------------------------ windowed MAP
constexpr int kMapSkipPost = 0x2000;
struct MapArgs {
  const int*      work;    // code block per slot, -1 = padding (slots of one warp share K)
  int             n_slots;
  const CbDev*    cbs;
  CbState*        state;
  int16_t*        ws;
  const int16_t*  tails;
  const uint16_t* qpp;
  int*            gmax; // per CB: max|syst|, max|par0|, max|par1|, max|extrinsic handed to the next half-iteration|
  int             mode; // 0: decode every active CB; 1: Fast16 attempt (flags redo); 2: replay only CBs flagged redo;
                        // | kMapSkipPost: no hard decision follows this launch, the a-posteriori plane is not written
  u32*            ck_scratch; // global beta-checkpoint scratch: ck_slots x (grid threads) x 8 words
  int             ck_slots;
  const uint32_t* counters; // DecideArgs::counters
  int             iter;     // half-iteration index of this launch within the batch
  // tensor-map staging (k_map_t): per warp {index of its K-group's pair of tensor maps, block coordinate of its first slot}
  const int*         winfo;
  const CUtensorMap* tmaps; // per K-group: [2g] box of 3 planes, [2g+1] box of 2 planes
};

// Row source staged through shared memory with cp.async (LDGSTS): chunk c+1 streams in while chunk c is being
// processed, so the global-memory latency is off the dependent-instruction chain.  Each thread stages and reads
// only its own words, so no block barrier is involved: cp.async.wait_group orders a thread's own copies.
//
// The beta checkpoints (8 words per thread and segment) go to a per-thread global scratch -- written with two
// 128-bit stores in the backward pass, prefetched back with the rows of the segment that needs them.  At one
// block per SM the scratch in flight is a few tens of MB, i.e. L2-resident; shared memory then only holds the
// double-buffered staging area (2 x (3L + 8) words per thread), which is what lets more warps share an SM.
//
// staging layout: kStages buffers; per buffer: rows: word (i, a) at (i*4 + a) * NT with a = in, parity, a-priori, QPP
// table; then per thread 8 checkpoint words.  Chunk c uses buffer c % kStages and is requested kAhead chunks early.
template <int L, int NT, int T>
struct StagedSrc {
  static constexpr int kAhead  = 2;
  static constexpr int kStages = 3;
  static constexpr int kRowWords = L * 4 * NT;              // words of row staging per buffer
  static constexpr int kBufWords = kRowWords + 8 * NT;      // + checkpoint staging
  const u32 *in, *apr, *par, *lut;
  int        j;
  u32*       stage;  // smem base (generic pointer), already offset by threadIdx.x for rows
  unsigned   stage_s; // the same as a shared-window address
  u32*       ckg;    // global scratch of this thread: slot s at ckg[s * ck_stride .. +8)
  size_t     ck_stride;
  __device__ __forceinline__ static void cp4(unsigned dst_s, const u32* src)
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst_s), "l"(src) : "memory");
  }
  __device__ __forceinline__ static void cp16(unsigned dst_s, const u32* src)
  {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_s), "l"(src) : "memory");
  }
  __device__ __forceinline__ unsigned row_s(int buf, int i, int a) const { return stage_s + 4u * (unsigned)(buf * kBufWords + (i * 4 + a) * NT); }
  __device__ __forceinline__ const u32* row_p(int buf) const { return stage + buf * kBufWords; }
  // checkpoint words of this thread in buffer buf: after the rows, 8 consecutive words per thread
  __device__ __forceinline__ int ck_word(int buf) const { return buf * kBufWords + kRowWords + 7 * (int)threadIdx.x; }
  __device__ __forceinline__ void prefetch(int c, int p0, int lo, int hi, int ck_slot)
  {
    const int       buf  = c % kStages;
    const ptrdiff_t row0 = (ptrdiff_t)p0 * T + j;
    const u32 *     gi = in + row0, *gp = par + row0, *ga = apr + row0, *gl = lut + row0;
    const bool      full = p0 >= lo && p0 + L <= hi;
    if (full) { // interior chunk: constant offsets only
#pragma unroll
      for (int i = 0; i < L; i++) {
        cp4(row_s(buf, i, 0), gi + i * T);
        cp4(row_s(buf, i, 1), gp + i * T);
      }
      if (apr) {
#pragma unroll
        for (int i = 0; i < L; i++)
          cp4(row_s(buf, i, 2), ga + i * T);
      }
      if (ck_slot >= 0) {
#pragma unroll
        for (int i = 0; i < L; i++)
          cp4(row_s(buf, i, 3), gl + i * T);
      }
    } else {
#pragma unroll
      for (int i = 0; i < L; i++) {
        const int p = p0 + i;
        if (p >= lo && p < hi) {
          cp4(row_s(buf, i, 0), gi + i * T);
          cp4(row_s(buf, i, 1), gp + i * T);
          if (apr)
            cp4(row_s(buf, i, 2), ga + i * T);
          if (ck_slot >= 0)
            cp4(row_s(buf, i, 3), gl + i * T);
        }
      }
    }
    if (ck_slot >= 0) {
      const unsigned d = stage_s + 4u * (unsigned)ck_word(buf);
      const u32*     g = ckg + (size_t)ck_slot * ck_stride;
      cp16(d, g);
      cp16(d + 16, g + 4);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  __device__ __forceinline__ void prefetch_none() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
  // kAhead newer groups were committed after chunk c's: everything older must have landed
  __device__ __forceinline__ void wait(int) { asm volatile("cp.async.wait_group 2;\n" ::: "memory"); }
  __device__ __forceinline__ void drain() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
  __device__ __forceinline__ u32 get_apr(int c, int i, int) const { return apr ? row_p(c % kStages)[(i * 4 + 2) * NT] : 0u; }
  __device__ __forceinline__ u32 get_lut(int c, int i, int) const { return row_p(c % kStages)[(i * 4 + 3) * NT]; }
  __device__ __forceinline__ void get(int c, int i, int, u32& vin, u32& vapr, u32& vpar) const
  {
    const u32* b = row_p(c % kStages);
    vin          = b[(i * 4 + 0) * NT];
    vpar         = b[(i * 4 + 1) * NT];
    vapr         = apr ? b[(i * 4 + 2) * NT] : 0u;
  }
  __device__ __forceinline__ void ck_put(int slot, const u32 (&st)[8])
  {
    uint4* g = reinterpret_cast<uint4*>(ckg + (size_t)slot * ck_stride);
    g[0]     = make_uint4(st[0], st[1], st[2], st[3]);
    g[1]     = make_uint4(st[4], st[5], st[6], st[7]);
  }
  __device__ __forceinline__ void ck_get(int c, int, u32 (&st)[8]) const
  {
    // (stage is offset by threadIdx.x already: ck_word uses 7*tid so that stage + ck_word = base + ... + 8*tid)
    const uint4* d = reinterpret_cast<const uint4*>(stage + ck_word(c % kStages));
    const uint4  a = d[0], b = d[1];
    st[0] = a.x; st[1] = a.y; st[2] = a.z; st[3] = a.w;
    st[4] = b.x; st[5] = b.y; st[6] = b.z; st[7] = b.w;
  }
};

// DEC1 epilogue: a-posteriori -> post (linear), extrinsic -> app2 through the QPP permutation (iter.h:117-121)
template <class P>
struct EpiDec1 {
  u32*       post;
  int16_t*   app2;
  int        T, j;
  uint32_t   sat_end;
  u32        ehi, elo; // running max / min of the extrinsic values emitted
  // apr == 0 when the call had no a-priori input (n_iter == 0): the subtraction is then the identity
  __device__ __forceinline__ void operator()(int p, u32 llr, u32, u32 apr, u32 r)
  {
    const int w = p * T + j;
    post[w]     = llr;
    const u32 e = P::glue_sub(llr, apr, (uint32_t)(2 * w) < sat_end, (uint32_t)(2 * w + 1) < sat_end);
    ehi = p_max(ehi, e);
    elo = p_min(elo, e);
    app2[r & 0xffffu] = (int16_t)lo16(e);
    app2[r >> 16]     = (int16_t)hi16(e);
  }
};

// DEC2 epilogue: a-posteriori -> post[fwd], extrinsic (= a-posteriori - own input) -> a-priori[fwd]
// (iter.h:107-109 fused with :127)
template <class P>
struct EpiDec2 {
  int16_t*   post;
  int16_t*   apr;
  uint32_t   sat_end;
  u32        ehi, elo;
  __device__ __forceinline__ void operator()(int, u32 llr, u32 x, u32, u32 f)
  {
    const uint32_t t0 = f & 0xffffu, t1 = f >> 16;
    const u32      a  = P::glue_sub(llr, x, t0 < sat_end, t1 < sat_end);
    ehi = p_max(ehi, a);
    elo = p_min(elo, a);
    apr[t0]           = (int16_t)lo16(a);
    apr[t1]           = (int16_t)hi16(a);
    post[t0]          = (int16_t)lo16(llr);
    post[t1]          = (int16_t)hi16(llr);
  }
};

template <class P, int N, int L, int NT, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) k_map_win(const MapArgs a)
{
  constexpr int T = N / 2;  // threads per code block
  constexpr int G = 32 / T; // code blocks per warp
  extern __shared__ __align__(16) u32 smem_ck[];

  if (a.iter > 0 && a.counters[4 + a.iter - 1] == 0)
    return; // nothing left to decode in this batch
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * (NT / 32) + (threadIdx.x >> 5)) * G + lane / T;
  const int j    = lane % T;
  const int cb   = slot < a.n_slots ? a.work[slot] : -1;
  if (cb < 0)
    return;
  const CbDev   d  = a.cbs[cb];
  const CbState st0 = a.state[cb];
  if (st0.done || st0.n_iter >= d.max_iter)
    return;
  if (a.mode == 2 && !st0.redo)
    return;
  const unsigned gmask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane / T * T));

  int16_t*       ws   = a.ws + d.ws_off;
  const int16_t* tl   = a.tails + (size_t)cb * 12;
  const bool     dec2 = st0.n_iter & 1u;

  const int S = (d.W + L - 1) / L;
  const uint16_t* q = a.qpp + d.qpp_off;
  MapWin<P, L, StagedSrc<L, NT, T>> m;
  m.W       = d.W;
  m.src.j   = j;
  m.src.stage   = smem_ck + threadIdx.x;
  m.src.stage_s = (unsigned)__cvta_generic_to_shared(smem_ck + threadIdx.x);
  // global checkpoint scratch: slot-major, 8 words per thread, threads of the whole grid contiguous
  m.src.ck_stride = (size_t)gridDim.x * NT * 8;
  m.src.ckg       = a.ck_scratch + ((size_t)blockIdx.x * NT + threadIdx.x) * 8;
  const int16_t *tin, *tpar;
  if (!dec2) {
    m.src.in  = (const u32*)(ws);
    m.src.par = (const u32*)(ws + d.ps);
    m.src.apr = st0.n_iter ? (const u32*)(ws + kPlApr * (size_t)d.ps) : nullptr;
    m.src.lut = (const u32*)(q + d.K); // rev[] as pairs
    tin       = tl;
    tpar      = tl + 3;
  } else {
    m.src.in  = (const u32*)(ws + kPlApp2 * (size_t)d.ps);
    m.src.par = (const u32*)(ws + kPlPar1 * (size_t)d.ps);
    m.src.apr = nullptr;
    m.src.lut = (const u32*)q; // fwd[] as pairs
    tin       = tl + 6;
    tpar      = tl + 9;
  }
  (void)S;
  m.begin();

  u32 st[8];
  // ---- backward: warm-up, hand the estimate to the lane below, tail for the last lane, checkpointed pass
  m.beta_warm(st);
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const u32 nx = __shfl_down_sync(gmask, st[s], 1, T);
    st[s]        = shift_down_lanes(st[s], nx);
  }
  if (j == T - 1) {
    int32_t t[8];
    tail_trellis<P>(tin, tpar, t);
#pragma unroll
    for (int s = 0; s < 8; s++)
      st[s] = (st[s] & 0xffffu) | ((u32)(uint16_t)t[s] << 16);
  }
  m.beta_main(st);

  // bound on every |branch metric| of this call: max|a-priori| + max|systematic| + max|parity|
  int* gm = a.gmax + (size_t)cb * 4;
  int  g  = 0;
  if (P::kMonitor) {
    g = dec2 ? gm[3] + gm[2] : (st0.n_iter ? gm[3] : 0) + gm[0] + gm[1];
    const bool bad = !fast16_beta_ok(m.mon_b.spread_lo(), g) || !fast16_beta_ok(m.mon_b.spread_hi(), g);
    if (__any_sync(gmask, bad)) { // the whole code block is replayed with the exact policy (mode 2 launch)
      if (j == 0)
        a.state[cb].redo = 1;
      m.src.drain();
      return;
    }
  }

  // ---- forward: warm-up, hand the estimate to the lane above, known start for lane 0, output pass
  m.alpha_warm(st);
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
  u32             ehi, elo;
  if (!dec2) {
    EpiDec1<P> e{(u32*)(ws + kPlPost * (size_t)d.ps), ws + kPlApp2 * (size_t)d.ps, T, j, d.sat_end, 0u, 0u};
    m.alpha_main(st, e);
    ehi = e.ehi;
    elo = e.elo;
  } else {
    EpiDec2<P> e{ws + kPlPost * (size_t)d.ps, ws + kPlApr * (size_t)d.ps, d.sat_end, 0u, 0u};
    m.alpha_main(st, e);
    ehi = e.ehi;
    elo = e.elo;
  }
  // max |extrinsic| handed to the next half-iteration (its a-priori / systematic input)
  int ge = max(max(lo16(ehi), hi16(ehi)), max(-lo16(elo), -hi16(elo)));
#pragma unroll
  for (int o = T / 2; o >= 1; o >>= 1)
    ge = max(ge, __shfl_xor_sync(gmask, ge, o, T));
  if (P::kMonitor) {
    const bool bad = !fast16_alpha_ok(m.mon_a.spread_lo(), m.mon_b.spread_lo(), g) ||
                     !fast16_alpha_ok(m.mon_a.spread_hi(), m.mon_b.spread_hi(), g) || (m.mon_a.ovf & 0x80008000u) != 0;
    if (__any_sync(gmask, bad)) {
      if (j == 0)
        a.state[cb].redo = 1;
      return;
    }
  }
  if (j == 0)
    gm[3] = ge;
}

// ------------------------------------------------------------------------------------------ TMA / mbarrier primitives
// (used by k_map_f16, map_f16.cuh).  Staging history: per-thread LDGSTS costs ~8 LSU cycles per warp instruction
// whatever its payload and saturated the L1 data pipe before the integer pipe; plain bulk copies (UBLKCP) take
// uniform-register operands, so a copy per (code block, plane) costs ~10 issue slots each; one tensor copy per warp
// and tile does neither.
__device__ __forceinline__ void mbar_init(unsigned bar_s, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar_s, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar_s, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "B200_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra B200_DONE;\n"
      "bra B200_WAIT;\n"
      "B200_DONE:\n"
      "}\n" ::"r"(bar_s),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_s, const void* src, unsigned bytes, unsigned bar_s)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_s), "l"(src), "r"(bytes),
               "r"(bar_s)
               : "memory");
}

// ------------------------------------------------------------------------------------------ windowed MAP, tensor-map staging
// One TMA tensor copy per warp and chunk.  The workspace of a K-group (code blocks of equal K, consecutive slots of
// the work list) is a 4-D tensor (T words | block | row | plane) with strides (4, 6*ps*2, 4T, ps*2) bytes; the box
// (T, G, L, P) is exactly what the G code blocks of a warp need for a chunk of L trellis steps: P = 3 planes
// {syst, par0, apr} for DEC1 with a-priori input, P = 2 for the first half-iteration {syst, par0} and for DEC2
// {app2, par1}.  Listing the block dimension second makes a box row 32 words = one word per lane of the warp, so
// the staging area is read without bank conflicts and with a single address per (plane, step).  Rows outside the
// tensor (top-aligned chunks start below row 0 when W is not a multiple of L) are zero-filled by the hardware.
__device__ __forceinline__ void tma_tile4(unsigned dst_s, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar_s)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst_s),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_s)
               : "memory");
}

} // namespace b200
#include "map_f16.cuh"
#include "map_lat.cuh"
#include "map_scan.cuh"
namespace b200 {
__device__ __forceinline__ uint32_t crc24_mulmod(uint32_t a, uint32_t b, uint32_t poly);
}
#include "map_fused.cuh"
namespace b200 {

// ------------------------------------------------------------------------------------------ generic MAP
// tdec_gen_dec: un-windowed, wrapping int16, natural order.  One thread decodes TWO code blocks of the same K
// (one per int16x2 half); beta lives in a global scratch interleaved by thread for coalescing.
struct GenArgs {
  const int*      work; // pairs: work[2t], work[2t+1] (second may be -1)
  int             n_pairs;
  const CbDev*    cbs;
  const CbState*  state;
  int16_t*        ws;
  const int16_t*  tails;
  const uint16_t* qpp;
  u32*            beta;  // (K_max+4) * 8 * n_threads
  int             n_threads;
};

__device__ __forceinline__ void gen_norm(uint32_t k, u32 (&o)[8])
{
  if ((k & 3u) == 0) {
#pragma unroll
    for (int i = 1; i < 8; i++)
      o[i] = p_sub_wrap(o[i], o[0]);
    o[0] = 0;
  }
}

struct Wrap16 { // turbodecoder_gen.c: plain C int16 arithmetic
  static constexpr int  kInf     = 10000;
  static constexpr bool kMonitor = false;
  B200_HD static u32 add(u32 a, u32 b) { return p_add_wrap(a, b); }
  B200_HD static u32 sub(u32 a, u32 b) { return p_sub_wrap(a, b); }
  B200_HD static u32 max(u32 a, u32 b) { return p_max(a, b); }
  B200_HD static u32 addmax(u32 a, u32 b, u32 c) { return p_addmax(a, b, c); }
  B200_HD static u32 addmax2(u32 a, u32 b, u32 c, u32 d) { return p_addmax(a, b, p_add_wrap(c, d)); }
  B200_HD static u32 sum0(u32 a, u32 b) { return p_add_wrap(a, b); }
  B200_HD static u32 summax(u32 a, u32 b, u32 c) { return p_addmax(a, b, c); }
  B200_HD static u32 sumfin(u32 m) { return m; }
  B200_HD static u32 cand(u32 a, u32 g) { return p_add_wrap(a, g); }
  B200_HD static u32 out(u32 v) { return v; }
};

__global__ void __launch_bounds__(64) k_map_gen(const GenArgs a)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_pairs)
    return;
  const int cb0 = a.work[2 * t], cb1r = a.work[2 * t + 1];
  if (cb0 < 0)
    return;
  const int     cb1 = cb1r < 0 ? cb0 : cb1r;
  const CbDev   d0 = a.cbs[cb0], d1 = a.cbs[cb1];
  const CbState s0 = a.state[cb0], s1 = a.state[cb1];
  // both halves always run (same K); results of an inactive half are simply not stored
  const bool act0 = !(s0.done || s0.n_iter >= d0.max_iter);
  const bool act1 = cb1r >= 0 && !(s1.done || s1.n_iter >= d1.max_iter);
  if (!act0 && !act1)
    return;
  const uint32_t K = d0.K;
  int16_t *      w0 = a.ws + d0.ws_off, *w1 = a.ws + d1.ws_off;
  const size_t   ps0 = d0.ps, ps1 = d1.ps;
  const bool     dec2_0 = s0.n_iter & 1u, dec2_1 = s1.n_iter & 1u;
  // per-half plane selection (the two blocks may be at different half-iterations)
  const int16_t* in0  = dec2_0 ? w0 + kPlApp2 * ps0 : w0;
  const int16_t* in1  = dec2_1 ? w1 + kPlApp2 * ps1 : w1;
  const int16_t* pa0  = dec2_0 ? w0 + kPlPar1 * ps0 : w0 + ps0;
  const int16_t* pa1  = dec2_1 ? w1 + kPlPar1 * ps1 : w1 + ps1;
  const int16_t* ap0  = (!dec2_0 && s0.n_iter) ? w0 + kPlApr * ps0 : nullptr;
  const int16_t* ap1  = (!dec2_1 && s1.n_iter) ? w1 + kPlApr * ps1 : nullptr;
  const int16_t* tl0  = a.tails + (size_t)cb0 * 12 + (dec2_0 ? 6 : 0);
  const int16_t* tl1  = a.tails + (size_t)cb1 * 12 + (dec2_1 ? 6 : 0);
  auto ldx = [&](uint32_t k, u32& x, u32& y, u32& ap) {
    int32_t x0, x1, y0, y1;
    if (k < K) {
      x0 = in0[k]; x1 = in1[k]; y0 = pa0[k]; y1 = pa1[k];
    } else {
      x0 = tl0[k - K]; x1 = tl1[k - K]; y0 = tl0[3 + k - K]; y1 = tl1[3 + k - K];
    }
    const int32_t a0 = (ap0 && k < K) ? ap0[k] : 0, a1 = (ap1 && k < K) ? ap1[k] : 0;
    ap = pack16(a0, a1);
    x  = p_add_wrap(pack16(x0, x1), ap);
    y  = pack16(y0, y1);
  };
  u32*      beta = a.beta + t;
  const int nt   = a.n_threads;
  // The recursions are a serial chain in registers, but nothing they READ depends on them: the rows of a chunk of kGenU
  // trellis steps (and, in the forward pass, the stored beta vectors) are loaded before the chunk's first step, so the chain
  // never waits for memory (a batch brings a few dozen warps at most: one load round trip per step was the whole run time)
  constexpr int kGenU = 8;
  u32           o[8];
  o[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++)
    o[i] = splat16(-Wrap16::kInf);
  for (int k0 = (int)K + 2; k0 >= 0; k0 -= kGenU) { // map_gen_beta, gen.c:71-111: k = K+2 .. 0
    u32 xs[kGenU], ys[kGenU];
#pragma unroll
    for (int u = 0; u < kGenU; u++) {
      u32 ap;
      if (k0 - u >= 0)
        ldx((uint32_t)(k0 - u), xs[u], ys[u], ap);
    }
#pragma unroll
    for (int u = 0; u < kGenU; u++) {
      const int k = k0 - u;
      if (k < 0)
        break;
      bwd_step<Wrap16>(o, xs[u], ys[u], p_add_wrap(xs[u], ys[u]));
      if (k >= 1 && k <= (int)K) {
#pragma unroll
        for (int s = 0; s < 8; s++)
          beta[((size_t)k * 8 + s) * nt] = o[s];
      }
      if ((uint32_t)k < K)
        gen_norm((uint32_t)k, o);
    }
  }
  o[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++)
    o[i] = splat16(-Wrap16::kInf);
  const uint16_t* q0 = a.qpp + d0.qpp_off;
  const uint16_t* q1 = a.qpp + d1.qpp_off;
  constexpr int   kGenA = 4;
  for (uint32_t k0 = 1; k0 <= K; k0 += kGenA) { // map_gen_alpha, gen.c:135-197: k = 1 .. K
    u32 xs[kGenA], ys[kGenA], aps[kGenA], bv[kGenA][8];
    uint32_t f0[kGenA], f1[kGenA];
#pragma unroll
    for (int u = 0; u < kGenA; u++) {
      const uint32_t k = k0 + u;
      if (k <= K) {
        ldx(k - 1, xs[u], ys[u], aps[u]);
#pragma unroll
        for (int s = 0; s < 8; s++)
          bv[u][s] = beta[((size_t)k * 8 + s) * nt];
        // scatter targets of the glue: rev[i] for decoder 1, fwd[i] for decoder 2 (per half)
        f0[u] = dec2_0 ? q0[k - 1] : q0[K + k - 1];
        f1[u] = dec2_1 ? q1[k - 1] : q1[K + k - 1];
      }
    }
#pragma unroll
    for (int u = 0; u < kGenA; u++) {
      const uint32_t k = k0 + u;
      if (k > K)
        break;
      RangeMon  nomon;
      const u32 llr = fwd_step_llr<Wrap16>(o, bv[u], xs[u], ys[u], p_add_wrap(xs[u], ys[u]), nomon);
      gen_norm(k, o);
      const uint32_t i = k - 1;
      // glue (iter.h:107-127), per half
      if (act0) {
        const int32_t l = lo16(llr);
        if (!dec2_0) {
          w0[kPlPost * ps0 + i]     = (int16_t)l;
          w0[kPlApp2 * ps0 + f0[u]] = (int16_t)(l - lo16(aps[u])); // app2[rev[i]] = ext1[i] - app1[i]
        } else {
          w0[kPlPost * ps0 + f0[u]] = (int16_t)l;
          w0[kPlApr * ps0 + f0[u]]  = (int16_t)(l - lo16(xs[u]));  // app1[fwd[i]] = ext2[i] - app2[i] (decoder 2: x is its own input)
        }
      }
      if (act1) {
        const int32_t l = hi16(llr);
        if (!dec2_1) {
          w1[kPlPost * ps1 + i]     = (int16_t)l;
          w1[kPlApp2 * ps1 + f1[u]] = (int16_t)(l - hi16(aps[u]));
        } else {
          w1[kPlPost * ps1 + f1[u]] = (int16_t)l;
          w1[kPlApr * ps1 + f1[u]]  = (int16_t)(l - hi16(xs[u]));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ decision + CRC
__device__ __forceinline__ uint32_t crc24_mulmod(uint32_t a, uint32_t b, uint32_t poly)
{
  uint32_t r = 0;
#pragma unroll 1
  for (int i = 23; i >= 0; i--) {
    r <<= 1;
    if (r & 0x1000000u)
      r ^= poly;
    if ((b >> i) & 1u)
      r ^= a;
  }
  return r & 0xffffffu;
}

// CRC24 of `nbytes` bytes produced by get_byte(i), computed by one warp: each lane runs the reference's byte
// recurrence (crc.h:56-63) over one chunk; chunks are combined with crc(A||B) = crc(A) x^(8|B|) + crc(B) mod g.
// The message is virtually left-padded with zero bytes to 32 equal chunks (leading zeros do not change a
// zero-initialised CRC), so the result equals the serial recurrence bit for bit.
// tab: 256-entry byte table of the polynomial in SHARED memory (lanes index it divergently, which constant memory
// would serialise)
// xpow[l] = x^(8*cb*2^l) mod g, cb = ceil(nbytes/32): chunk-combination constants precomputed on the host
// (crc24_xpows in engine.cu), so the warp runs 5 modular multiplications instead of 10
template <class GetByte>
__device__ __forceinline__ uint32_t warp_crc24(uint32_t nbytes, const uint32_t* tab, uint32_t poly, const uint32_t* xpow, GetByte get_byte)
{
  const int      lane = threadIdx.x & 31;
  const uint32_t cb   = (nbytes + 31) / 32, pad = 32 * cb - nbytes;
  uint32_t       crc  = 0;
  for (uint32_t q = 0; q < cb; q++) {
    const uint32_t pos = lane * cb + q;
    if (pos >= pad) {
      const uint32_t byte = get_byte(pos - pad);
      crc                 = ((crc << 8) ^ tab[((crc >> 16) & 0xffu) ^ byte]) & 0xffffffu;
    }
  }
#pragma unroll
  for (int lv = 0; lv < 5; lv++) {
    const int      l     = 1 << lv;
    const uint32_t other = __shfl_down_sync(0xffffffffu, crc, l); // chunk(s) to the right
    if ((lane & (2 * l - 1)) == 0)
      crc = crc24_mulmod(crc, xpow[lv], poly) ^ other;
  }
  return __shfl_sync(0xffffffffu, crc, 0);
}

// CRC24 of `nbytes` bytes in shared memory by a whole CTA of NT threads (every thread calls it; result in every thread): the
// chunk scheme of warp_crc24 with NT chunks -- a 9.4 KB transport block is 37 byte steps per thread instead of 295 for the
// lanes of one warp, whose two dependent shared-memory loads per byte made the CRC 20 of k_tb_finish's 45 us.  The chunk
// constants x^(8 chunk 2^l) mod g are computed here (square and multiply by thread 0, ~20 polynomial products).
// s_x: NT / 32 + 8 words of shared memory.
template <int NT>
__device__ __forceinline__ uint32_t cta_crc24(uint32_t nbytes, const uint8_t* bytes, const uint32_t* tab, uint32_t poly, uint32_t* s_x)
{
  constexpr int  kLevels = NT == 256 ? 8 : NT == 128 ? 7 : 5;
  const int      g = threadIdx.x, lane = g & 31, wid = g >> 5;
  const uint32_t cbk = (nbytes + NT - 1) / NT, pad = NT * cbk - nbytes;
  if (g == 0) {
    uint32_t r = 1u, base = 0x100u; // x^0, x^8
    for (uint32_t e = cbk; e; e >>= 1) {
      if (e & 1u)
        r = crc24_mulmod(r, base, poly);
      base = crc24_mulmod(base, base, poly);
    }
    s_x[NT / 32] = r; // x^(8 cbk)
    for (int lv = 1; lv < kLevels; lv++) {
      r                 = crc24_mulmod(r, r, poly);
      s_x[NT / 32 + lv] = r;
    }
  }
  uint32_t crc = 0;
  for (uint32_t q = 0; q < cbk; q++) {
    const uint32_t pos = (uint32_t)g * cbk + q;
    if (pos >= pad)
      crc = ((crc << 8) ^ tab[((crc >> 16) & 0xffu) ^ bytes[pos - pad]]) & 0xffffffu;
  }
  __syncthreads();
#pragma unroll
  for (int lv = 0; lv < 5; lv++) {
    const int      l     = 1 << lv;
    const uint32_t other = __shfl_down_sync(0xffffffffu, crc, l); // chunk(s) to the right
    if ((lane & (2 * l - 1)) == 0)
      crc = crc24_mulmod(crc, s_x[NT / 32 + lv], poly) ^ other;
  }
  if (lane == 0)
    s_x[wid] = crc;
  __syncthreads();
  crc = lane < NT / 32 ? s_x[lane] : 0u;
#pragma unroll
  for (int lv = 5; lv < kLevels; lv++) {
    const int      l     = 1 << (lv - 5);
    const uint32_t other = __shfl_down_sync(0xffffffffu, crc, l);
    if ((lane & (2 * l - 1)) == 0)
      crc = crc24_mulmod(crc, s_x[NT / 32 + lv], poly) ^ other;
  }
  return __shfl_sync(0xffffffffu, crc, 0);
}

// copies both CRC24 byte tables from constant to shared memory (call with all threads of the block, then sync)
// g_tab: the engine's copy of the two tables in global memory.  Constant memory serves ONE address per warp and access: 512
// different words are 512 serialised constant-cache accesses, most of them misses -- 30 of the 50 us k_tb_finish took for one
// transport block; from global memory the same copy is two coalesced loads per thread
__device__ __forceinline__ void load_crc_tables(uint32_t (*s_tab)[256], const uint32_t* g_tab = nullptr)
{
  for (int i = threadIdx.x; i < 512; i += blockDim.x)
    s_tab[i >> 8][i & 255] = g_tab ? g_tab[i] : c_crc_tab[i >> 8][i & 255];
}

struct DecideArgs {
  const int*     list;
  int            n;
  const CbDev*   cbs;
  CbState*       state;
  const int16_t* ws;
  uint8_t*       cb_out;
  uint32_t*      counters; // [0]: half-iterations replayed with the exact policy, [1]: half-iterations run,
                           // [4 + t]: code blocks still undecided after half-iteration t of this batch
  int            iter;     // half-iteration index t of this launch within the batch
  const uint32_t* crc_tab; // [2][256] CRC24A, CRC24B byte tables in global memory
};

// One warp per code block.  Hard decisions (win.h:925-993 / gen.c:260-277): the a-posteriori LLRs sit in lane
// layout (row = trellis step, N lanes per row).  Each lane loads one whole row with vector loads, and one warp
// ballot per decoder lane turns 32 rows into 32 consecutive natural-order bits of that lane; the bit rows are then
// cut into MSB-first bytes (bit n of the block is step n % W of lane n / W), the bytes go out coalesced and feed
// the warp-parallel CRC.
// Hard decisions of a windowed decoder: lane l of the warp loads row r0 + l (N int16 = NQ uint4) of the a-posteriori
// plane, one ballot per decoder lane turns 32 rows into 32 consecutive bits of that lane.  The loads of the next 32 rows
// are issued before the ballots of the current ones (the chain is otherwise one DRAM round trip per 32 rows).
template <int NQ>
__device__ __forceinline__ void decide_rows(const int16_t* post, uint32_t W, u32* bits, uint32_t wpr, int lane)
{
  constexpr uint32_t N = 8 * NQ;
  // kAhead chunks of 32 rows in flight: a lone warp pays the whole memory latency of every batch of loads (one chunk at a
  // time was 12 round trips for K = 5824; the decision kernel sits on the per-TTI latency path after every half-iteration)
  constexpr int kAhead = NQ == 4 ? 3 : 6;
  uint4         v[kAhead][NQ];
  for (uint32_t r0 = 0; r0 < W; r0 += 32 * kAhead) {
#pragma unroll
    for (int c = 0; c < kAhead; c++) {
      const uint32_t p = r0 + 32 * c + lane;
#pragma unroll
      for (int q = 0; q < NQ; q++)
        v[c][q] = p < W ? *reinterpret_cast<const uint4*>(post + (size_t)p * N + 8 * q) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int c = 0; c < kAhead; c++) {
      if (r0 + 32 * c < W) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const u32 ww[4] = {v[c][q].x, v[c][q].y, v[c][q].z, v[c][q].w};
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const int32_t val = (e & 1) ? hi16(ww[e >> 1]) : lo16(ww[e >> 1]);
            const u32     b   = __ballot_sync(0xffffffffu, val > 0);
            if (lane == 0)
              bits[(8 * q + e) * wpr + (r0 + 32 * c) / 32] = b;
          }
        }
      }
    }
  }
}

constexpr int kDecideWarps = 4;
__global__ void __launch_bounds__(kDecideWarps * 32) k_decide_crc(const DecideArgs a)
{
  __shared__ u32 s_bits[kDecideWarps][6144 / 32 + 64]; // bit rows: lane d at word d * wpr
  __shared__ u32 s_bytes[kDecideWarps][6144 / 32];     // decided bytes, 4 per word
  __shared__ uint32_t s_tab[2][256];
  if (a.iter > 0 && a.counters[4 + a.iter - 1] == 0)
    return; // every code block of the batch was finished by an earlier half-iteration
  const int wib = threadIdx.x >> 5;
  const int w   = blockIdx.x * kDecideWarps + wib;
  const int cb  = w < a.n ? a.list[w] : -1;
  CbDev     d;
  CbState*  s = nullptr;
  uint32_t  n_iter0 = 0;
  bool      active = false;
  if (cb >= 0) {
    d       = a.cbs[cb];
    s       = &a.state[cb];
    n_iter0 = s->n_iter;
    active  = !(s->done || n_iter0 >= d.max_iter);
  }
  const uint32_t n_iter = n_iter0 + 1; // the half-iteration that just ran
  const int      lane   = threadIdx.x & 31;
  const bool     need   = active && (d.crc_poly != 0 || n_iter >= d.max_iter);
  if (!__syncthreads_or(active))
    return; // none of this block's code blocks is still being decoded
  if (__syncthreads_or(need && d.crc_poly != 0)) { // (run_all semantics: no CRC, the tables are not needed)
    load_crc_tables(s_tab, a.crc_tab);
    __syncthreads();
  }
  if (!active)
    return;
  uint32_t       crc    = 1;
  if (need) {
    const int16_t* post = a.ws + d.ws_off + kPlPost * (size_t)d.ps;
    const uint32_t K = d.K, N = d.N ? d.N : 1, W = d.N ? d.W : K;
    const uint32_t wpr = (W + 31) / 32 + 1; // words per bit row (+1 so a 64-bit window never runs off the row)
    u32*           bits = s_bits[wib];
    for (uint32_t i = lane; i < N * wpr; i += 32)
      bits[i] = 0;
    __syncwarp();
    if (N == 8)
      decide_rows<1>(post, W, bits, wpr, lane);
    else if (N == 16)
      decide_rows<2>(post, W, bits, wpr, lane);
    else if (N == 32)
      decide_rows<4>(post, W, bits, wpr, lane);
    for (uint32_t r0 = 0; r0 < W && N < 8; r0 += 32) {
      const uint32_t p     = r0 + lane;
      const bool     valid = p < W;
      {
        const int32_t val = valid ? (int32_t)post[p] : 0;
        const u32     b   = __ballot_sync(0xffffffffu, val > 0);
        if (lane == 0)
          bits[r0 / 32] = b;
      }
    }
    __syncwarp();
    // bytes: natural bit n = 8*byte + i lives at bit (n % W) of row n / W; rows are LSB-first in step order
    const uint32_t nbytes = K / 8;
    u32*           obytes = s_bytes[wib];
    uint8_t*       ob8    = reinterpret_cast<uint8_t*>(obytes);
    for (uint32_t b = lane; b < nbytes; b += 32) {
      const uint32_t n = 8 * b, dl = n / W, p = n - dl * W;
      const uint32_t* row = bits + dl * wpr;
      const uint64_t  win = ((uint64_t)row[p / 32 + 1] << 32) | row[p / 32];
      u32             v   = (u32)(win >> (p & 31)) & 0xffu;
      if (p + 8 > W) { // the byte straddles two decoder lanes
        const uint32_t c1 = W - p;
        v = (v & ((1u << c1) - 1u)) | ((bits[(dl + 1) * wpr] << c1) & 0xffu);
      }
      ob8[b] = (uint8_t)(__brev(v) >> 24); // first bit in time is the MSB
    }
    __syncwarp();
    u32* out = reinterpret_cast<u32*>(a.cb_out + d.out_off);
    if ((d.out_off & 3u) == 0) {
      for (uint32_t i = lane; i < nbytes / 4; i += 32)
        out[i] = obytes[i];
      for (uint32_t i = (nbytes / 4) * 4 + lane; i < nbytes; i += 32)
        a.cb_out[d.out_off + i] = ob8[i];
    } else {
      for (uint32_t i = lane; i < nbytes; i += 32)
        a.cb_out[d.out_off + i] = ob8[i];
    }
    if (d.crc_poly != 0) {
      const bool is_a = d.crc_poly == kCrc24A;
      crc = warp_crc24(nbytes, s_tab[is_a ? 0 : 1], is_a ? kCrc24A : kCrc24B, d.crc_xp, [&](uint32_t b) -> uint32_t { return ob8[b]; });
    }
  }
  if (lane == 0) {
    s->n_iter = n_iter;
    if (!(d.crc_poly != 0 && crc == 0) && n_iter < d.max_iter)
      atomicAdd(&a.counters[4 + a.iter], 1u);
    atomicAdd(&a.counters[1], 1u);
    if (s->redo) {
      s->redo = 0;
      s->n_redo++;
      atomicAdd(&a.counters[0], 1u);
    }
    if (d.crc_poly != 0 && crc == 0) { // early stop (sch.c:441-450); the iteration limit is checked against n_iter
      s->crc_ok = 1;
      s->done   = 1;
    }
    s->crc = crc;
  }
}

} // namespace b200
#include "map_gen_fused.cuh"
namespace b200 {

// ------------------------------------------------------------------------------------------ soft demodulation
// srslte_demod_soft_demodulate_{s,b} (modem/demod_soft.c:896-945) + srslte_scrambling_{s,sb}_offset
// (scrambling/scrambling.c:43-53), the two steps between the equaliser and decode_tb (phch/pdsch.c:832-852), fused:
// one thread per symbol, Qm LLRs computed, descrambled and stored.  The reference's integer results depend on where a
// symbol sits -- its SSE bodies (groups of 4 symbols for int16, 8 for int8; 16 values for the QPSK conversion) round
// to nearest, pack with saturation and subtract integer thresholds; the scalar tails truncate, wrap and subtract float
// thresholds -- so both forms are implemented and selected by the symbol index.  Conversions reproduce cvt(t)ps2dq
// including its "integer indefinite" result for out-of-range inputs.
struct DemodDev {
  const float*   sym; // nsym complex values (re, im)
  const uint8_t* scr; // packed scrambling sequence (first bit = MSB) or nullptr
  void*          out; // nsym * Qm LLRs
  uint32_t       nsym;
  uint32_t       mod; // srslte_mod_t: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM
  const float*   csi;     // channel state information per symbol (csi_correction, pdsch.c:628-741) or nullptr
  float*         csi_max; // its maximum over the codeword, written by k_csi_max
};
struct DemodConst { // thresholds evaluated on the host exactly as the reference's expressions are (float arithmetic)
  float   qpsk_scale_s, qpsk_scale_b;
  float   thr16_s, thr16_b; // 2 * SCALE / sqrtf(10)
  int32_t off16_s, off16_b;
  int32_t off64a_s, off64b_s, off64a_b, off64b_b; // 4 * SCALE / sqrtf(42), 2 * SCALE / sqrtf(42), truncated
  float   c8, c4, c2; // 8, 4, 2 / sqrtf(170)
};
__device__ __forceinline__ int32_t cvt_rne(float f) { return f < 2147483648.0f ? __float2int_rn(f) : (int32_t)0x80000000; }
__device__ __forceinline__ int32_t cvt_trunc(float f) { return f < 2147483648.0f ? __float2int_rz(f) : (int32_t)0x80000000; }
__device__ __forceinline__ int32_t cvt_trunc_d(double d) { return d < 2147483648.0 ? __double2int_rz(d) : (int32_t)0x80000000; }
template <typename T>
__device__ __forceinline__ int32_t wrapT(int32_t v)
{
  return sizeof(T) == 2 ? (int32_t)(int16_t)(uint16_t)v : (int32_t)(int8_t)(uint8_t)v;
}
template <typename T>
__device__ __forceinline__ int32_t satT(int32_t v) // _mm_packs_epi32 [+ _mm_packs_epi16]
{
  const int32_t s16 = min(max(v, -32768), 32767);
  return sizeof(T) == 2 ? s16 : min(max(s16, -128), 127);
}
template <typename T>
__device__ __forceinline__ int32_t absT(int32_t v) { return wrapT<T>(v < 0 ? -v : v); } // abs(-min) = -min, like pabsw/pabsb

constexpr int kDemodU = 4;
// csi_max = csi[srslte_vec_max_fi(csi, nsym)] of every codeword that carries channel state information (pdsch.c:651-655)
__global__ void __launch_bounds__(256) k_csi_max(const DemodDev* __restrict__ cws)
{
  __shared__ float s_m[8];
  const DemodDev d = cws[blockIdx.x];
  if (!d.csi)
    return;
  float m = d.csi[0];
  for (uint32_t i = threadIdx.x; i < d.nsym; i += blockDim.x)
    m = fmaxf(m, d.csi[i]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
      m = fmaxf(m, s_m[w]);
    *d.csi_max = m;
  }
}

template <typename T, int MOD> // one instantiation per modulation (the host groups the codewords): no per-symbol switch
__global__ void __launch_bounds__(256) k_demod_descramble(const DemodDev* __restrict__ cws, const DemodConst c)
{
  constexpr bool S     = sizeof(T) == 2;
  const DemodDev d     = cws[blockIdx.y];
  const int      n     = (int)d.nsym;
  const int      grp   = S ? 4 : 8;              // symbols per SSE trip of the 16QAM / 64QAM bodies
  const int      n_sse = n / grp * grp;
  const int      qlen = 2 * n, q_sse = qlen >= 16 ? ((qlen - 16) / 16 + 1) * 16 : 0; // values converted by the QPSK SIMD body
  constexpr int  Qm   = MOD == 0 ? 1 : 2 * MOD;
  T*             out  = (T*)d.out;
  const bool     vec4 = (((uintptr_t)out) & 3u) == 0 && ((Qm * sizeof(T)) & 3u) == 0;
  // kDemodU symbols per thread and trip, their loads (symbol + sequence bytes) issued before the first use
  for (int i0 = blockIdx.x * (kDemodU * 256) + threadIdx.x; i0 < n; i0 += gridDim.x * (kDemodU * 256)) {
    float2   sy_u[kDemodU];
    uint32_t win_u[kDemodU];
#pragma unroll
    for (int u = 0; u < kDemodU; u++) {
      const int i = i0 + u * 256;
      if (i < n) {
        sy_u[u] = reinterpret_cast<const float2*>(d.sym)[i];
        if (d.scr) { // 16 sequence bits starting at byte b0 / 8 (the second byte only when the symbol's bits reach into it)
          const uint32_t b0 = (uint32_t)i * (uint32_t)Qm;
          win_u[u] = ((uint32_t)d.scr[b0 >> 3] << 8) | (((b0 & 7u) + (uint32_t)Qm > 8u) ? (uint32_t)d.scr[(b0 >> 3) + 1] : 0u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kDemodU; u++) {
    const int i = i0 + u * 256;
    if (i >= n)
      break;
    const float  re = sy_u[u].x, im = sy_u[u].y;
    int32_t      v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    switch (MOD) {
      case 0: { // demod_bpsk_lte_{s,b}
        const float t = __fmul_rn(S ? -100.0f : -20.0f, __fadd_rn(re, im));
        v[0]          = wrapT<T>(cvt_trunc_d(__dmul_rn((double)t, 0.70710678118654752440)));
        break;
      }
      case 1: { // srslte_vec_convert_f{i,b}: truncation; the SIMD body saturates, the tail wraps
        const float sc = S ? c.qpsk_scale_s : c.qpsk_scale_b;
        const int32_t a = cvt_trunc(__fmul_rn(re, sc)), b = cvt_trunc(__fmul_rn(im, sc));
        v[0] = 2 * i < q_sse ? satT<T>(a) : wrapT<T>(a);
        v[1] = 2 * i + 1 < q_sse ? satT<T>(b) : wrapT<T>(b);
        break;
      }
      case 2: {
        const float sc = S ? 400.0f : 30.0f;
        if (i < n_sse) {
          const int32_t off = S ? c.off16_s : c.off16_b;
          v[0] = satT<T>(cvt_rne(__fmul_rn(re, -sc)));
          v[1] = satT<T>(cvt_rne(__fmul_rn(im, -sc)));
          v[2] = wrapT<T>(absT<T>(v[0]) - off);
          v[3] = wrapT<T>(absT<T>(v[1]) - off);
        } else {
          const float   thr = S ? c.thr16_s : c.thr16_b;
          const int32_t yre = wrapT<T>(cvt_trunc(__fmul_rn(sc, re))), yim = wrapT<T>(cvt_trunc(__fmul_rn(sc, im)));
          v[0] = wrapT<T>(-yre);
          v[1] = wrapT<T>(-yim);
          v[2] = wrapT<T>(cvt_trunc(__fsub_rn((float)abs(yre), thr)));
          v[3] = wrapT<T>(cvt_trunc(__fsub_rn((float)abs(yim), thr)));
        }
        break;
      }
      case 3: {
        const float   sc = S ? 700.0f : 40.0f;
        const int32_t o1 = S ? c.off64a_s : c.off64a_b, o2 = S ? c.off64b_s : c.off64b_b;
        if (i < n_sse) {
          v[0] = satT<T>(cvt_rne(__fmul_rn(re, -sc)));
          v[1] = satT<T>(cvt_rne(__fmul_rn(im, -sc)));
          v[2] = wrapT<T>(absT<T>(v[0]) - o1);
          v[3] = wrapT<T>(absT<T>(v[1]) - o1);
          v[4] = wrapT<T>(absT<T>(v[2]) - o2);
          v[5] = wrapT<T>(absT<T>(v[3]) - o2);
        } else {
          const int32_t yre = wrapT<T>(cvt_trunc(__fmul_rn(sc, re))), yim = wrapT<T>(cvt_trunc(__fmul_rn(sc, im)));
          v[0] = wrapT<T>(-yre);
          v[1] = wrapT<T>(-yim);
          v[2] = wrapT<T>(wrapT<T>(abs(yre)) - o1);
          v[3] = wrapT<T>(wrapT<T>(abs(yim)) - o1);
          v[4] = wrapT<T>(wrapT<T>(abs(v[2])) - o2);
          v[5] = wrapT<T>(wrapT<T>(abs(v[3])) - o2);
        }
        break;
      }
      default: { // demod_256qam_lte_{s,b}: scalar float arithmetic, truncating conversions
        const float sc = S ? 1000.0f : 50.0f;
        float       r = -re, q = -im;
        v[0] = wrapT<T>(cvt_trunc(__fmul_rn(sc, r)));
        v[1] = wrapT<T>(cvt_trunc(__fmul_rn(sc, q)));
        r = __fsub_rn(fabsf(r), c.c8);
        q = __fsub_rn(fabsf(q), c.c8);
        v[2] = wrapT<T>(cvt_trunc(__fmul_rn(sc, r)));
        v[3] = wrapT<T>(cvt_trunc(__fmul_rn(sc, q)));
        r = __fsub_rn(fabsf(r), c.c4);
        q = __fsub_rn(fabsf(q), c.c4);
        v[4] = wrapT<T>(cvt_trunc(__fmul_rn(sc, r)));
        v[5] = wrapT<T>(cvt_trunc(__fmul_rn(sc, q)));
        r = __fsub_rn(fabsf(r), c.c2);
        q = __fsub_rn(fabsf(q), c.c2);
        v[6] = wrapT<T>(cvt_trunc(__fmul_rn(sc, r)));
        v[7] = wrapT<T>(cvt_trunc(__fmul_rn(sc, q)));
        break;
      }
    }
    // csi_correction (pdsch.c:628-741), between the demodulator and the descrambler: soft bits scaled by the channel state
    // information of their resource element, with the arithmetic of the reference's x86 build (see oracle.c:orc_csi_correction):
    // SSE bodies e = (e * sat16(rne(csi * (32767 / max)))) >> 16 -- and in the QPSK / 64QAM bodies the two symbols of a pair
    // swap (part of) their csi (_mm_blend_ps operand order) --, scalar tails and the int8 path e = (T)((float)e * (csi / max))
    if (d.csi) {
      const float    cmax = *d.csi_max;
      const uint32_t NB   = (uint32_t)n * (uint32_t)Qm;
      // symbols covered by the SSE body of this modulation (int16 only)
      const uint32_t body = !S ? 0u : MOD == 1 ? 2u * (NB / 4u) : MOD == 2 ? NB / 4u : MOD == 3 ? 2u * (NB / 12u) : MOD == 4 ? NB / 8u : 0u;
      if ((uint32_t)i < body) {
        const float scale = __fdiv_rn(32767.0f, cmax);
        auto        c16   = [&](float cs) -> int32_t { return min(max(cvt_rne(__fmul_rn(cs, scale)), -32768), 32767); };
        const int32_t own = c16(d.csi[i]);
        const int32_t oth = (MOD == 1 || MOD == 3) ? c16(d.csi[i ^ 1]) : own;
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < Qm) {
            int32_t cc = own;
            if (MOD == 1)
              cc = oth;
            if (MOD == 3 && (((i & 1) == 0 && k >= 4) || ((i & 1) == 1 && k < 2)))
              cc = oth;
            v[k] = (v[k] * cc) >> 16; // _mm_mulhi_pi16
          }
      } else {
        const float cs = __fdiv_rn(d.csi[i], cmax);
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < Qm)
            v[k] = wrapT<T>(cvt_trunc(__fmul_rn((float)v[k], cs)));
      }
    }
    // descrambling: wrapping negation where the sequence bit is 1 (srslte_vec_neg_* against c_short / c_char = +-1)
    const uint32_t b0 = (uint32_t)i * (uint32_t)Qm;
    if (d.scr) {
      const uint32_t win = win_u[u];
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k < Qm && ((win >> (15 - (b0 & 7u) - k)) & 1u))
          v[k] = wrapT<T>(-v[k]);
    }
    if (vec4) {
      u32*      o  = reinterpret_cast<u32*>(out + b0);
      const int nw = Qm * (int)sizeof(T) / 4;
#pragma unroll
      for (int w = 0; w < 4; w++)
        if (w < nw)
          o[w] = S ? pack16(v[2 * w], v[2 * w + 1])
                   : ((u32)(uint8_t)v[4 * w] | ((u32)(uint8_t)v[4 * w + 1] << 8) | ((u32)(uint8_t)v[4 * w + 2] << 16) | ((u32)(uint8_t)v[4 * w + 3] << 24));
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k < Qm)
          out[b0 + k] = (T)v[k];
    }
    }
  }
}

// ------------------------------------------------------------------------------------------ TB finish
struct TbArgs {
  const TbDev*   tbs;
  int            n_tb;
  const CbDev*   cbs;
  const CbState* state;
  const uint8_t* cb_out;
  TbResult*      res;
  const uint32_t* crc_tab; // [2][256] CRC24A, CRC24B byte tables in global memory
};

// One CTA per transport block: assemble the payload (sch.c:390,422-424,462-467), TB CRC24A (sch.c:546-552), HARQ
// bookkeeping (sch.c:469-484).  All threads copy; warp 0 computes the CRC over the assembled bytes.
constexpr int kTbThreads = 256;
constexpr uint32_t kTbStage = 38 * 1024; // bytes of a transport block staged in shared memory for its CRC (TBS <= 299 856 bits: LTE's largest)
__global__ void __launch_bounds__(kTbThreads) k_tb_finish(const TbArgs a)
{
  __shared__ uint32_t s_tab[2][256];
  __shared__ uint32_t s_ok, s_iter;
  const int   w    = blockIdx.x;
  const int   tid  = threadIdx.x;
  const TbDev t    = a.tbs[w];
  uint8_t*    data = t.data;
  uint8_t*    hdat = t.hdata;
  uint8_t*    hcrc = t.hcrc;
  load_crc_tables(s_tab, a.crc_tab);
  if (tid == 0) {
    s_ok   = 0;
    s_iter = 0;
  }
  __syncthreads();
  // one thread per code block: the descriptor / state loads of a block are independent of the others', and so are the copies
  // of the blocks' bytes below (a loop over the blocks paid four dependent round trips per block: 45 of the kernel's 52 us)
  __shared__ const uint8_t* s_src[32];
  __shared__ uint32_t       s_off[32], s_nb[32];
  if ((uint32_t)tid < t.C) {
    const uint32_t c    = tid;
    const uint32_t cb   = t.first_cb + c;
    const CbDev    d    = a.cbs[cb];
    const uint32_t rlen = t.rlen_bytes[c < t.C1 ? 0 : 1];
    s_off[c] = c * rlen;
    if (d.skip) {
      atomicOr(&s_ok, 1u << c);
      s_src[c] = hdat + (size_t)c * 768;
      s_nb[c]  = rlen;
    } else {
      const CbState st = a.state[cb];
      if (st.crc_ok)
        atomicOr(&s_ok, 1u << c);
      atomicAdd(&s_iter, st.n_iter);
      s_src[c] = a.cb_out + d.out_off;
      // the reference writes K/8 bytes per CB; all but the last CB's trailing CRC bytes are overwritten by the
      // next CB's copy, so write payload only, plus the 3 CRC bytes of the last CB
      s_nb[c] = (c + 1 == t.C) ? d.K / 8 : rlen;
    }
  }
  // data[tbs/8 .. +2] = 0 happens before the CB loop in the reference; the payload copies below overwrite it
  if (tid < 3)
    data[t.tbs / 8 + tid] = 0;
  __syncthreads();
  const uint32_t ok_mask = s_ok;
#pragma unroll 8
  for (uint32_t idx = tid; idx < t.C * 768u; idx += kTbThreads) { // (a block has at most 768 bytes; the regions are disjoint)
    const uint32_t c = idx / 768u, i = idx - c * 768u;
    if (i < s_nb[c])
      data[(size_t)s_off[c] + i] = s_src[c][i];
  }
  __syncthreads();
  const bool all_ok = ok_mask == (t.C >= 32 ? 0xffffffffu : ((1u << t.C) - 1u));
  if (hcrc) {
    for (uint32_t c = tid; c < t.C; c += kTbThreads)
      hcrc[c] = (ok_mask >> c) & 1u;
    if (!all_ok) {
      for (uint32_t c = 0; c < t.C; c++) {
        if (((ok_mask >> c) & 1u) && !a.cbs[t.first_cb + c].skip) {
          const uint32_t rlen = t.rlen_bytes[c < t.C1 ? 0 : 1];
          for (uint32_t i = tid; i < rlen; i += kTbThreads)
            hdat[(size_t)c * 768 + i] = data[(size_t)c * rlen + i];
        }
      }
    }
  }
  // The transport block's bytes for the CRC come through shared memory: fetched by the whole CTA with independent loads (one
  // round trip), not byte by byte from global memory inside the CRC recurrence of one warp
  __shared__ __align__(16) uint8_t s_tb[kTbStage];
  const uint32_t     n_tb   = t.tbs / 8;
  const bool         staged = all_ok && n_tb <= kTbStage;
  if (staged) {
    if ((reinterpret_cast<uintptr_t>(data) & 3u) == 0) {
      const uint32_t* d32 = reinterpret_cast<const uint32_t*>(data);
      uint32_t*       s32 = reinterpret_cast<uint32_t*>(s_tb);
#pragma unroll 8
      for (uint32_t i = tid; i < (n_tb + 3) / 4; i += kTbThreads)
        s32[i] = d32[i]; // (the last word may reach into the three CRC bytes that follow the payload: written above)
    } else {
#pragma unroll 8
      for (uint32_t i = tid; i < n_tb; i += kTbThreads)
        s_tb[i] = data[i];
    }
  }
  __syncthreads();
  __shared__ uint32_t s_x[kTbThreads / 32 + 8];
  // ... by all threads when the block is long; a short block (the all-sizes workload c3 brings 12 000 of at most 768 bytes in
  // one batch) is cheaper on one warp with the host's chunk constants than the CTA-wide scheme's fixed cost (constants by
  // square and multiply in thread 0, two combination stages: ~4 us per CTA, 330 us per c3 batch)
  const bool use_cta = staged && n_tb > 3072;
  uint32_t   crc_cta = 0;
  if (use_cta) // (uniform over the CTA)
    crc_cta = cta_crc24<kTbThreads>(n_tb, s_tb, s_tab[0], kCrc24A, s_x);
  if (tid < 32) {
    int32_t  ret    = -1;
    uint32_t par_rx = 0;
    if (all_ok) {
      par_rx = use_cta ? crc_cta
                       : staged ? warp_crc24(n_tb, s_tab[0], kCrc24A, t.crc_xp, [&](uint32_t b) -> uint32_t { return s_tb[b]; })
                                : warp_crc24(n_tb, s_tab[0], kCrc24A, t.crc_xp, [&](uint32_t b) -> uint32_t { return data[b]; });
      const uint32_t o      = t.tbs / 8;
      const uint32_t par_tx = ((uint32_t)data[o] << 16) | ((uint32_t)data[o + 1] << 8) | data[o + 2];
      ret                   = (par_rx == par_tx && par_rx != 0) ? 0 : -1;
    }
    if (tid == 0) {
      a.res[w].ret      = ret;
      a.res[w].sum_iter = s_iter;
      a.res[w].cb_crc   = ok_mask;
      a.res[w].par_rx   = par_rx;
    }
  }
}

// plain CRC of a byte buffer (device side of srslte_crc_checksum_byte); one warp, generic order <= 24 handled by
// the host for orders other than 24
struct CrcXp {
  uint32_t v[5];
};
__global__ void k_crc24_bytes(const uint8_t* data, uint32_t nbytes, int tab, uint32_t poly, CrcXp xp, uint32_t* out)
{
  __shared__ uint32_t s_tab[2][256];
  load_crc_tables(s_tab);
  __syncthreads();
  const uint32_t c = warp_crc24(nbytes, s_tab[tab], poly, xp.v, [&](uint32_t b) -> uint32_t { return data[b]; });
  if (threadIdx.x == 0)
    *out = c;
}

// ------------------------------------------------------------------------------------------ transmit mirror
// encode_tb_off (phch/sch.c:235-349) = TB CRC24A, segmentation, CB CRC24B, srslte_tcod_encode_lut (fec/turbocoder.c:190-372)
// and srslte_rm_turbo_tx_lut (fec/rm_turbo.c:349-395), for many transport blocks at once.
struct EncTbDev {
  const uint8_t* data;    // tbs/8 payload bytes
  uint32_t*      e_words; // packed e-bits of the TB (first bit = MSB of byte 0), 4-byte aligned; zero-filled by k_enc_tb_crc
  uint32_t       n_words; // (nof_e_bits + 31) / 32
  uint32_t       tbs;
  uint32_t       crc;     // CRC24A of the payload, written by k_enc_tb_crc
  uint32_t       crc_xp[5];
};
struct EncCbDev {
  uint32_t tb;        // owning transport block
  uint32_t K;
  uint32_t rp_bytes;  // offset of this block's payload in the TB
  uint32_t nd_bytes;  // payload bytes taken from the TB data
  uint32_t last;      // last block of the TB: the 3 TB CRC bytes follow the payload
  uint32_t cb_crc;    // C > 1: CRC24B over payload (+ TB CRC) is appended
  uint32_t E;         // rate-matched bits of this block
  uint32_t wp;        // bit offset of this block in the TB's e-bit stream
  uint32_t qpp_off;   // natural-order QPP table fwd[K] in the QPP pool
  uint32_t rm_off;    // base rate-matching table (standard layout) in the rm pool
  uint32_t rm_start;  // rank of the first transmitted entry for this rv
  uint32_t crc_xp[5]; // CRC24B combination constants for (K - 24) / 8 bytes
};

__global__ void __launch_bounds__(128) k_enc_tb_crc(EncTbDev* __restrict__ tbs, int n_tb)
{
  __shared__ uint32_t s_tab[2][256];
  load_crc_tables(s_tab);
  __syncthreads();
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= n_tb)
    return;
  const EncTbDev t = tbs[w];
  const uint32_t c = warp_crc24(t.tbs / 8, s_tab[0], kCrc24A, t.crc_xp, [&](uint32_t b) -> uint32_t { return t.data[b]; });
  if ((threadIdx.x & 31) == 0)
    tbs[w].crc = c;
  for (uint32_t i = threadIdx.x & 31; i < t.n_words; i += 32) // code blocks OR their boundary words into place
    t.e_words[i] = 0;
}

// recursive systematic convolutional encoder g0 = 1 + D^2 + D^3 (feedback), g1 = 1 + D + D^3: one step on the packed
// state (bit 0 = newest register).  Returns the parity bit.
__device__ __forceinline__ uint32_t rsc_step(uint32_t& st, uint32_t u)
{
  const uint32_t s0 = st & 1u, s1 = (st >> 1) & 1u, s2 = (st >> 2) & 1u;
  const uint32_t fb = u ^ s1 ^ s2;
  st                = ((st << 1) & 6u) | fb;
  return fb ^ s0 ^ s2;
}

// One CTA per code block.  The two constituent encoders run chunk-parallel: every thread first encodes its chunk from the
// zero state, the chunk start states follow from a 256-step scan with the (linear) zero-input transition of one chunk
// length, then every thread encodes its chunk again from its true start state.
__global__ void __launch_bounds__(256) k_enc_cb(const EncCbDev* __restrict__ cbs, const EncTbDev* __restrict__ tbs, const uint16_t* __restrict__ qpp,
                                                const uint16_t* __restrict__ rm_pool)
{
  extern __shared__ __align__(16) uint8_t s_enc[];
  __shared__ uint32_t s_tab[2][256];
  __shared__ uint8_t  s_end[2][256], s_start[2][256], s_zi[8];
  const EncCbDev d  = cbs[blockIdx.x];
  const EncTbDev t  = tbs[d.tb];
  const uint32_t K = d.K, nb = K / 8;
  uint8_t*       bytes = s_enc;                      // K/8 block bytes (payload [+ TB CRC] [+ CB CRC])
  uint8_t*       c     = s_enc + 768;                // K bits
  uint8_t*       buf   = c + kMaxKDev;               // 3K + 12 coded bits in the decoder's input order
  const int      tid = threadIdx.x;
  load_crc_tables(s_tab);
  for (uint32_t i = tid; i < d.nd_bytes; i += 256)
    bytes[i] = t.data[d.rp_bytes + i];
  if (d.last && tid < 3)
    bytes[d.nd_bytes + tid] = (uint8_t)(t.crc >> (16 - 8 * tid));
  __syncthreads();
  if (d.cb_crc && tid < 32) {
    const uint32_t crc = warp_crc24(nb - 3, s_tab[1], kCrc24B, d.crc_xp, [&](uint32_t b) -> uint32_t { return bytes[b]; });
    if (tid < 3)
      bytes[nb - 3 + tid] = (uint8_t)(crc >> (16 - 8 * tid));
  }
  __syncthreads();
  for (uint32_t i = tid; i < K; i += 256)
    c[i] = (bytes[i >> 3] >> (7 - (i & 7))) & 1u;
  // zero-input transition of one chunk length
  const uint32_t L = (K + 255) / 256, nc = (K + L - 1) / L;
  if (tid < 8) {
    uint32_t st = tid;
    for (uint32_t i = 0; i < L; i++)
      rsc_step(st, 0);
    s_zi[tid] = (uint8_t)st;
  }
  __syncthreads();
  const uint16_t* fwd = qpp + d.qpp_off;
  const uint32_t  lo = tid * L, hi = min(K, lo + L);
  // pass 1: chunk end states from the zero state (encoder 0: natural order, encoder 1: interleaved order)
  if ((uint32_t)tid < nc) {
    uint32_t s0 = 0, s1 = 0;
    for (uint32_t i = lo; i < hi; i++) {
      rsc_step(s0, c[i]);
      rsc_step(s1, c[fwd[i]]);
    }
    s_end[0][tid] = (uint8_t)s0;
    s_end[1][tid] = (uint8_t)s1;
  }
  __syncthreads();
  if (tid == 0 || tid == 32) {
    const int e  = tid >> 5;
    uint32_t  st = 0;
    for (uint32_t k = 0; k < nc; k++) {
      s_start[e][k] = (uint8_t)st;
      st            = s_zi[st] ^ s_end[e][k]; // linear: response to the start state + response to the input
    }
  }
  __syncthreads();
  // pass 2: parity bits; the last chunk continues into the trellis termination
  if ((uint32_t)tid < nc) {
    uint32_t s0 = s_start[0][tid], s1 = s_start[1][tid];
    for (uint32_t i = lo; i < hi; i++) {
      buf[3 * i]     = c[i];
      buf[3 * i + 1] = (uint8_t)rsc_step(s0, c[i]);
      buf[3 * i + 2] = (uint8_t)rsc_step(s1, c[fwd[i]]);
    }
    if ((uint32_t)tid == nc - 1) {
      // tail: the input that drives the feedback to zero, and its parity; x z x z x z | x' z' x' z' x' z'
      for (int k = 0; k < 3; k++) {
        const uint32_t x0 = ((s0 >> 1) ^ (s0 >> 2)) & 1u, x1 = ((s1 >> 1) ^ (s1 >> 2)) & 1u;
        buf[3 * K + 2 * k]         = (uint8_t)x0;
        buf[3 * K + 2 * k + 1]     = (uint8_t)rsc_step(s0, x0);
        buf[3 * K + 6 + 2 * k]     = (uint8_t)x1;
        buf[3 * K + 6 + 2 * k + 1] = (uint8_t)rsc_step(s1, x1);
      }
    }
  }
  __syncthreads();
  // rate matching (gather through the same table the receiver scatters with) + packing into the TB's e-bit words
  const uint16_t* tab = rm_pool + d.rm_off;
  const uint32_t  Lr = 3 * K + 12, w0 = d.wp >> 5, w1 = (d.wp + d.E + 31) >> 5;
  for (uint32_t w = w0 + tid; w < w1; w += 256) {
    uint32_t word = 0;
    const uint32_t b_lo = max(w << 5, d.wp), b_hi = min((w + 1) << 5, d.wp + d.E);
    uint32_t       pos = (b_lo - d.wp + d.rm_start) % Lr;
    for (uint32_t bb = b_lo; bb < b_hi; bb++) {
      const uint32_t bit = buf[tab[pos]];
      pos                = pos + 1 == Lr ? 0 : pos + 1;
      word |= bit << (8 * ((bb & 31) >> 3) + 7 - (bb & 7)); // first bit of the stream = MSB of byte 0
    }
    if (b_lo == (w << 5) && b_hi == ((w + 1) << 5))
      t.e_words[w] = word; // interior word: owned by this block
    else
      atomicOr(&t.e_words[w], word); // shared with the neighbouring block
  }
}

// ------------------------------------------------------------------------------------------ ALU roofline probe
// Issue rate of ONE packed instruction at a time, the way the MAP kernels use it: 8 independent dependency chains
// per thread, 32 operations per chain and loop trip, every operation exactly one SASS instruction (checked with
// cuobjdump: VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2 / VIMNMX3.S16x2).  op 4 is the saturating add
// (__vaddss2, a multi-instruction emulation on sm_100a), counted per source operation.
// Measured on B200: every one of the four native instructions issues at 0.5 warp-instructions per clock and SM
// sub-partition (the integer pipe is 16 lanes wide), i.e. 64 packed operations = 128 int16 lane-ops per clock and SM.
template <int OP>
__global__ void __launch_bounds__(256) k_alu_probe(u32* out, int iters, u32 seed)
{
  u32 a[8];
#pragma unroll
  for (int i = 0; i < 8; i++)
    a[i] = seed * (threadIdx.x + 1) + i * 0x00010003u;
  const u32 g = seed | 0x00010001u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (OP == 0)
          a[i] = p_add_wrap(a[i], g);
        if (OP == 1)
          a[i] = p_max(a[i], a[(i + 1) & 7]);
        if (OP == 2)
          a[i] = p_addmax(a[i], g, a[(i + 3) & 7]);
        if (OP == 3)
          a[i] = p_max3(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);
        if (OP == 4)
          a[i] = p_add_sat(a[i], g);
      }
    }
  }
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
    r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ------------------------------------------------------------------------------------------ PUSCH pre-steps
// The data movement of srslte_ulsch_decode (phch/sch.c:1105-1180) between the descrambler and decode_tb; geometry and
// index arithmetic in ulsch_core.cuh.  A CTA moves a tile of kUlRows rows of the interleaver matrix through shared
// memory: per column one contiguous run of q_bits in, one contiguous run of g_bits out.  The kernel is HBM-bound work
// whose enemy is the index arithmetic per 32-bit word, so: W (words per symbol) is a template parameter; full tiles
// load with 128-bit accesses, all of a thread's loads in flight before the first use; the ACK symbols (bottom rows) are
// zeroed in shared memory; all rows above the RI symbols are stored with a thread mapping whose (column, word) is
// fixed per thread (only the row advances), and only the few bottom rows that hold RI symbols take the per-word path.
template <int W>
__global__ void __launch_bounds__(256, 4) k_ulsch_deinterleave(const UlschDev* __restrict__ tbs)
{
  constexpr uint32_t cs = kUlRows * W + 4; // column stride: 16-byte aligned columns, 4 banks apart
  __shared__ __align__(16) u32 s_tile[kUlMaxCols * cs];
  const UlschDev d = tbs[blockIdx.y];
  const uint32_t rows = d.rows, cols = d.cols, rw = cols * W;
  const uint32_t inv_cols  = d.inv_cols;                 // sym / cols == sym * inv_cols >> 16 for the symbols of a tile
  const uint32_t ack_rows  = (d.q_ack + 3) / 4, ri_rows = (d.q_ri + 3) / 4; // bottom rows that hold ACK / RI symbols
  const uint32_t cqi_words = d.uci ? d.q_cqi * W : 0u;
  const bool     vec = ((rows * W) & 3u) == 0 && (reinterpret_cast<uintptr_t>(d.q) & 15u) == 0;
  const int16_t* qe  = reinterpret_cast<const int16_t*>(d.q);
  u32*           cqi = d.uci ? reinterpret_cast<u32*>(d.uci + 2 * W * (d.q_ack + d.q_ri)) : nullptr;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t j0 = blockIdx.x * kUlRows; j0 < rows; j0 += gridDim.x * kUlRows) {
    const uint32_t rt = min((uint32_t)kUlRows, rows - j0), run = rt * W, half = run / 2;
    __syncthreads();
    // ---- in: per column one contiguous run of q_bits
    if (rt == kUlRows && vec) {
      constexpr uint32_t kRunV = kUlRows * W / 4, kTrips = (kUlMaxCols * kRunV + 255) / 256;
      uint4              v[kTrips];
#pragma unroll
      for (uint32_t k = 0; k < kTrips; k++) {
        const uint32_t idx = threadIdx.x + 256 * k, c = idx / kRunV, xv = idx - c * kRunV;
        if (c < cols)
          v[k] = *reinterpret_cast<const uint4*>(d.q + ((size_t)c * rows + j0) * W + 4 * xv);
      }
#pragma unroll
      for (uint32_t k = 0; k < kTrips; k++) {
        const uint32_t idx = threadIdx.x + 256 * k, c = idx / kRunV, xv = idx - c * kRunV;
        if (c < cols)
          *reinterpret_cast<uint4*>(&s_tile[c * cs + 4 * xv]) = v[k];
      }
    } else {
      for (uint32_t u = warp; u < 2 * cols; u += 8) { // one half of a column's run per warp and trip
        const uint32_t c = u >> 1, x0 = (u & 1) ? half : 0u, x1 = (u & 1) ? run : half;
        const u32*     src = d.q + ((size_t)c * rows + j0) * W;
#pragma unroll 4
        for (uint32_t x = x0 + lane; x < x1; x += 32)
          s_tile[c * cs + x] = src[x];
      }
    }
    __syncthreads();
    if (j0 + rt + ack_rows > rows) { // ACK symbols are zeroed before the de-interleaver sees them (sch.c:1067-1070)
      for (uint32_t idx = threadIdx.x; idx < d.q_ack * W; idx += 256) {
        const uint32_t r = idx / W, w = idx - r * W, j = rows - 1 - r / 4;
        if (j >= j0 && j < j0 + rt)
          s_tile[ul_col(d.ack_cols, r) * cs + (j - j0) * W + w] = 0;
      }
      __syncthreads();
    }
    // ---- out: the rows above the RI symbols leave as one contiguous run; (column, word) of a thread never change,
    //      its row advances by rpp
    const uint32_t n_plain = j0 + rt + ri_rows <= rows ? rt : (j0 + ri_rows >= rows ? 0u : rows - ri_rows - j0);
    const size_t   base = (size_t)j0 * rw;
    {
      u32*           dst = d.g + base;
      const uint32_t rpp = d.rpp;
      if (threadIdx.x < rpp * rw) {
        const uint32_t r0 = (threadIdx.x * d.inv_rw) >> 16, t = threadIdx.x - r0 * rw, c = t / W, w = t - c * W;
        const u32*     sp = s_tile + c * cs + r0 * W + w;
        u32*           gp = dst + threadIdx.x;
#pragma unroll 4
        for (uint32_t jr = r0; jr < n_plain; jr += rpp, sp += rpp * W, gp += rpp * rw)
          *gp = *sp;
      }
      if (n_plain && (base < cqi_words || (j0 == 0 && d.clobber > 0))) {
        // the head of g_bits: the reference's clobbered first LLR, and the CQI LLRs (sch.c:1152-1171)
        const uint32_t n_fix = (uint32_t)min((size_t)n_plain * rw, max((size_t)cqi_words, base + 1) - base);
        for (uint32_t idx = threadIdx.x; idx < n_fix; idx += 256) {
          const uint32_t sym = idx / W, w = idx - sym * W;
          const uint32_t jr = (sym * inv_cols) >> 16, c = sym - jr * cols;
          u32            v  = s_tile[c * cs + jr * W + w];
          if (base + idx == 0 && d.clobber > 0) {
            v      = (v & 0xffff0000u) | (uint16_t)qe[d.clobber];
            dst[0] = v; // same thread as the plain store above (thread 0), program order
          }
          if (base + idx < cqi_words)
            cqi[base + idx] = v;
        }
      }
    }
    // ---- out: the bottom rows, RI symbols skipped (ulsch_interleave_gen, sch.c:658-679)
    for (uint32_t idx = n_plain * rw + threadIdx.x; idx < rt * rw; idx += 256) {
      const uint32_t sym = idx / W, w = idx - sym * W;
      const uint32_t jr = (sym * inv_cols) >> 16, c = sym - jr * cols;
      const uint32_t m    = rows - 1 - (j0 + jr);
      const uint32_t n_ri = ul_row_count(m, d.q_ri);
      if (ul_holds(n_ri, d.ri_cols, c))
        continue;
      const uint32_t o = ((j0 + jr) * cols + c - ul_ri_before(m, n_ri, d.q_ri, d.ri_cols, c)) * W + w;
      u32            v = s_tile[c * cs + jr * W + w];
      if (o == 0 && d.clobber > (int32_t)(((size_t)c * rows + j0 + jr) * 2 * W)) // the later store to index 0 wins
        v = (v & 0xffff0000u) | (uint16_t)qe[d.clobber];
      d.g[o] = v;
      if (o < cqi_words)
        cqi[o] = v;
    }
  }
  // LLRs at the ACK and RI positions, in the order srslte_uci_decode_ack_ri walks them (uci.c:843-857)
  if (blockIdx.x == 0 && d.uci) {
    constexpr uint32_t Qm = 2 * W;
    for (uint32_t idx = threadIdx.x; idx < (d.q_ack + d.q_ri) * Qm; idx += blockDim.x) {
      const uint32_t r0 = idx / Qm, k = idx - r0 * Qm;
      const bool     ri = r0 >= d.q_ack;
      d.uci[idx] = qe[ul_uci_element(ri ? d.ri_cols : d.ack_cols, ri ? r0 - d.q_ack : r0, rows, Qm, k)];
    }
  }
}

} // namespace b200
