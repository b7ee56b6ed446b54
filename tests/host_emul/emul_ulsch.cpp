// CPU emulation of k_ulsch_deinterleave (srsran_b200/csrc/kernels.cuh): the same tile loops (half a column run per warp
// and trip in, one contiguous run out, the plain-tile fast path) and the very index functions of ulsch_core.cuh, compiled
// with g++; the "threads" of a CTA run one after the other between the barriers.
// TEST INFRASTRUCTURE (tests/test_host_emul.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../srsran_b200/csrc/ulsch_core.cuh"

using namespace b200;

template <int W>
static void emul_cta(const UlschDev& d, uint32_t bx, uint32_t grid_x)
{
  constexpr uint32_t cs = kUlRows * W + 4;
  std::vector<u32>   s_tile(kUlMaxCols * cs);
  const uint32_t     rows = d.rows, cols = d.cols, rw = cols * W;
  const uint32_t inv_cols  = d.inv_cols;                 // sym / cols == sym * inv_cols >> 16 for the symbols of a tile
  const uint32_t     ack_rows  = (d.q_ack + 3) / 4, ri_rows = (d.q_ri + 3) / 4;
  const uint32_t     cqi_words = d.uci ? d.q_cqi * W : 0u;
  const bool         vec = ((rows * W) & 3u) == 0 && (reinterpret_cast<uintptr_t>(d.q) & 15u) == 0;
  const int16_t*     qe  = reinterpret_cast<const int16_t*>(d.q);
  u32*               cqi = d.uci ? reinterpret_cast<u32*>(d.uci + 2 * W * (d.q_ack + d.q_ri)) : nullptr;
  for (uint32_t j0 = bx * kUlRows; j0 < rows; j0 += grid_x * kUlRows) {
    const uint32_t rt = std::min<uint32_t>(kUlRows, rows - j0), run = rt * W, half = run / 2;
    std::fill(s_tile.begin(), s_tile.end(), 0xdeadbeefu);
    if (rt == kUlRows && vec) {
      constexpr uint32_t kRunV = kUlRows * W / 4, kTrips = (kUlMaxCols * kRunV + 255) / 256;
      for (uint32_t tid = 0; tid < 256; tid++)
        for (uint32_t k = 0; k < kTrips; k++) {
          const uint32_t idx = tid + 256 * k, c = idx / kRunV, xv = idx - c * kRunV;
          if (c < cols)
            memcpy(&s_tile[c * cs + 4 * xv], d.q + ((size_t)c * rows + j0) * W + 4 * xv, 16);
        }
    } else {
      for (uint32_t warp = 0; warp < 8; warp++)
        for (uint32_t lane = 0; lane < 32; lane++)
          for (uint32_t u = warp; u < 2 * cols; u += 8) {
            const uint32_t c = u >> 1, x0 = (u & 1) ? half : 0u, x1 = (u & 1) ? run : half;
            const u32*     src = d.q + ((size_t)c * rows + j0) * W;
            for (uint32_t x = x0 + lane; x < x1; x += 32)
              s_tile[c * cs + x] = src[x];
          }
    }
    if (j0 + rt + ack_rows > rows) {
      for (uint32_t idx = 0; idx < d.q_ack * W; idx++) {
        const uint32_t r = idx / W, w = idx - r * W, j = rows - 1 - r / 4;
        if (j >= j0 && j < j0 + rt)
          s_tile[ul_col(d.ack_cols, r) * cs + (j - j0) * W + w] = 0;
      }
    }
    const uint32_t n_plain = j0 + rt + ri_rows <= rows ? rt : (j0 + ri_rows >= rows ? 0u : rows - ri_rows - j0);
    const size_t   base = (size_t)j0 * rw;
    {
      u32*           dst = d.g + base;
      const uint32_t rpp = d.rpp;
      for (uint32_t tid = 0; tid < 256; tid++)
        if (tid < rpp * rw) {
          const uint32_t r0 = (tid * d.inv_rw) >> 16, t = tid - r0 * rw, c = t / W, w = t - c * W;
          const u32*     sp = s_tile.data() + c * cs + r0 * W + w;
          u32*           gp = dst + tid;
          for (uint32_t jr = r0; jr < n_plain; jr += rpp, sp += rpp * W, gp += rpp * rw)
            *gp = *sp;
        }
      if (n_plain && (base < cqi_words || (j0 == 0 && d.clobber > 0))) {
        const uint32_t n_fix = (uint32_t)std::min((size_t)n_plain * rw, std::max((size_t)cqi_words, base + 1) - base);
        for (uint32_t idx = 0; idx < n_fix; idx++) {
          const uint32_t sym = idx / W, w = idx - sym * W;
          const uint32_t jr = (sym * inv_cols) >> 16, c = sym - jr * cols;
          u32            v  = s_tile[c * cs + jr * W + w];
          if (base + idx == 0 && d.clobber > 0) {
            v      = (v & 0xffff0000u) | (uint16_t)qe[d.clobber];
            dst[0] = v;
          }
          if (base + idx < cqi_words)
            cqi[base + idx] = v;
        }
      }
    }
    for (uint32_t idx = n_plain * rw; idx < rt * rw; idx++) {
      const uint32_t sym = idx / W, w = idx - sym * W;
      const uint32_t jr = (sym * inv_cols) >> 16, c = sym - jr * cols;
      const uint32_t m    = rows - 1 - (j0 + jr);
      const uint32_t n_ri = ul_row_count(m, d.q_ri);
      if (ul_holds(n_ri, d.ri_cols, c))
        continue;
      const uint32_t o = ((j0 + jr) * cols + c - ul_ri_before(m, n_ri, d.q_ri, d.ri_cols, c)) * W + w;
      u32            v = s_tile[c * cs + jr * W + w];
      if (o == 0 && d.clobber > (int32_t)(((size_t)c * rows + j0 + jr) * 2 * W))
        v = (v & 0xffff0000u) | (uint16_t)qe[d.clobber];
      d.g[o] = v;
      if (o < cqi_words)
        cqi[o] = v;
    }
  }
  if (bx == 0 && d.uci) {
    constexpr uint32_t Qm = 2 * W;
    for (uint32_t idx = 0; idx < (d.q_ack + d.q_ri) * Qm; idx++) {
      const uint32_t r0 = idx / Qm, k = idx - r0 * Qm;
      const bool     ri = r0 >= d.q_ack;
      d.uci[idx] = qe[ul_uci_element(ri ? d.ri_cols : d.ack_cols, ri ? r0 - d.q_ack : r0, rows, Qm, k)];
    }
  }
}

extern "C" int emul_ulsch(const int16_t* q_bits, uint32_t Qm, uint32_t H, uint32_t nsym, uint32_t qa, uint32_t qr, uint32_t qc, int16_t* g_bits,
                          int16_t* uci, uint32_t grid_x)
{
  UlschDev d;
  d.q = (const u32*)q_bits;
  d.g = (u32*)g_bits;
  d.uci = uci;
  d.W = Qm / 2;
  d.cols = nsym;
  d.rows = H / nsym;
  d.q_ack = qa; d.q_ri = qr; d.q_cqi = qc;
  ul_finish_descriptor(d, nsym);
  for (uint32_t bx = 0; bx < grid_x; bx++) {
    switch (d.W) {
      case 1: emul_cta<1>(d, bx, grid_x); break;
      case 2: emul_cta<2>(d, bx, grid_x); break;
      case 3: emul_cta<3>(d, bx, grid_x); break;
      default: emul_cta<4>(d, bx, grid_x); break;
    }
  }
  return 0;
}
