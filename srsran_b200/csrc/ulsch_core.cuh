// ulsch_core.cuh -- index arithmetic of k_ulsch_deinterleave (kernels.cuh), __host__ __device__ so that the CPU suite can
// run the kernel's loops with g++ (tests/host_emul/emul_ulsch.cpp).
//
// The channel interleaver matrix of TS 36.212 5.2.2.8 has `cols` = N_pusch_symbs columns and `rows` = H' / cols rows of
// Qm-LLR symbols; q_bits hold it column by column, g_bits row by row with the RI symbols left out (ulsch_interleave_gen,
// phch/sch.c:658-679).  Coded ACK / RI symbol r sits in row rows-1-r/4, column set[(3r) % 4]
// (uci_ulsch_interleave_{ack,ri}_gen, phch/uci.c:551-605): the bottom rows hold up to four of each.
#pragma once
#include <stdint.h>

#include "arith.cuh"

namespace b200 {

struct UlschDev {
  const u32* q;      // H' * W words (W = Qm / 2: every Qm is even, the unit is one 32-bit word = two LLRs)
  u32*       g;      // (H' - Q'_ri) * W words
  int16_t*   uci;    // [Q'_ack * Qm | Q'_ri * Qm | Q'_cqi * Qm] LLRs for the host's UCI decoders, or nullptr
  uint32_t   W, rows, cols;
  uint32_t   q_ack, q_ri, q_cqi;
  uint32_t   ack_cols, ri_cols; // the four columns of each set, one per byte, indexed by (3r) % 4
  uint32_t   inv_cols, inv_rw, rpp; // host-computed: 65536 / cols + 1, 65536 / (cols * W) + 1 (exact quotients by multiply
                                // and shift for the operand ranges of a tile), 256 / (cols * W) rows per store pass
  int32_t    clobber;           // element of q_bits that the reference leaves in g_bits[0] (its table sends every RI
                                // position to index 0 and the last store wins, sch.c:670-671 + vector.c:136-141), or -1
};
constexpr int kUlRows    = 128; // rows of the matrix per tile
constexpr int kUlMaxCols = 14;
constexpr int kUlMaxW    = 4;

// column sets of uci.c:558-559, 586-587 for N_pusch_symbs > 10 (normal CP) and <= 10 (extended CP)
constexpr uint32_t kUlAckNorm = 2u | 3u << 8 | 8u << 16 | 9u << 24, kUlAckExt = 1u | 2u << 8 | 6u << 16 | 7u << 24;
constexpr uint32_t kUlRiNorm = 1u | 4u << 8 | 7u << 16 | 10u << 24, kUlRiExt = 0u | 3u << 8 | 5u << 16 | 8u << 24;

B200_HD uint32_t ul_col(uint32_t colset, uint32_t r) { return (colset >> (8 * ((3 * r) & 3))) & 0xffu; }
// coded symbols of one UCI kind (q of them in all) in the row that is m rows above the bottom one
B200_HD uint32_t ul_row_count(uint32_t m, uint32_t q) { return 4 * (uint64_t)m >= q ? 0u : (q - 4 * m < 4u ? q - 4 * m : 4u); }
// does column c of a row with n_row such symbols hold one?
B200_HD bool ul_holds(uint32_t n_row, uint32_t colset, uint32_t c)
{
  bool h = false;
  for (uint32_t t = 0; t < 4; t++)
    h |= t < n_row && ul_col(colset, t) == c;
  return h;
}
// RI symbols that precede (row j, column c) in row-major order; j's row is m = rows-1-j above the bottom, holds n_ri
B200_HD uint32_t ul_ri_before(uint32_t m, uint32_t n_ri, uint32_t q_ri, uint32_t ri_cols, uint32_t c)
{
  const uint64_t upto = 4 * ((uint64_t)m + 1);
  uint32_t       before = upto >= q_ri ? 0u : q_ri - (uint32_t)upto; // rows above this one
  for (uint32_t t = 0; t < 4; t++)
    before += t < n_ri && ul_col(ri_cols, t) < c;
  return before;
}
// element index in q_bits of LLR k of coded ACK / RI symbol r
B200_HD size_t ul_uci_element(uint32_t colset, uint32_t r, uint32_t rows, uint32_t Qm, uint32_t k)
{
  return ((size_t)ul_col(colset, r) * rows + (rows - 1 - r / 4)) * Qm + k;
}

// fills the derived fields of a descriptor whose W, rows, cols, q_ack, q_ri are set (host side)
inline void ul_finish_descriptor(UlschDev& d, uint32_t N_pusch_symbs)
{
  d.ack_cols = N_pusch_symbs > 10 ? kUlAckNorm : kUlAckExt;
  d.ri_cols  = N_pusch_symbs > 10 ? kUlRiNorm : kUlRiExt;
  d.inv_cols = 65536u / d.cols + 1;
  d.inv_rw   = 65536u / (d.cols * d.W) + 1;
  d.rpp      = 256u / (d.cols * d.W);
  d.clobber  = -1;
  for (uint32_t r = 0; r < d.q_ri; r++) {
    const int32_t e = (int32_t)ul_uci_element(d.ri_cols, r, d.rows, 2 * d.W, 2 * d.W - 1);
    d.clobber       = e > d.clobber ? e : d.clobber;
  }
}

} // namespace b200
