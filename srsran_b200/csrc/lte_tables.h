// Host-side LTE tables for the B200 turbo-decode engine (init-time work).
// Product code: does not include or link anything under oracle/.
#pragma once
#include <cstdint>
#include <vector>

namespace b200 {

constexpr int      kNofCbSizes   = 188;   // TS 36.212 Table 5.1.3-3
constexpr uint32_t kMaxK         = 6144;
constexpr uint32_t kSoftbufElems = 18600; // reference SOFTBUFFER_SIZE (softbuffer.h:50)
constexpr uint32_t kSbPad        = 32;    // plane padding of the lane layout (rm_turbo.c:272, turbodecoder_iter.h:93-94)

int      cb_size(uint32_t idx);          // srslte_cbsegm_cbsize   (cbsegm.c:132-139)
int      cb_index(uint32_t len);         // srslte_cbsegm_cbindex  (cbsegm.c:114-125): smallest K >= len
bool     cb_size_valid(uint32_t K);      // srslte_cbsegm_cbsize_isvalid (cbsegm.c:147-155)
uint32_t auto_lanes16(uint32_t K);       // srslte_tdec_autoimp_get_subblocks      (turbodecoder.c:381-393)
uint32_t auto_lanes8(uint32_t K);        // srslte_tdec_autoimp_get_subblocks_8bit (turbodecoder.c:410-424)

struct CbSegm { // mirrors srslte_cbsegm_t (cbsegm.h:32-42)
  uint32_t F, C, K1, K2, K1_idx, K2_idx, C1, C2, tbs;
};
int cb_segm(CbSegm* s, uint32_t tbs); // srslte_cbsegm (cbsegm.c:48-103)

// QPP interleaver pi(i) = (f1 i + f2 i^2) mod K in the index space of an N-lane layout (N<=1: natural order).
// fwd/rev as srslte_tc_interl_LTE_gen_interl (tc_interl_lte.c:69-109) would fill them.
void qpp_tables(uint32_t K, uint32_t lanes, uint16_t* fwd, uint16_t* rev);

// Rate de-matching: positions (in the chosen layout) of the 3K+12 non-NULL circular-buffer entries in
// circular-buffer order starting at slot 0, plus the rank of the first transmitted entry for each rv.
// The reference's table for one rv (rm_turbo.c:177-251 + :263-277) is base[(i + start[rv]) % (3K+12)].
struct RmTable {
  std::vector<uint16_t> base;
  uint32_t              start[4];
};
void rm_table(uint32_t K, uint32_t lanes, RmTable* t);

// index helpers for the lane layout (tc_interl_lte.c:66-67)
inline uint32_t to_lane(uint32_t n, uint32_t K, uint32_t N) { return (n % (K / N)) * N + n / (K / N); }
inline uint32_t from_lane(uint32_t j, uint32_t K, uint32_t N) { return (j % N) * (K / N) + j / N; }

} // namespace b200
