// Host-side LTE tables (see lte_tables.h).  Written from TS 36.212; behaviour checked against the
// reference through the test-suite (tests/test_host_tables.py).
#include "lte_tables.h"
#include "lte_qpp_table.h"

namespace b200 {

int cb_size(uint32_t idx) { return idx < (uint32_t)kNofCbSizes ? (int)lte_qpp_table[idx].K : -1; }

int cb_index(uint32_t len)
{
  // table is sorted: binary search for the first K >= len
  int lo = 0, hi = kNofCbSizes;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (lte_qpp_table[mid].K < len)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo == kNofCbSizes ? -1 : lo;
}

bool cb_size_valid(uint32_t K)
{
  int i = cb_index(K);
  return i >= 0 && lte_qpp_table[i].K == K;
}

uint32_t auto_lanes16(uint32_t K)
{
  if (K % 16 == 0 && K > 800)
    return 16;
  if (K % 8 == 0 && K > 400)
    return 8;
  return 0;
}

uint32_t auto_lanes8(uint32_t K)
{
  if (K % 32 == 0 && K > 2048)
    return 32;
  return auto_lanes16(K);
}

int cb_segm(CbSegm* s, uint32_t tbs)
{
  *s = CbSegm{};
  if (tbs == 0)
    return 0;
  const uint32_t B = tbs + 24;
  uint32_t       Bp;
  s->tbs = tbs;
  if (B <= kMaxK) {
    s->C = 1;
    Bp   = B;
  } else {
    // the reference evaluates ceil(B / 6120) in single precision (cbsegm.c:65); keep that
    const float q = (float)B / (float)(kMaxK - 24);
    uint32_t    c = (uint32_t)q;
    if ((float)c < q)
      c++;
    s->C = c;
    Bp   = B + 24 * c;
  }
  const int i1 = cb_index((Bp - 1) / s->C + 1);
  if (i1 < 0)
    return -1;
  s->K1     = lte_qpp_table[i1].K;
  s->K1_idx = (uint32_t)i1;
  if (s->C == 1) {
    s->C1 = 1;
  } else {
    s->K2_idx = i1 > 0 ? (uint32_t)i1 - 1 : 0;
    s->K2     = lte_qpp_table[s->K2_idx].K;
    s->C2     = (s->C * s->K1 - Bp) / (s->K1 - s->K2);
    s->C1     = s->C - s->C2;
  }
  s->F = s->C1 * s->K1 + s->C2 * s->K2 - Bp;
  return 0;
}

void qpp_tables(uint32_t K, uint32_t lanes, uint16_t* fwd, uint16_t* rev)
{
  const int      ci = cb_index(K);
  const uint32_t f1 = lte_qpp_table[ci].f1, f2 = lte_qpp_table[ci].f2;
  // pi(i+1) = pi(i) + g(i), g(i+1) = g(i) + 2 f2 (mod K): no 64-bit products needed
  std::vector<uint16_t> nat(K);
  uint32_t              pi = 0, g = (f1 + f2) % K;
  const uint32_t        dg = (2 * f2) % K;
  for (uint32_t i = 0; i < K; i++) {
    nat[i] = (uint16_t)pi;
    pi     = (pi + g) % K;
    g      = (g + dg) % K;
  }
  for (uint32_t j = 0; j < K; j++) {
    const uint32_t n = lanes > 1 ? from_lane(j, K, lanes) : j;
    const uint32_t t = lanes > 1 ? to_lane(nat[n], K, lanes) : nat[n];
    fwd[j]           = (uint16_t)t;
    rev[t]           = (uint16_t)j;
  }
}

void rm_table(uint32_t K, uint32_t lanes, RmTable* t)
{
  static const uint8_t P[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
  const uint32_t D = K + 4, R = (D + 31) / 32, Kp = 32 * R, ND = Kp - D, Ncb = 3 * Kp;
  // position in the decoder's buffer of bit k of stream s
  auto place = [&](uint32_t k, uint32_t s) -> uint16_t {
    if (lanes == 0)
      return (uint16_t)(3 * k + s);
    if (k < K)
      return (uint16_t)(s * (K + kSbPad) + to_lane(k, K, lanes));
    return (uint16_t)(3 * (K + kSbPad) + 3 * (k - K) + s);
  };
  t->base.clear();
  t->base.reserve(3 * K + 12);
  uint32_t k0[4];
  for (uint32_t rv = 0; rv < 4; rv++) {
    k0[rv]       = R * (2 * ((Ncb + 8 * R - 1) / (8 * R)) * rv + 2);
    t->start[rv] = 0;
  }
  // walk the circular buffer w = [v0 | v1,v2 interlaced] once; slot j of v_s is column j / R, row j % R
  for (uint32_t slot = 0; slot < Ncb; slot++) {
    for (uint32_t rv = 0; rv < 4; rv++)
      if (slot == k0[rv] % Ncb)
        t->start[rv] = (uint32_t)t->base.size();
    uint32_t s, j;
    if (slot < Kp) {
      s = 0;
      j = slot;
    } else {
      s = 1 + ((slot - Kp) & 1);
      j = (slot - Kp) >> 1;
    }
    uint32_t y = (j % R) * 32 + P[j / R];
    if (s == 2)
      y = (y + 1) % Kp; // third stream uses the permutation shifted by one (TS 36.212 5.1.4.1.1)
    if (y >= ND)
      t->base.push_back(place(y - ND, s));
  }
}

} // namespace b200
