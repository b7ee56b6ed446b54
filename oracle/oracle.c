/*
 * oracle.c -- CPU restatement of the reference's LTE turbo-decode hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may build, load or call it.  The product library
 * (srsran_b200/csrc -> libsrslte_fec_b200.so) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function below is checked element-for-element against the
 * UNMODIFIED reference compiled from /root/reference by oracle/build_ref.sh
 * (tests/test_oracle_vs_ref.py; golden vectors produced by that binary are committed under
 * tests/golden/ by tests/golden/make_golden.py), and against the reference's own known-answer
 * material (crc_test.h:36-39, turbodecoder_test.h:70-125 via the golden files).
 *
 * Everything is written scalar, one sub-block lane at a time, from the numerical contract in
 * SURVEY.md 8a-3; each function cites the reference lines it restates (paths relative to
 * /root/reference/lib).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lte_qpp_table.h"

#define ORC_MAX_K 6144
#define ORC_SOFTBUFFER_SIZE 18600 /* include/srslte/phy/fec/softbuffer.h:50 */
#define ORC_WIN_OVERLAP 40        /* include/srslte/phy/fec/turbodecoder_win.h:54 */

/* ============================================================ code block sizes / segmentation */

/* src/phy/fec/cbsegm.c:29-39 (table), :114-125 (lookup of the smallest K >= len) */
int orc_cbsize(uint32_t idx) { return idx < LTE_NOF_CB_SIZES ? (int)lte_qpp_table[idx].K : -1; }
int orc_cbindex(uint32_t len)
{
  for (int j = 0; j < LTE_NOF_CB_SIZES; j++)
    if (lte_qpp_table[j].K >= len)
      return j;
  return -1;
}

/* src/phy/fec/cbsegm.c:48-103.  out = {F, C, K1, K2, K1_idx, K2_idx, C1, C2, tbs}.
 * C = ceil(B/6120) is evaluated in single-precision float in the reference (:65). */
int orc_cbsegm(uint32_t tbs, uint32_t out[9])
{
  memset(out, 0, 9 * sizeof(uint32_t));
  if (tbs == 0)
    return 0;
  uint32_t B = tbs + 24, C, Bp;
  if (B <= ORC_MAX_K) {
    C  = 1;
    Bp = B;
  } else {
    float q = (float)B / (float)(ORC_MAX_K - 24);
    C       = (uint32_t)q;
    if ((float)C < q)
      C++;
    Bp = B + 24 * C;
  }
  int idx1 = orc_cbindex((Bp - 1) / C + 1);
  if (idx1 < 0)
    return -1;
  uint32_t K1 = lte_qpp_table[idx1].K, K2 = 0, K2i = 0, C1 = 1, C2 = 0;
  if (C > 1) {
    /* for idx1 == 0 the reference reuses K1 for K2 (division by zero follows); unreachable for C > 1 */
    K2i = idx1 > 0 ? (uint32_t)idx1 - 1 : 0;
    K2  = lte_qpp_table[K2i].K;
    C2  = (C * K1 - Bp) / (K1 - K2);
    C1  = C - C2;
  }
  out[0] = C1 * K1 + C2 * K2 - Bp;
  out[1] = C;
  out[2] = K1;
  out[3] = K2;
  out[4] = (uint32_t)idx1;
  out[5] = K2i;
  out[6] = C1;
  out[7] = C2;
  out[8] = tbs;
  return 0;
}

/* src/phy/fec/turbodecoder.c:381-424: number of sub-block lanes of the AUTO decoder (AVX2 build) */
uint32_t orc_subblocks16(uint32_t K) { return (K % 16 == 0 && K > 800) ? 16 : (K % 8 == 0 && K > 400) ? 8 : 0; }
uint32_t orc_subblocks8(uint32_t K)
{
  return (K % 32 == 0 && K > 2048) ? 32 : (K % 16 == 0 && K > 800) ? 16 : (K % 8 == 0 && K > 400) ? 8 : 0;
}

/* ============================================================================ QPP interleaver */

/* natural index n -> lane layout: element (lane n / W, step n % W) lives at step*N + lane
 * (src/phy/fec/tc_interl_lte.c:66-67 "deinter"/"inter") */
static inline uint32_t to_lane(uint32_t n, uint32_t K, uint32_t N) { return (n % (K / N)) * N + n / (K / N); }
static inline uint32_t from_lane(uint32_t j, uint32_t K, uint32_t N) { return (j % N) * (K / N) + j / N; }

/* src/phy/fec/tc_interl_lte.c:69-109.  nsb <= 1: natural order. */
int orc_qpp(uint32_t K, uint32_t nsb, uint16_t* fwd, uint16_t* rev)
{
  int ci = orc_cbindex(K);
  if (ci < 0 || lte_qpp_table[ci].K != K)
    return -1;
  uint64_t  f1 = lte_qpp_table[ci].f1, f2 = lte_qpp_table[ci].f2;
  uint16_t* f = malloc(K * sizeof(uint16_t));
  uint16_t* r = malloc(K * sizeof(uint16_t));
  for (uint64_t i = 0; i < K; i++) {
    uint32_t j = (uint32_t)((f1 * i + f2 * i * i) % K);
    f[i]       = (uint16_t)j;
    r[j]       = (uint16_t)i;
  }
  for (uint32_t i = 0; i < K; i++) {
    if (nsb > 1) {
      fwd[i] = (uint16_t)to_lane(f[from_lane(i, K, nsb)], K, nsb);
      rev[i] = (uint16_t)to_lane(r[from_lane(i, K, nsb)], K, nsb);
    } else {
      fwd[i] = f[i];
      rev[i] = r[i];
    }
  }
  free(f);
  free(r);
  return 0;
}

/* ============================================================ rate de-matching (TS 36.212 5.1.4.1) */

static const uint8_t rm_col_perm[32] = {0, 16, 8,  24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                        1, 17, 9,  25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

/* Restates srslte_rm_turbo_gentable_receive (src/phy/fec/rm_turbo.c:177-251) from the standard:
 * table[i] = 3*k + s  <=>  the i-th transmitted bit (starting at k0(rv), NULLs skipped) is bit k of
 * stream s (s = 0 systematic, 1 parity0, 2 parity1; k < K+4, the 4 extra columns carry the 12 tail bits in
 * the order the turbo coder emits them).  nsb > 0 additionally applies interleave_table_sb (:263-277). */
int orc_rm_table(uint32_t K, uint32_t rv, uint32_t nsb, uint16_t* table)
{
  uint32_t D = K + 4, R = (D - 1) / 32 + 1, Kp = 32 * R, ND = Kp - D, Ncb = 3 * Kp;
  uint32_t k0 = R * (2 * ((Ncb + 8 * R - 1) / (8 * R)) * rv + 2);
  /* w[j] = codeword position held by circular-buffer slot j, or -1 for a <NULL> */
  int32_t* w = malloc(Ncb * sizeof(int32_t));
  for (uint32_t col = 0; col < 32; col++) {
    for (uint32_t row = 0; row < R; row++) {
      uint32_t j = col * R + row;
      int32_t  y = (int32_t)(row * 32 + rm_col_perm[col]) - (int32_t)ND; /* streams 0 and 1 */
      w[j]          = y >= 0 ? 3 * y + 0 : -1;
      w[Kp + 2 * j] = y >= 0 ? 3 * y + 1 : -1;
      int32_t y2    = (int32_t)((rm_col_perm[col] + 32 * row + 1) % Kp) - (int32_t)ND; /* stream 2: shifted perm */
      w[Kp + 2 * j + 1] = y2 >= 0 ? 3 * y2 + 2 : -1;
    }
  }
  uint32_t n = 0, len = 3 * K + 12;
  for (uint32_t j = 0; n < len; j++) {
    int32_t v = w[(k0 + j) % Ncb];
    if (v < 0)
      continue;
    if (nsb > 0) {
      if ((uint32_t)v < 3 * K)
        v = (v % 3) * (int32_t)(K + 32) + (int32_t)to_lane((uint32_t)v / 3, K, nsb);
      else
        v = v - 3 * (int32_t)K + 3 * (int32_t)(K + 32);
    }
    table[n++] = (uint16_t)v;
  }
  free(w);
  return 0;
}

/* src/phy/fec/rm_turbo.c:397-454 / :456-493: out[tab[i mod (3K+12)]] += in[i], wrapping.
 * layout: 0 = standard (3k+s), N = lane layout with N lanes. */
int orc_rm_rx16(const int16_t* in, int16_t* out, uint32_t E, uint32_t K, uint32_t rv, uint32_t nsb)
{
  uint32_t  len = 3 * K + 12;
  uint16_t* tab = malloc(len * sizeof(uint16_t));
  orc_rm_table(K, rv, nsb, tab);
  for (uint32_t i = 0; i < E; i++) {
    uint16_t p = tab[i % len];
    out[p]     = (int16_t)(uint16_t)((uint16_t)out[p] + (uint16_t)in[i]);
  }
  free(tab);
  return 0;
}
int orc_rm_rx8(const int8_t* in, int8_t* out, uint32_t E, uint32_t K, uint32_t rv, uint32_t nsb)
{
  uint32_t  len = 3 * K + 12;
  uint16_t* tab = malloc(len * sizeof(uint16_t));
  orc_rm_table(K, rv, nsb, tab);
  for (uint32_t i = 0; i < E; i++) {
    uint16_t p = tab[i % len];
    out[p]     = (int8_t)(uint8_t)((uint8_t)out[p] + (uint8_t)in[i]);
  }
  free(tab);
  return 0;
}

/* ======================================================================================= CRC */

/* src/phy/fec/crc.c:30-46,143-157 + include/srslte/phy/fec/crc.h:56-70: MSB-first, init 0, no reflection,
 * no final xor.  Restated bit-serially (the table method is the same polynomial division). */
uint32_t orc_crc_bytes(uint32_t poly, int order, const uint8_t* data, uint32_t nbytes)
{
  uint64_t reg = 0, top = 1ull << order, mask = top - 1;
  for (uint32_t i = 0; i < nbytes; i++) {
    for (int b = 7; b >= 0; b--) {
      reg = (reg << 1) | ((data[i] >> b) & 1u);
      if (reg & top)
        reg ^= ((uint64_t)poly | top);
    }
  }
  /* flush `order` zero bits (the table form multiplies the message by x^order) */
  for (int b = 0; b < order; b++) {
    reg <<= 1;
    if (reg & top)
      reg ^= ((uint64_t)poly | top);
  }
  return (uint32_t)(reg & mask);
}
/* src/phy/fec/crc.c:102-140: same CRC over an array of unpacked bits (one bit per byte) */
uint32_t orc_crc_bits(uint32_t poly, int order, const uint8_t* bits, uint32_t nbits)
{
  uint64_t reg = 0, top = 1ull << order, mask = top - 1;
  for (uint32_t i = 0; i < nbits + (uint32_t)order; i++) {
    reg = (reg << 1) | (i < nbits ? (bits[i] & 1u) : 0u);
    if (reg & top)
      reg ^= ((uint64_t)poly | top);
  }
  return (uint32_t)(reg & mask);
}

/* ============================================================================== arithmetic */

typedef struct {
  int      bits;        /* 8 or 16 */
  uint32_t N;           /* lanes; 0 = generic (un-windowed) decoder */
  int32_t  inf;         /* 10000 / 0   (turbodecoder_win.h:56,151,186,287) */
  int      norm_max;    /* 8-bit: subtract the max every step (win.h:180-181,289-290) */
  int      out_shift;   /* 8-bit: output >>= 1 (win.h:184,293) */
} orc_cfg_t;

static inline int32_t sat(const orc_cfg_t* c, int32_t v)
{
  int32_t hi = c->bits == 16 ? 32767 : 127, lo = -hi - 1;
  return v > hi ? hi : (v < lo ? lo : v);
}
static inline int32_t wrap(const orc_cfg_t* c, int32_t v)
{
  return c->bits == 16 ? (int32_t)(int16_t)(uint16_t)v : (int32_t)(int8_t)(uint8_t)v;
}
/* tail-trellis adder (turbodecoder_win.h:470-478): 16-bit wraps; 8-bit saturates only upwards */
static inline int32_t tail_add(const orc_cfg_t* c, int32_t a, int32_t b)
{
  if (c->bits == 16)
    return wrap(c, a + b);
  int32_t z = (int32_t)(int16_t)(a + b);
  return z > 127 ? 127 : (int32_t)(int8_t)(uint8_t)z;
}

/* backward branch metrics (turbodecoder_win.h:643-659 == turbodecoder_gen.c:80-96) */
#define BWD_STEP(ADD)                                                                                                  \
  do {                                                                                                                 \
    mb[0] = ADD(o[4], xy); mb[1] = o[4];        mb[2] = ADD(o[5], y);  mb[3] = ADD(o[5], x);                            \
    mb[4] = ADD(o[6], x);  mb[5] = ADD(o[6], y); mb[6] = o[7];         mb[7] = ADD(o[7], xy);                           \
    nw[0] = o[0];          nw[1] = ADD(o[0], xy); nw[2] = ADD(o[1], x); nw[3] = ADD(o[1], y);                           \
    nw[4] = ADD(o[2], y);  nw[5] = ADD(o[2], x);  nw[6] = ADD(o[3], xy); nw[7] = o[3];                                  \
  } while (0)
/* forward branch metrics (turbodecoder_win.h:769-785 == turbodecoder_gen.c:148-164) */
#define FWD_STEP(ADD)                                                                                                  \
  do {                                                                                                                 \
    mb[0] = o[0];          mb[1] = ADD(o[3], y);  mb[2] = ADD(o[4], y); mb[3] = o[7];                                   \
    mb[4] = o[1];          mb[5] = ADD(o[2], y);  mb[6] = ADD(o[5], y); mb[7] = o[6];                                   \
    nw[0] = ADD(o[1], xy); nw[1] = ADD(o[2], x);  nw[2] = ADD(o[5], x); nw[3] = ADD(o[6], xy);                          \
    nw[4] = ADD(o[0], xy); nw[5] = ADD(o[3], x);  nw[6] = ADD(o[4], x); nw[7] = ADD(o[7], xy);                          \
  } while (0)

/* turbodecoder_win.h:480-498 */
static void win_normalize(const orc_cfg_t* c, uint32_t k, int32_t o[8])
{
  if (c->norm_max) {
    if (k == 0)
      return;
    int32_t m = o[0];
    for (int i = 1; i < 8; i++)
      m = o[i] > m ? o[i] : m;
    for (int i = 0; i < 8; i++)
      o[i] = sat(c, o[i] - m);
  } else {
    if (k % 2 != 0 || k == 0)
      return;
    for (int i = 1; i < 8; i++)
      o[i] = sat(c, o[i] - o[0]);
    o[0] = 0;
  }
}

/* Windowed max-log-MAP: tdec_win*_dec = beta + alpha (turbodecoder_win.h:551-681, 684-832, 860-868).
 * in/app/par/out are in lane layout (element (lane d, step p) at p*N+d); in[K..K+2], par[K..K+2] hold the tail.
 * app may be NULL.  All values are carried in int32 but always inside the range of the configured width. */
#define SADD(a, b) sat(c, (a) + (b))
static void map_win(const orc_cfg_t* c, const int32_t* in, const int32_t* app, const int32_t* par, int32_t* out,
                    uint32_t K)
{
  const uint32_t N = c->N, W = K / N;
  int32_t*       beta = malloc((size_t)(W + 1) * N * 8 * sizeof(int32_t)); /* beta[p][d][s] */
  int32_t(*warm)[8]   = malloc(N * sizeof(*warm));
  int32_t o[8], mb[8], nw[8];

  /* ---- beta, pass 0: 40 warm-up steps on each lane's own first steps (win.h:622-630) */
  for (uint32_t d = 0; d < N; d++) {
    for (int i = 0; i < 8; i++)
      o[i] = -c->inf;
    for (int k = ORC_WIN_OVERLAP - 1; k >= 0; k--) {
      int32_t x = in[k * N + d], y = par[k * N + d];
      if (app)
        x = SADD(app[k * N + d], x);
      int32_t xy = SADD(x, y);
      BWD_STEP(SADD);
      for (int i = 0; i < 8; i++)
        o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
      win_normalize(c, (uint32_t)k, o);
    }
    memcpy(warm[d], o, sizeof(o));
  }
  /* ---- beta, pass 1 (win.h:577-620, 632-679) */
  for (uint32_t d = 0; d < N; d++) {
    if (d + 1 < N) {
      memcpy(o, warm[d + 1], sizeof(o)); /* lane d starts from what lane d+1 estimated at its step 0 */
    } else {
      /* last lane: 3 tail steps from the known zero state, no a-priori (win.h:500-548) */
      o[0] = 0;
      for (int i = 1; i < 8; i++)
        o[i] = -c->inf;
      for (uint32_t k = K + 2; k >= K; k--) {
        int32_t x = in[k], y = par[k];
#define TADD(a, b) tail_add(c, (a), (b))
        int32_t xy = TADD(x, y);
        BWD_STEP(TADD);
        for (int i = 0; i < 8; i++)
          o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
      }
    }
    memcpy(&beta[((size_t)W * N + d) * 8], o, sizeof(o));
    for (int k = (int)W - 1; k >= 0; k--) {
      int32_t x = in[k * N + d], y = par[k * N + d];
      if (app)
        x = SADD(app[k * N + d], x);
      int32_t xy = SADD(x, y);
      BWD_STEP(SADD);
      for (int i = 0; i < 8; i++)
        o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
      memcpy(&beta[((size_t)k * N + d) * 8], o, sizeof(o)); /* stored before normalising (win.h:666-678) */
      win_normalize(c, (uint32_t)k, o);
    }
  }
  /* ---- alpha, pass 0: warm-up on each lane's own last 40 steps (win.h:747-756) */
  for (uint32_t d = 0; d < N; d++) {
    for (int i = 0; i < 8; i++)
      o[i] = -c->inf;
    for (uint32_t k = 0; k < ORC_WIN_OVERLAP; k++) {
      uint32_t p = W - ORC_WIN_OVERLAP + k;
      int32_t  x = in[p * N + d], y = par[p * N + d];
      if (app)
        x = SADD(app[p * N + d], x);
      int32_t xy = SADD(x, y);
      FWD_STEP(SADD);
      for (int i = 0; i < 8; i++)
        o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
      win_normalize(c, k, o);
    }
    memcpy(warm[d], o, sizeof(o));
  }
  /* ---- alpha, pass 1 with LLR output (win.h:715-746, 758-830) */
  for (uint32_t d = 0; d < N; d++) {
    if (d > 0) {
      memcpy(o, warm[d - 1], sizeof(o));
    } else {
      o[0] = 0;
      for (int i = 1; i < 8; i++)
        o[i] = -c->inf;
    }
    for (uint32_t k = 0; k < W; k++) {
      int32_t x = in[k * N + d], y = par[k * N + d];
      if (app)
        x = SADD(app[k * N + d], x);
      int32_t xy = SADD(x, y);
      FWD_STEP(SADD);
      const int32_t* b  = &beta[((size_t)(k + 1) * N + d) * 8];
      int32_t        m1 = SADD(b[0], nw[0]), m0 = SADD(b[0], mb[0]);
      for (int i = 1; i < 8; i++) {
        int32_t t1 = SADD(b[i], nw[i]), t0 = SADD(b[i], mb[i]);
        m1 = t1 > m1 ? t1 : m1;
        m0 = t0 > m0 ? t0 : m0;
      }
      int32_t l = sat(c, m1 - m0);
      if (c->out_shift)
        l >>= c->out_shift; /* arithmetic shift per element (win.h:188-193,811-813) */
      out[k * N + d] = l;
      for (int i = 0; i < 8; i++)
        o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
      win_normalize(c, k, o);
    }
  }
  free(beta);
  free(warm);
}

/* Generic decoder: tdec_gen_dec (src/phy/fec/turbodecoder_gen.c:58-112, 114-198, 226-236).
 * Natural order, un-windowed, every add/sub WRAPS in int16. */
#define WADD(a, b) ((int32_t)(int16_t)(uint16_t)((a) + (b)))
static void map_gen(const int32_t* in, const int32_t* app, const int32_t* par, int32_t* out, uint32_t K)
{
  const int32_t INF  = 10000;
  int32_t*      beta = malloc((size_t)(K + 4) * 8 * sizeof(int32_t));
  int32_t       o[8], mb[8], nw[8];
  o[0] = 0;
  for (int i = 1; i < 8; i++)
    o[i] = -INF;
  memcpy(&beta[(size_t)(K + 3) * 8], o, sizeof(o));
  for (int k = (int)K + 2; k >= 0; k--) {
    int32_t x = in[k];
    if (app && (uint32_t)k < K)
      x = WADD(x, app[k]);
    int32_t y = par[k], xy = WADD(x, y);
    BWD_STEP(WADD);
    for (int i = 0; i < 8; i++)
      o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
    memcpy(&beta[(size_t)k * 8], o, sizeof(o));
    if (k % 4 == 0 && (uint32_t)k < K) {
      for (int i = 1; i < 8; i++)
        o[i] = WADD(o[i], -o[0]);
      o[0] = 0;
    }
  }
  o[0] = 0;
  for (int i = 1; i < 8; i++)
    o[i] = -INF;
  for (uint32_t k = 1; k <= K; k++) {
    int32_t x = in[k - 1];
    if (app)
      x = WADD(x, app[k - 1]);
    int32_t y = par[k - 1], xy = WADD(x, y);
    FWD_STEP(WADD);
    const int32_t* b  = &beta[(size_t)k * 8];
    int32_t        m1 = WADD(nw[0], b[0]), m0 = WADD(mb[0], b[0]);
    for (int i = 1; i < 8; i++) {
      int32_t t1 = WADD(nw[i], b[i]), t0 = WADD(mb[i], b[i]);
      m1 = t1 > m1 ? t1 : m1;
      m0 = t0 > m0 ? t0 : m0;
    }
    for (int i = 0; i < 8; i++)
      o[i] = mb[i] > nw[i] ? mb[i] : nw[i];
    if (k % 4 == 0) {
      for (int i = 1; i < 8; i++)
        o[i] = WADD(o[i], -o[0]);
      o[0] = 0;
    }
    out[k - 1] = WADD(m1, -m0);
  }
  free(beta);
}

/* ============================================================================ decoder object */

/* implementation selector; values follow srslte_tdec_impl_type_t (include/srslte/phy/fec/turbodecoder_impl.h:28-38) */
enum { ORC_AUTO = 0, ORC_GENERIC = 1, ORC_SSE = 2, ORC_SSE_WIN = 3, ORC_NEON_WIN = 4, ORC_AVX_WIN = 5, ORC_SSE8_WIN = 6, ORC_AVX8_WIN = 7 };

typedef struct {
  int       dec_type, force_not_sb;
  uint32_t  K;
  int       n_iter;
  orc_cfg_t cfg;          /* active MAP configuration */
  int       llr8;         /* arithmetic of the active decoder is 8-bit */
  int       input_sb;     /* input arrives in lane layout */
  uint16_t *fwd, *rev;
  int32_t * syst, *par0, *par1, *app1, *app2, *ext1, *ext2; /* K+3 each */
} orc_tdec_t;

void* orc_tdec_new(int dec_type, int force_not_sb)
{
  orc_tdec_t* h   = calloc(1, sizeof(*h));
  h->dec_type     = dec_type;
  h->force_not_sb = force_not_sb;
  size_t n        = ORC_MAX_K + 16;
  h->fwd          = malloc(n * 2);
  h->rev          = malloc(n * 2);
  int32_t** a[]   = {&h->syst, &h->par0, &h->par1, &h->app1, &h->app2, &h->ext1, &h->ext2};
  for (int i = 0; i < 7; i++)
    *a[i] = calloc(n, sizeof(int32_t));
  return h;
}
void orc_tdec_del(void* hh)
{
  orc_tdec_t* h = hh;
  if (!h)
    return;
  free(h->fwd); free(h->rev); free(h->syst); free(h->par0); free(h->par1);
  free(h->app1); free(h->app2); free(h->ext1); free(h->ext2); free(h);
}

static orc_cfg_t mk_cfg(int bits, uint32_t N)
{
  orc_cfg_t c;
  c.bits      = bits;
  c.N         = N;
  c.inf       = bits == 16 ? 10000 : 0;
  c.norm_max  = bits == 8;
  c.out_shift = bits == 8 ? 1 : 0;
  return c;
}

/* srslte_tdec_new_cb (src/phy/fec/turbodecoder.c:511-526) */
int orc_tdec_new_cb(void* hh, uint32_t K)
{
  orc_tdec_t* h  = hh;
  int         ci = orc_cbindex(K);
  if (K > ORC_MAX_K || ci < 0 || lte_qpp_table[ci].K != K)
    return -1;
  h->K      = K;
  h->n_iter = 0;
  return 0;
}

/* dispatch of tdec_iteration_16 / tdec_iteration_8 (src/phy/fec/turbodecoder.c:458-508) for `in_bits`-wide input.
 * patched8: for int8 input and 400 < K <= 800 convert the WHOLE lane-layout buffer (documented deviation from the
 * reference, which reads uninitialised memory there: SURVEY 8a-4 #1). */
static void select_impl(orc_tdec_t* h, int in_bits)
{
  uint32_t K = h->K, N;
  switch (h->dec_type) {
    case ORC_AUTO:
      if (in_bits == 16) {
        N           = orc_subblocks16(K);
        h->cfg      = mk_cfg(16, N);
        h->llr8     = 0;
        h->input_sb = N > 0; /* iter.h:41 input_is_interleaved = current_dec > 0 */
      } else {
        N = orc_subblocks8(K);
        if (N >= 16) {
          h->cfg  = mk_cfg(8, N);
          h->llr8 = 1;
        } else {
          h->cfg  = mk_cfg(16, N); /* widened to int16, then gen (N==0) or sse16 (N==8) */
          h->llr8 = 0;
        }
        h->input_sb = N > 0;
      }
      break;
    case ORC_GENERIC:
      h->cfg = mk_cfg(16, 0); h->llr8 = 0; h->input_sb = 0;
      break;
    case ORC_SSE_WIN:
      h->cfg = mk_cfg(16, 8); h->llr8 = 0; h->input_sb = 0; /* manual 16-bit: current_dec == 0 -> standard layout */
      break;
    case ORC_AVX_WIN:
      h->cfg = mk_cfg(16, 16); h->llr8 = 0; h->input_sb = 0;
      break;
    case ORC_SSE8_WIN:
      h->cfg = mk_cfg(8, 16); h->llr8 = 1; h->input_sb = 1;
      break;
    case ORC_AVX8_WIN:
      h->cfg = mk_cfg(8, 32); h->llr8 = 1; h->input_sb = 1;
      break;
    default:
      h->cfg = mk_cfg(16, 0); h->llr8 = 0; h->input_sb = 0;
  }
  if (h->force_not_sb)
    h->input_sb = 0;
}

/* srslte_vec_sub_sss (16-bit: wraps) / srslte_vec_sub_bbb (8-bit: saturates for i < floor(K/32)*32, then wraps;
 * AVX2 build, src/phy/utils/vector_simd.c:132-190) */
static void vec_sub(const orc_tdec_t* h, const int32_t* a, const int32_t* b, int32_t* z)
{
  uint32_t K = h->K, simd_end = h->llr8 ? (K / 32) * 32 : 0;
  for (uint32_t i = 0; i < K; i++) {
    int32_t v = a[i] - b[i];
    z[i]      = (h->llr8 && i < simd_end) ? sat(&h->cfg, v) : wrap(&h->cfg, v);
  }
}

/* One half-iteration: run_tdec_iteration_{16,8}bit (include/srslte/phy/fec/turbodecoder_iter.h:72-144).
 * input: the caller's LLR buffer widened to int32 (3K+12 standard, or 3(K+32)+12 lane layout). */
static void half_iteration(orc_tdec_t* h, const int32_t* input)
{
  uint32_t K = h->K, N = h->cfg.N;
  if (h->n_iter == 0) {
    orc_qpp(K, N, h->fwd, h->rev);
    if (h->input_sb) {
      /* iter.h:59-69,90-97: three planes (K+32 apart) + tail */
      memcpy(h->syst, input, K * sizeof(int32_t));
      memcpy(h->par0, input + (K + 32), K * sizeof(int32_t));
      memcpy(h->par1, input + 2 * (K + 32), K * sizeof(int32_t));
      const int32_t* t = input + 3 * (K + 32);
      for (uint32_t i = 0; i < 3; i++) {
        h->syst[K + i] = t[2 * i];
        h->par0[K + i] = t[2 * i + 1];
        h->app2[K + i] = t[6 + 2 * i];
        h->par1[K + i] = t[6 + 2 * i + 1];
      }
    } else {
      /* extract_input: win.h:880-923 (to lane layout) / gen.c:238-258 (natural order) */
      for (uint32_t n = 0; n < K; n++) {
        uint32_t j = N > 0 ? to_lane(n, K, N) : n;
        h->syst[j] = input[3 * n];
        h->par0[j] = input[3 * n + 1];
        h->par1[j] = input[3 * n + 2];
      }
      const int32_t* t = input + 3 * K;
      for (uint32_t i = 0; i < 3; i++) {
        h->syst[K + i] = t[2 * i];
        h->par0[K + i] = t[2 * i + 1];
        h->app2[K + i] = t[6 + 2 * i];
        h->par1[K + i] = t[6 + 2 * i + 1];
      }
    }
  }
  if (h->n_iter % 2 == 0) {
    if (h->n_iter)
      vec_sub(h, h->app1, h->ext1, h->app1);
    if (N)
      map_win(&h->cfg, h->syst, h->n_iter ? h->app1 : NULL, h->par0, h->ext1, K);
    else
      map_gen(h->syst, h->n_iter ? h->app1 : NULL, h->par0, h->ext1, K);
  } else {
    if (h->n_iter > 1)
      vec_sub(h, h->ext1, h->app1, h->ext1);
    for (uint32_t i = 0; i < K; i++)
      h->app2[h->rev[i]] = h->ext1[i];
    if (N)
      map_win(&h->cfg, h->app2, NULL, h->par1, h->ext2, K);
    else
      map_gen(h->app2, NULL, h->par1, h->ext2, K);
    for (uint32_t i = 0; i < K; i++)
      h->app1[h->fwd[i]] = h->ext2[i];
  }
  h->n_iter++;
}

/* tdec_decision_byte (turbodecoder.c:370-378) -> win.h:925-993 / gen.c:260-277 */
static void decision(const orc_tdec_t* h, uint8_t* out)
{
  const int32_t* llr = (h->n_iter % 2 == 0) ? h->app1 : h->ext1;
  uint32_t       K = h->K, N = h->cfg.N;
  memset(out, 0, K / 8);
  for (uint32_t n = 0; n < K; n++) {
    uint32_t j = N > 0 ? to_lane(n, K, N) : n;
    if (llr[j] > 0)
      out[n >> 3] |= (uint8_t)(0x80u >> (n & 7));
  }
}

static void widen(orc_tdec_t* h, const void* input, int in_bits, int32_t* dst)
{
  uint32_t K = h->K, n = h->input_sb ? 3 * (K + 32) + 12 : 3 * K + 12;
  for (uint32_t i = 0; i < n; i++) {
    int32_t v = in_bits == 16 ? (int32_t)((const int16_t*)input)[i] : (int32_t)((const int8_t*)input)[i];
    /* manual 8-bit decoder driven through the 16-bit entry point: convert_16_to_8 truncates
     * (src/phy/fec/turbodecoder.c:451-456, 499-504) */
    dst[i] = (in_bits == 16 && h->llr8) ? (int32_t)(int8_t)(uint8_t)v : v;
  }
}

/* srslte_tdec_iteration / srslte_tdec_iteration_8bit (turbodecoder.c:528-535, 552-558) */
void orc_tdec_iteration(void* hh, const void* input, int in_bits, uint8_t* out)
{
  orc_tdec_t* h = hh;
  if (h->K == 0)
    return;
  select_impl(h, in_bits);
  int32_t* w = malloc((3 * (ORC_MAX_K + 32) + 12) * sizeof(int32_t));
  if (h->n_iter == 0)
    widen(h, input, in_bits, w);
  half_iteration(h, w);
  decision(h, out);
  free(w);
}
/* srslte_tdec_run_all[_8bit] (turbodecoder.c:537-550, 560-578): do-while => at least one half-iteration */
int orc_tdec_run_all(void* hh, const void* input, int in_bits, uint8_t* out, uint32_t nof_iter, uint32_t K)
{
  orc_tdec_t* h = hh;
  if (orc_tdec_new_cb(h, K))
    return -1;
  select_impl(h, in_bits);
  int32_t* w = malloc((3 * (ORC_MAX_K + 32) + 12) * sizeof(int32_t));
  widen(h, input, in_bits, w);
  do {
    half_iteration(h, w);
  } while ((uint32_t)h->n_iter < nof_iter);
  decision(h, out);
  free(w);
  return 0;
}
int orc_tdec_n_iter(void* hh) { return ((orc_tdec_t*)hh)->n_iter; }
/* which: 0 app1, 1 app2, 2 ext1, 3 ext2 (as int16) */
void orc_tdec_get_llr(void* hh, int which, int16_t* dst, uint32_t n)
{
  orc_tdec_t*    h = hh;
  const int32_t* s = which == 0 ? h->app1 : which == 1 ? h->app2 : which == 2 ? h->ext1 : h->ext2;
  for (uint32_t i = 0; i < n; i++)
    dst[i] = (int16_t)s[i];
}

/* ================================================================== transport block (sch.c) */

#define ORC_MAX_CB 32 /* include/srslte/phy/common/phy_common.h:63 SRSLTE_MAX_CODEBLOCKS */
#define CRC24A 0x1864CFB
#define CRC24B 0x1800063

typedef struct {
  int16_t buffer[ORC_MAX_CB][ORC_SOFTBUFFER_SIZE]; /* int8 decoders view the same memory as int8 */
  uint8_t data[ORC_MAX_CB][ORC_MAX_K / 8];
  uint8_t cb_crc[ORC_MAX_CB];
  uint8_t tb_crc;
  void*   tdec;
} orc_softbuffer_t;

void* orc_softbuffer_new(void)
{
  orc_softbuffer_t* s = calloc(1, sizeof(*s));
  s->tdec             = orc_tdec_new(ORC_AUTO, 0);
  return s;
}
void orc_softbuffer_del(void* ss)
{
  orc_softbuffer_t* s = ss;
  if (s) {
    orc_tdec_del(s->tdec);
    free(s);
  }
}
/* srslte_softbuffer_rx_reset_cb (src/phy/fec/softbuffer.c:133-155), whole buffer */
void orc_softbuffer_reset(void* ss)
{
  orc_softbuffer_t* s = ss;
  void*             t = s->tdec;
  memset(s, 0, sizeof(*s));
  s->tdec = t;
}
int orc_softbuffer_get(void* ss, uint32_t cb, int16_t* dst, uint32_t n)
{
  memcpy(dst, ((orc_softbuffer_t*)ss)->buffer[cb], n * sizeof(int16_t));
  return 0;
}

/* decode_tb + decode_tb_cb (src/phy/phch/sch.c:363-488, 503-570).
 * Returns 0 / -1 / -2 like the reference; n_iter_out[cb] = half-iterations spent on CB cb (0 if skipped);
 * *avg_iter = sum / C.  is8 selects int8 e_bits and the 8-bit AUTO decoders. */
int orc_decode_tb(void* ss, uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const void* e_bits, int is8,
                  uint32_t max_iter, uint8_t* data, uint32_t* n_iter_out, float* avg_iter)
{
  orc_softbuffer_t* sb = ss;
  uint32_t          seg[9];
  if (orc_cbsegm(tbs, seg))
    return -1;
  uint32_t F = seg[0], C = seg[1], K1 = seg[2], K2 = seg[3], C1 = seg[6];
  if (tbs == 0 || C == 0)
    return 0;
  if (F)
    return -2;
  if (C > ORC_MAX_CB)
    return -1;
  data[tbs / 8 + 0] = 0;
  data[tbs / 8 + 1] = 0;
  data[tbs / 8 + 2] = 0;
  uint32_t total = 0;
  for (uint32_t cb = 0; cb < C; cb++) {
    uint32_t K = cb < C1 ? K1 : K2, rlen = C == 1 ? K : K - 24;
    if (n_iter_out)
      n_iter_out[cb] = 0;
    if (sb->cb_crc[cb]) {
      memcpy(&data[cb * rlen / 8], sb->data[cb], rlen / 8);
      continue;
    }
    uint32_t Gp = nof_e_bits / Qm, gamma = Gp % C, n_e = Qm * (Gp / C), rp = cb * n_e, n_e2 = n_e;
    if (cb > C - gamma) { /* strict '>' as in sch.c:398 */
      n_e2 = n_e + Qm;
      rp   = (C - gamma) * n_e + (cb - (C - gamma)) * n_e2;
    }
    if (is8)
      orc_rm_rx8((const int8_t*)e_bits + rp, (int8_t*)sb->buffer[cb], n_e2, K, rv, orc_subblocks8(K));
    else
      orc_rm_rx16((const int16_t*)e_bits + rp, sb->buffer[cb], n_e2, K, rv, orc_subblocks16(K));
    orc_tdec_new_cb(sb->tdec, K);
    uint32_t noi = 0;
    int      ok  = 0;
    do {
      orc_tdec_iteration(sb->tdec, sb->buffer[cb], is8 ? 8 : 16, &data[cb * rlen / 8]);
      noi++;
      uint32_t len_crc = C > 1 ? K : tbs + 24;
      if (orc_crc_bytes(C > 1 ? CRC24B : CRC24A, 24, &data[cb * rlen / 8], len_crc / 8) == 0) {
        sb->cb_crc[cb] = 1;
        ok             = 1;
      }
    } while (noi < max_iter && !ok);
    total += noi;
    if (n_iter_out)
      n_iter_out[cb] = noi;
  }
  sb->tb_crc = 1;
  for (uint32_t i = 0; i < C; i++)
    sb->tb_crc = sb->tb_crc && sb->cb_crc[i];
  if (!sb->tb_crc) {
    for (uint32_t i = 0; i < C; i++) {
      if (sb->cb_crc[i]) {
        uint32_t K = i < C1 ? K1 : K2, rlen = C == 1 ? K : K - 24;
        memcpy(sb->data[i], &data[i * rlen / 8], rlen / 8);
      }
    }
  }
  if (avg_iter)
    *avg_iter = (float)total / (float)C;
  if (!sb->tb_crc)
    return -1;
  uint32_t par_rx = orc_crc_bytes(CRC24A, 24, data, tbs / 8);
  uint32_t par_tx = ((uint32_t)data[tbs / 8] << 16) | ((uint32_t)data[tbs / 8 + 1] << 8) | data[tbs / 8 + 2];
  return (par_rx == par_tx && par_rx) ? 0 : -1;
}
void orc_softbuffer_get_crc(void* ss, uint8_t* cb_crc, uint32_t n)
{
  memcpy(cb_crc, ((orc_softbuffer_t*)ss)->cb_crc, n);
}

/* ------------------------------------------------------------------ single MAP call, exported for unit tests
 * (lane-layout int16 arrays in, K+3 elements for in/par; app may be NULL) */
void orc_map_win(int bits, uint32_t N, const int16_t* in, const int16_t* app, const int16_t* par, int16_t* out, uint32_t K)
{
  orc_cfg_t c = mk_cfg(bits, N);
  int32_t * i32 = calloc(K + 3, 4), *a32 = app ? calloc(K + 3, 4) : NULL, *p32 = calloc(K + 3, 4), *o32 = calloc(K + 3, 4);
  for (uint32_t i = 0; i < K + 3; i++) {
    i32[i] = in[i];
    p32[i] = par[i];
    if (a32)
      a32[i] = i < K ? app[i] : 0;
  }
  map_win(&c, i32, a32, p32, o32, K);
  for (uint32_t i = 0; i < K; i++)
    out[i] = (int16_t)o32[i];
  free(i32); free(p32); free(o32);
  if (a32) free(a32);
}

/* ============================================================== transmit side (test-input synthesis only)
 * Restates the reference encoder chain so tests can build valid code words / transport blocks anywhere:
 * srslte_tcod_encode (src/phy/fec/turbocoder.c:80-187), encode_tb_off (src/phy/phch/sch.c:235-349) and the
 * bit-selection of TS 36.212 5.1.4.1.2 via the receive table (e_k = codeword[table[k mod (3K+12)]]). */
static inline void rsc_step(uint8_t bit, uint8_t s[3], uint8_t* parity)
{
  uint8_t fb = bit ^ s[1] ^ s[2];   /* g0 = 1 + D^2 + D^3 */
  *parity    = fb ^ s[0] ^ s[2];    /* g1 = 1 + D + D^3   */
  s[2]       = s[1];
  s[1]       = s[0];
  s[0]       = fb;
}
/* in: K bits (one per byte) -> out: 3K+12 bits: (x_k, z_k, z'_k) triples then 12 tail bits */
int orc_tcod_encode(const uint8_t* in, uint8_t* out, uint32_t K)
{
  int ci = orc_cbindex(K);
  if (ci < 0 || lte_qpp_table[ci].K != K)
    return -1;
  uint64_t f1 = lte_qpp_table[ci].f1, f2 = lte_qpp_table[ci].f2;
  uint8_t  s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
  for (uint64_t i = 0; i < K; i++) {
    uint8_t p;
    out[3 * i] = in[i];
    rsc_step(in[i], s1, &p);
    out[3 * i + 1] = p;
    rsc_step(in[(f1 * i + f2 * i * i) % K], s2, &p);
    out[3 * i + 2] = p;
  }
  uint32_t k = 3 * K;
  for (int e = 0; e < 2; e++) {
    uint8_t* s = e ? s2 : s1;
    for (int j = 0; j < 3; j++) {
      uint8_t bit = s[1] ^ s[2], p; /* drives the register back to zero */
      out[k++]    = bit;
      rsc_step(bit, s, &p);
      out[k++] = p;
    }
  }
  return 0;
}

static void put_crc24(uint32_t crc, uint8_t* dst)
{
  dst[0] = (uint8_t)(crc >> 16);
  dst[1] = (uint8_t)(crc >> 8);
  dst[2] = (uint8_t)crc;
}

/* data: tbs/8 bytes -> e_bits: nof_e_bits bits (one per byte).  Returns 0 or -1 / -2 like encode_tb_off. */
int orc_encode_tb(uint32_t tbs, uint32_t Qm, uint32_t rv, uint32_t nof_e_bits, const uint8_t* data, uint8_t* e_bits)
{
  uint32_t seg[9];
  if (orc_cbsegm(tbs, seg))
    return -1;
  uint32_t F = seg[0], C = seg[1], K1 = seg[2], K2 = seg[3], C2 = seg[7];
  if (F)
    return -2;
  uint32_t B    = tbs + 24;
  uint8_t* tb   = calloc(B / 8 + 8, 1);
  memcpy(tb, data, tbs / 8);
  put_crc24(orc_crc_bytes(CRC24A, 24, tb, tbs / 8), tb + tbs / 8);
  uint32_t  Gp = nof_e_bits / Qm, gamma = Gp % C, rp = 0, wp = 0;
  uint8_t*  cb  = malloc(ORC_MAX_K / 8 + 8);
  uint8_t*  cbb = malloc(ORC_MAX_K);
  uint8_t*  cw  = malloc(3 * ORC_MAX_K + 12);
  uint16_t* tab = malloc((3 * ORC_MAX_K + 12) * sizeof(uint16_t));
  for (uint32_t i = 0; i < C; i++) {
    uint32_t K    = i < C2 ? K2 : K1; /* encoder order (sch.c:277-283) */
    uint32_t rlen = C > 1 ? K - 24 : K;
    uint32_t n_e  = (i + gamma + 1 <= C) ? Qm * (Gp / C) : Qm * ((Gp + C - 1) / C); /* sch.c:289-293 */
    memcpy(cb, tb + rp / 8, rlen / 8);
    if (C > 1)
      put_crc24(orc_crc_bytes(CRC24B, 24, cb, rlen / 8), cb + rlen / 8);
    for (uint32_t b = 0; b < K; b++)
      cbb[b] = (cb[b >> 3] >> (7 - (b & 7))) & 1u;
    orc_tcod_encode(cbb, cw, K);
    orc_rm_table(K, rv, 0, tab);
    for (uint32_t k = 0; k < n_e && wp + k < nof_e_bits; k++)
      e_bits[wp + k] = cw[tab[k % (3 * K + 12)]];
    rp += rlen;
    wp += n_e;
  }
  free(tb); free(cb); free(cbb); free(cw); free(tab);
  return 0;
}

/* ============================================================ soft demodulation + descrambling
 *
 * Restates src/phy/modem/demod_soft.c as built for x86 with SSE (LV_HAVE_SSE): srslte_demod_soft_demodulate_s (:896-919)
 * and _b (:921-945).  The reference's results depend on WHERE a symbol sits: the SSE bodies convert with
 * round-to-nearest and saturating packs and subtract integer thresholds, the scalar tails (the last n mod 4 / n mod 8
 * symbols) truncate, wrap and subtract float thresholds.  Both are written down element-wise here. */
#include <math.h>

static int32_t orc_cvt_rne(float f) /* _mm_cvtps_epi32 */
{
  if (!(f >= -2147483648.0f && f < 2147483648.0f))
    return INT32_MIN;
  return (int32_t)lrintf(f); /* default rounding mode: to nearest even */
}
static int32_t orc_cvt_trunc(float f) /* _mm_cvttps_epi32 / cvttss2si */
{
  if (!(f > -2147483904.0f && f < 2147483648.0f))
    return INT32_MIN;
  return (int32_t)f;
}
static int32_t orc_cvt_trunc_d(double d) /* cvttsd2si */
{
  if (!(d > -2147483649.0 && d < 2147483648.0))
    return INT32_MIN;
  return (int32_t)d;
}
static int16_t orc_sat16(int32_t v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : (int16_t)v); } /* _mm_packs_epi32 */
static int8_t  orc_sat8(int16_t v) { return v > 127 ? 127 : (v < -128 ? -128 : (int8_t)v); }             /* _mm_packs_epi16 */
static int16_t orc_w16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static int8_t  orc_w8(int32_t v) { return (int8_t)(uint8_t)(uint32_t)v; }
static int16_t orc_abs16(int16_t v) { return v < 0 ? orc_w16(-(int32_t)v) : v; } /* _mm_abs_epi16: abs(-32768) = -32768 */
static int8_t  orc_abs8(int8_t v) { return v < 0 ? orc_w8(-(int32_t)v) : v; }

/* srslte_mod_t: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM (phy_common.h) */
int orc_demod_s(int mod, const float* sym, int16_t* llr, int n)
{
  switch (mod) {
    case 0: /* demod_bpsk_lte_s :118-123 */
      for (int i = 0; i < n; i++) {
        const float t = -100.0f * (sym[2 * i] + sym[2 * i + 1]);
        llr[i]        = orc_w16(orc_cvt_trunc_d((double)t * M_SQRT1_2));
      }
      return 0;
    case 1: { /* demod_qpsk_lte_s :137-140 -> srslte_vec_convert_fi_simd (vector_simd.c:436-472, AVX2: 16 values per trip) */
      const float scale = (float)(-100 * M_SQRT2);
      const int   len = 2 * n, nb = len >= 16 ? ((len - 16) / 16 + 1) * 16 : 0;
      for (int e = 0; e < len; e++) {
        const int32_t v = orc_cvt_trunc(sym[e] * scale);
        llr[e]          = e < nb ? orc_sat16(v) : orc_w16(v);
      }
      return 0;
    }
    case 2: { /* demod_16qam_lte_s_sse :273-322 */
      const int16_t off  = orc_w16(orc_cvt_trunc(2 * 400 / sqrtf(10)));
      const float   thrf = 2 * 400 / sqrtf(10);
      for (int i = 0; i < n; i++) {
        const float re = sym[2 * i], im = sym[2 * i + 1];
        if (i < 4 * (n / 4)) {
          const int16_t vr = orc_sat16(orc_cvt_rne(re * -400.0f)), vi = orc_sat16(orc_cvt_rne(im * -400.0f));
          llr[4 * i + 0] = vr;
          llr[4 * i + 1] = vi;
          llr[4 * i + 2] = orc_w16((int32_t)orc_abs16(vr) - off);
          llr[4 * i + 3] = orc_w16((int32_t)orc_abs16(vi) - off);
        } else {
          const int16_t yre = orc_w16(orc_cvt_trunc(400.0f * re)), yim = orc_w16(orc_cvt_trunc(400.0f * im));
          llr[4 * i + 0] = orc_w16(-(int32_t)yre);
          llr[4 * i + 1] = orc_w16(-(int32_t)yim);
          llr[4 * i + 2] = orc_w16(orc_cvt_trunc((float)abs((int)yre) - thrf));
          llr[4 * i + 3] = orc_w16(orc_cvt_trunc((float)abs((int)yim) - thrf));
        }
      }
      return 0;
    }
    case 3: { /* demod_64qam_lte_s_sse :594-669 */
      const int16_t off1 = orc_w16(orc_cvt_trunc(4 * 700 / sqrtf(42))), off2 = orc_w16(orc_cvt_trunc(2 * 700 / sqrtf(42)));
      for (int i = 0; i < n; i++) {
        const float re = sym[2 * i], im = sym[2 * i + 1];
        int16_t     vr, vi;
        if (i < 4 * (n / 4)) {
          vr = orc_sat16(orc_cvt_rne(re * -700.0f));
          vi = orc_sat16(orc_cvt_rne(im * -700.0f));
          llr[6 * i + 0] = vr;
          llr[6 * i + 1] = vi;
          llr[6 * i + 2] = orc_w16((int32_t)orc_abs16(vr) - off1);
          llr[6 * i + 3] = orc_w16((int32_t)orc_abs16(vi) - off1);
          llr[6 * i + 4] = orc_w16((int32_t)orc_abs16(llr[6 * i + 2]) - off2);
          llr[6 * i + 5] = orc_w16((int32_t)orc_abs16(llr[6 * i + 3]) - off2);
        } else {
          vr = orc_w16(orc_cvt_trunc(700.0f * re));
          vi = orc_w16(orc_cvt_trunc(700.0f * im));
          llr[6 * i + 0] = orc_w16(-(int32_t)vr);
          llr[6 * i + 1] = orc_w16(-(int32_t)vi);
          llr[6 * i + 2] = orc_w16((int32_t)orc_w16(abs((int)vr)) - off1);
          llr[6 * i + 3] = orc_w16((int32_t)orc_w16(abs((int)vi)) - off1);
          llr[6 * i + 4] = orc_w16((int32_t)orc_w16(abs((int)llr[6 * i + 2])) - off2);
          llr[6 * i + 5] = orc_w16((int32_t)orc_w16(abs((int)llr[6 * i + 3])) - off2);
        }
      }
      return 0;
    }
    case 4: { /* demod_256qam_lte_s :849-869 */
      const float c8 = 8.0f / sqrtf(170.0f), c4 = 4.0f / sqrtf(170.0f), c2 = 2.0f / sqrtf(170.0f);
      for (int i = 0; i < n; i++) {
        float re = -sym[2 * i], im = -sym[2 * i + 1];
        llr[8 * i + 0] = orc_w16(orc_cvt_trunc(1000.0f * re));
        llr[8 * i + 1] = orc_w16(orc_cvt_trunc(1000.0f * im));
        re = fabsf(re) - c8;
        im = fabsf(im) - c8;
        llr[8 * i + 2] = orc_w16(orc_cvt_trunc(1000.0f * re));
        llr[8 * i + 3] = orc_w16(orc_cvt_trunc(1000.0f * im));
        re = fabsf(re) - c4;
        im = fabsf(im) - c4;
        llr[8 * i + 4] = orc_w16(orc_cvt_trunc(1000.0f * re));
        llr[8 * i + 5] = orc_w16(orc_cvt_trunc(1000.0f * im));
        re = fabsf(re) - c2;
        im = fabsf(im) - c2;
        llr[8 * i + 6] = orc_w16(orc_cvt_trunc(1000.0f * re));
        llr[8 * i + 7] = orc_w16(orc_cvt_trunc(1000.0f * im));
      }
      return 0;
    }
  }
  return -1;
}

int orc_demod_b(int mod, const float* sym, int8_t* llr, int n)
{
  switch (mod) {
    case 0: /* demod_bpsk_lte_b :111-116 */
      for (int i = 0; i < n; i++) {
        const float t = -20.0f * (sym[2 * i] + sym[2 * i + 1]);
        llr[i]        = orc_w8(orc_cvt_trunc_d((double)t * M_SQRT1_2));
      }
      return 0;
    case 1: { /* demod_qpsk_lte_b :132-135 -> srslte_vec_convert_fb_simd (vector_simd.c:524-590, SSE: 16 values per trip) */
      const float scale = (float)(-20 * M_SQRT2);
      const int   len = 2 * n, nb = len >= 16 ? ((len - 16) / 16 + 1) * 16 : 0;
      for (int e = 0; e < len; e++) {
        const int32_t v = orc_cvt_trunc(sym[e] * scale);
        llr[e]          = e < nb ? orc_sat8(orc_sat16(v)) : orc_w8(v);
      }
      return 0;
    }
    case 2: { /* demod_16qam_lte_b_sse :324-382 */
      const int8_t off  = orc_w8(orc_cvt_trunc(2 * 30 / sqrtf(10)));
      const float  thrf = 2 * 30 / sqrtf(10);
      for (int i = 0; i < n; i++) {
        const float re = sym[2 * i], im = sym[2 * i + 1];
        if (i < 8 * (n / 8)) {
          const int8_t vr = orc_sat8(orc_sat16(orc_cvt_rne(re * -30.0f))), vi = orc_sat8(orc_sat16(orc_cvt_rne(im * -30.0f)));
          llr[4 * i + 0] = vr;
          llr[4 * i + 1] = vi;
          llr[4 * i + 2] = orc_w8((int32_t)orc_abs8(vr) - off);
          llr[4 * i + 3] = orc_w8((int32_t)orc_abs8(vi) - off);
        } else {
          const int yre = orc_w8(orc_cvt_trunc(30.0f * re)), yim = orc_w8(orc_cvt_trunc(30.0f * im));
          llr[4 * i + 0] = orc_w8(-yre);
          llr[4 * i + 1] = orc_w8(-yim);
          llr[4 * i + 2] = orc_w8(orc_cvt_trunc((float)abs(yre) - thrf));
          llr[4 * i + 3] = orc_w8(orc_cvt_trunc((float)abs(yim) - thrf));
        }
      }
      return 0;
    }
    case 3: { /* demod_64qam_lte_b_sse :671-755 */
      const int8_t off1 = orc_w8(orc_cvt_trunc(4 * 40 / sqrtf(42))), off2 = orc_w8(orc_cvt_trunc(2 * 40 / sqrtf(42)));
      for (int i = 0; i < n; i++) {
        const float re = sym[2 * i], im = sym[2 * i + 1];
        if (i < 8 * (n / 8)) {
          const int8_t vr = orc_sat8(orc_sat16(orc_cvt_rne(re * -40.0f))), vi = orc_sat8(orc_sat16(orc_cvt_rne(im * -40.0f)));
          llr[6 * i + 0] = vr;
          llr[6 * i + 1] = vi;
          llr[6 * i + 2] = orc_w8((int32_t)orc_abs8(vr) - off1);
          llr[6 * i + 3] = orc_w8((int32_t)orc_abs8(vi) - off1);
          llr[6 * i + 4] = orc_w8((int32_t)orc_abs8(llr[6 * i + 2]) - off2);
          llr[6 * i + 5] = orc_w8((int32_t)orc_abs8(llr[6 * i + 3]) - off2);
        } else {
          const int8_t vr = orc_w8(orc_cvt_trunc(40.0f * re)), vi = orc_w8(orc_cvt_trunc(40.0f * im));
          llr[6 * i + 0] = orc_w8(-(int32_t)vr);
          llr[6 * i + 1] = orc_w8(-(int32_t)vi);
          llr[6 * i + 2] = orc_w8((int32_t)orc_w8(abs((int)vr)) - off1);
          llr[6 * i + 3] = orc_w8((int32_t)orc_w8(abs((int)vi)) - off1);
          llr[6 * i + 4] = orc_w8((int32_t)orc_w8(abs((int)llr[6 * i + 2])) - off2);
          llr[6 * i + 5] = orc_w8((int32_t)orc_w8(abs((int)llr[6 * i + 3])) - off2);
        }
      }
      return 0;
    }
    case 4: { /* demod_256qam_lte_b :827-847 */
      const float c8 = 8.0f / sqrtf(170.0f), c4 = 4.0f / sqrtf(170.0f), c2 = 2.0f / sqrtf(170.0f);
      for (int i = 0; i < n; i++) {
        float re = -sym[2 * i], im = -sym[2 * i + 1];
        llr[8 * i + 0] = orc_w8(orc_cvt_trunc(50.0f * re));
        llr[8 * i + 1] = orc_w8(orc_cvt_trunc(50.0f * im));
        re = fabsf(re) - c8;
        im = fabsf(im) - c8;
        llr[8 * i + 2] = orc_w8(orc_cvt_trunc(50.0f * re));
        llr[8 * i + 3] = orc_w8(orc_cvt_trunc(50.0f * im));
        re = fabsf(re) - c4;
        im = fabsf(im) - c4;
        llr[8 * i + 4] = orc_w8(orc_cvt_trunc(50.0f * re));
        llr[8 * i + 5] = orc_w8(orc_cvt_trunc(50.0f * im));
        re = fabsf(re) - c2;
        im = fabsf(im) - c2;
        llr[8 * i + 6] = orc_w8(orc_cvt_trunc(50.0f * re));
        llr[8 * i + 7] = orc_w8(orc_cvt_trunc(50.0f * im));
      }
      return 0;
    }
  }
  return -1;
}

/* Pseudo-random sequence of TS 36.211 7.2 (src/phy/common/sequence.c: srslte_sequence_LTE_pr), packed as the
 * reference's c_bytes (srslte_bit_pack_vector: first bit = MSB).  out: (len + 7) / 8 bytes. */
void orc_sequence_bytes(uint32_t c_init, uint32_t len, uint8_t* out)
{
  const uint32_t Nc = 1600;
  uint8_t*       x1 = (uint8_t*)calloc(Nc + len + 31, 1);
  uint8_t*       x2 = (uint8_t*)calloc(Nc + len + 31, 1);
  x1[0] = 1;
  for (int i = 0; i < 31; i++)
    x2[i] = (c_init >> i) & 1u;
  for (uint32_t n = 0; n + 31 < Nc + len + 31; n++) {
    x1[n + 31] = x1[n + 3] ^ x1[n];
    x2[n + 31] = x2[n + 3] ^ x2[n + 2] ^ x2[n + 1] ^ x2[n];
  }
  memset(out, 0, (len + 7) / 8);
  for (uint32_t n = 0; n < len; n++)
    if (x1[n + Nc] ^ x2[n + Nc])
      out[n / 8] |= (uint8_t)(0x80u >> (n % 8));
  free(x1);
  free(x2);
}

/* srslte_scrambling_s_offset / _sb_offset (scrambling.c:43-53) with the sequence given as packed bytes:
 * srslte_vec_neg_{sss,bbb} against c_short / c_char = +-1 -> wrapping negation where the sequence bit is 1 */
void orc_descramble_s(const uint8_t* c_bytes, int16_t* data, int len)
{
  for (int i = 0; i < len; i++)
    if (c_bytes[i / 8] & (0x80u >> (i % 8)))
      data[i] = orc_w16(-(int32_t)data[i]);
}
void orc_descramble_b(const uint8_t* c_bytes, int8_t* data, int len)
{
  for (int i = 0; i < len; i++)
    if (c_bytes[i / 8] & (0x80u >> (i % 8)))
      data[i] = orc_w8(-(int32_t)data[i]);
}

/* ============================================================ PUSCH pre-steps
 * What srslte_ulsch_decode (src/phy/phch/sch.c:1105-1180) does between the descrambler and decode_tb: locate the
 * HARQ-ACK and RI resource elements, take their LLRs out, zero the ACK ones, de-interleave the rest (TS 36.212
 * 5.2.2.8) and offset past the CQI.  The UCI payload decoders themselves (srslte_uci_decode_ack_ri's correlators,
 * the CQI block/Viterbi decoders) are control-plane code and stay in the reference. */

/* 36.213 Tables 8.6.3-1/-2/-3 as the reference holds them (sch.c:42-87); out-of-range indexes fall back like there */
float orc_beta_offset(int which, uint32_t idx)
{
  static const float harq[16] = {2.0, 2.5, 3.125, 4.0, 5.0, 6.250, 8.0, 10.0, 12.625, 15.875, 20.0, 31.0, 50.0, 80.0, 126.0, -1.0};
  static const float ri[16]   = {1.25, 1.625, 2.0, 2.5, 3.125, 4.0, 5.0, 6.25, 8.0, 10.0, 12.625, 15.875, 20.0, -1.0, -1.0, -1.0};
  static const float cqi[16]  = {-1.0, -1.0, 1.125, 1.25, 1.375, 1.625, 1.750, 2.0, 2.25, 2.5, 2.875, 3.125, 3.5, 4.0, 5.0, 6.25};
  switch (which) {
    case 0: return idx < 15 ? harq[idx] : harq[0];
    case 1: return idx < 13 ? ri[idx] : ri[0];
    default: return (idx > 1 && idx < 16) ? cqi[idx] : cqi[2];
  }
}

/* Q_prime_ri_ack (src/phy/phch/uci.c:606-630): coded symbols of an O-bit ACK or RI; K_segm = C1*K1 + C2*K2 (sch.c:1123) */
uint32_t orc_qprime_ri_ack(uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb, uint32_t O, uint32_t O_cqi, float beta)
{
  uint32_t K = K_segm;
  if (K == 0)
    K = O_cqi <= 11 ? O_cqi : O_cqi + 8;
  float    f = (float)O * L_prb * 12 * nof_symb * beta / K;
  uint32_t x = (uint32_t)ceilf(f);
  uint32_t m = 4 * L_prb * 12;
  return x < m ? x : m;
}

/* Q_prime_cqi (uci.c:329-345) */
uint32_t orc_qprime_cqi(uint32_t K_segm, uint32_t L_prb, uint32_t nof_symb, uint32_t O, float beta, uint32_t Q_prime_ri)
{
  uint32_t L = O < 11 ? 0 : 8;
  uint32_t x = 999999;
  if (K_segm > 0)
    x = (uint32_t)ceilf((float)(O + L) * L_prb * 12 * nof_symb * beta / K_segm);
  uint32_t m = L_prb * 12 * nof_symb - Q_prime_ri;
  return x < m ? x : m;
}

/* uci_ulsch_interleave_ack_gen / _ri_gen (uci.c:551-605): first of the Qm positions of coded symbol idx, or -1 where
 * the reference logs an error (the symbol would sit above the first row) */
int64_t orc_ulsch_uci_position(int is_ri, uint32_t idx, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs)
{
  static const uint32_t ack_norm[4] = {2, 3, 8, 9}, ack_ext[4] = {1, 2, 6, 7};
  static const uint32_t ri_norm[4] = {1, 4, 7, 10}, ri_ext[4] = {0, 3, 5, 8};
  uint32_t rows = H_prime_total / N_pusch_symbs;
  if (rows < 1 + idx / 4)
    return -1;
  uint32_t row    = rows - 1 - idx / 4;
  uint32_t colidx = (3 * idx) % 4;
  uint32_t col    = N_pusch_symbs > 10 ? (is_ri ? ri_norm : ack_norm)[colidx] : (is_ri ? ri_ext : ack_ext)[colidx];
  return (int64_t)row * Qm + (int64_t)rows * col * Qm;
}

/* The data movement of srslte_ulsch_decode, in its order:
 *   1. ACK LLRs are read at their positions (uci.c:846-857) and the positions zeroed in q_bits (sch.c:1067-1070);
 *   2. RI LLRs are read at theirs (they stay in q_bits);
 *   3. ulsch_deinterleave (sch.c:992-1019): the table of ulsch_interleave_gen (:658-679) numbers the positions that
 *      hold no RI row by row, column by column, and sends every RI position to index 0; srslte_vec_lut_sis
 *      (utils/vector.c:136-141) then stores g[lut[i]] = q[i] for i ascending -- so with RI present g[0] ends up
 *      holding the LLR of the LAST RI position, not the first data/CQI LLR (reference behaviour, kept).
 * q_bits is modified like the reference modifies it; g_bits receives (H' - Q'_ri) * Qm values, the tail is untouched.
 * Returns 0, or -1 for geometries the reference cannot index (see orc_ulsch_uci_position). */
int orc_ulsch_deinterleave(int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, uint32_t Q_prime_ack,
                           uint32_t Q_prime_ri, int16_t* g_bits, int16_t* ack_llr, int16_t* ri_llr)
{
  if (Qm == 0 || N_pusch_symbs == 0 || H_prime_total < N_pusch_symbs || H_prime_total % N_pusch_symbs)
    return -1;
  const uint32_t rows = H_prime_total / N_pusch_symbs, cols = N_pusch_symbs, nof_bits = H_prime_total * Qm;
  if (Q_prime_ack > 4 * rows || Q_prime_ri > 4 * rows)
    return -1;
  for (uint32_t i = 0; i < Q_prime_ack; i++) {
    int64_t p = orc_ulsch_uci_position(0, i, Qm, H_prime_total, N_pusch_symbs);
    for (uint32_t k = 0; k < Qm; k++) {
      if (ack_llr)
        ack_llr[i * Qm + k] = q_bits[p + k];
      q_bits[p + k] = 0;
    }
  }
  uint8_t* ri_present = (uint8_t*)calloc(nof_bits, 1);
  for (uint32_t i = 0; i < Q_prime_ri; i++) {
    int64_t p = orc_ulsch_uci_position(1, i, Qm, H_prime_total, N_pusch_symbs);
    for (uint32_t k = 0; k < Qm; k++) {
      if (ri_llr)
        ri_llr[i * Qm + k] = q_bits[p + k];
      ri_present[p + k] = 1;
    }
  }
  uint32_t* lut = (uint32_t*)malloc(sizeof(uint32_t) * nof_bits);
  uint32_t  idx = 0;
  for (uint32_t j = 0; j < rows; j++)
    for (uint32_t i = 0; i < cols; i++)
      for (uint32_t k = 0; k < Qm; k++) {
        uint32_t p = j * Qm + i * rows * Qm + k;
        lut[p]     = ri_present[p] ? 0 : idx++;
      }
  for (uint32_t i = 0; i < nof_bits; i++)
    g_bits[lut[i]] = q_bits[i];
  free(lut);
  free(ri_present);
  return 0;
}

/* ================================================================== csi_correction (lib/src/phy/phch/pdsch.c:628-741)
 * Scales the soft bits of a codeword by the channel state information of their resource element, between the soft
 * demodulator and the descrambler (pdsch.c:843-846, only when cfg->csi_enable).  Restated with the reference's quirks:
 *  - int16, SSE bodies: e = (e * c16) >> 16 with c16 = sat16(rne(csi * (32767 / csi_max))) -- _mm_cvtps_pi16 + _mm_mulhi_pi16,
 *    i.e. HALF the scale of the scalar tail;
 *  - QPSK body (groups of 4 soft bits = two symbols, :676-686): _mm_blend_ps(csi1, csi2, 3) takes elements 0,1 from csi2, so
 *    the two symbols of a pair use EACH OTHER's csi;
 *  - 64QAM body (groups of 12 = two symbols, :697-711): soft bits 4,5 of the first symbol use the second symbol's csi and soft
 *    bits 0,1 of the second symbol the first one's;
 *  - scalar tails and the whole int8 path: e = (T)((float)e * (csi / csi_max)), truncation;
 *  - BPSK has no SSE body.
 * mod: srslte_mod_t 0..4; nof_bits = symbols * Qm. */
static int16_t orc_csi_c16(float csi, float scale)
{
  int32_t v = orc_cvt_rne(csi * scale); /* cvtps2pi: round to nearest even */
  return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); /* packssdw */
}
void orc_csi_correction(const float* csi, void* e, uint32_t nof_bits, int mod, int is8)
{
  const uint32_t qm = mod == 0 ? 1 : 2 * (uint32_t)mod, nsym = nof_bits / qm;
  /* srslte_vec_max_fi: index of the maximum (the value is what matters) */
  float csi_max = 1.0f;
  if (nsym > 0) {
    csi_max = csi[0];
    for (uint32_t i = 1; i < nsym; i++)
      if (csi[i] > csi_max)
        csi_max = csi[i];
  }
  if (is8) {
    int8_t* eb = e;
    for (uint32_t i = 0; i < nsym; i++) {
      const float c = csi[i] / csi_max;
      for (uint32_t k = 0; k < qm; k++)
        eb[qm * i + k] = (int8_t)(int32_t)((float)eb[qm * i + k] * c);
    }
    return;
  }
  int16_t*    es    = e;
  const float scale = (float)32767 / csi_max;
  uint32_t    i     = 0; /* soft-bit index while in the SSE bodies */
#define ORC_MULHI(x, c) ((int16_t)(((int32_t)(x) * (int32_t)(c)) >> 16))
  switch (mod) {
    case 1:
      for (; i + 3 < nof_bits; i += 4) {
        const int16_t c0 = orc_csi_c16(csi[i / 2], scale), c1 = orc_csi_c16(csi[i / 2 + 1], scale);
        es[i] = ORC_MULHI(es[i], c1); es[i + 1] = ORC_MULHI(es[i + 1], c1);
        es[i + 2] = ORC_MULHI(es[i + 2], c0); es[i + 3] = ORC_MULHI(es[i + 3], c0);
      }
      break;
    case 2:
      for (; i + 3 < nof_bits; i += 4) {
        const int16_t c = orc_csi_c16(csi[i / 4], scale);
        for (int k = 0; k < 4; k++)
          es[i + k] = ORC_MULHI(es[i + k], c);
      }
      break;
    case 3:
      for (; i + 11 < nof_bits; i += 12) {
        const int16_t c0 = orc_csi_c16(csi[i / 6], scale), c1 = orc_csi_c16(csi[i / 6 + 1], scale);
        for (int k = 0; k < 4; k++)
          es[i + k] = ORC_MULHI(es[i + k], c0);
        es[i + 4] = ORC_MULHI(es[i + 4], c1); es[i + 5] = ORC_MULHI(es[i + 5], c1);
        es[i + 6] = ORC_MULHI(es[i + 6], c0); es[i + 7] = ORC_MULHI(es[i + 7], c0);
        for (int k = 8; k < 12; k++)
          es[i + k] = ORC_MULHI(es[i + k], c1);
      }
      break;
    case 4:
      for (; i + 7 < nof_bits; i += 8) {
        const int16_t c = orc_csi_c16(csi[i / 8], scale);
        for (int k = 0; k < 8; k++)
          es[i + k] = ORC_MULHI(es[i + k], c);
      }
      break;
    default:
      break;
  }
#undef ORC_MULHI
  for (uint32_t s = i / qm; s < nsym; s++) {
    const float c = csi[s] / csi_max;
    for (uint32_t k = 0; k < qm; k++)
      es[qm * s + k] = (int16_t)(int32_t)((float)es[qm * s + k] * c);
  }
}
