/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Thin flat-C shim over the UNMODIFIED reference (srsLTE 20.10.1) compiled by
 * oracle/build_ref.sh into oracle/_ref/libsrslte_ref.so.  It only forwards to
 * the reference's own public entry points so that Python (ctypes) tests and
 * bench.py's CPU-baseline arm can drive them:
 *
 *   srslte_tdec_*            lib/include/srslte/phy/fec/turbodecoder.h:97-121
 *   srslte_rm_turbo_rx_lut*  lib/include/srslte/phy/fec/rm_turbo.h:75-89
 *   srslte_crc_*             lib/include/srslte/phy/fec/crc.h:48-74
 *   srslte_cbsegm*           lib/include/srslte/phy/fec/cbsegm.h:46-52
 *   srslte_dlsch_encode2 / srslte_dlsch_decode2   lib/src/phy/phch/sch.c:577,611
 *   srslte_demod_soft_demodulate_{s,b}             lib/src/phy/modem/demod_soft.c:896-945
 *   srslte_ulsch_encode / srslte_ulsch_decode      lib/src/phy/phch/sch.c:1105-1330
 *   srslte_sequence_LTE_pr, srslte_scrambling_{s,sb}_offset   lib/src/phy/common/sequence.c, scrambling/scrambling.c:43-53
 *
 * Nothing in here re-implements reference arithmetic.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "srslte/phy/common/phy_common.h"
#include "srslte/phy/fec/cbsegm.h"
#include "srslte/phy/fec/crc.h"
#include "srslte/phy/fec/rm_turbo.h"
#include "srslte/phy/fec/softbuffer.h"
#include "srslte/phy/fec/tc_interl.h"
#include "srslte/phy/fec/turbocoder.h"
#include "srslte/phy/fec/turbodecoder.h"
#include "srslte/phy/phch/pdsch_cfg.h"
#include "srslte/phy/phch/sch.h"
#include "srslte/phy/common/sequence.h"
#include "srslte/phy/modem/demod_soft.h"
#include "srslte/phy/scrambling/scrambling.h"

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* pin the calling thread to the t-th CPU of the process's affinity mask (pinned threads) */
static void pin_to_cpu(int t)
{
  cpu_set_t all, one;
  if (sched_getaffinity(0, sizeof(all), &all) != 0)
    return;
  int n = CPU_COUNT(&all);
  if (n <= 0)
    return;
  int want = t % n, seen = 0;
  for (int c = 0; c < CPU_SETSIZE; c++) {
    if (!CPU_ISSET(c, &all))
      continue;
    if (seen++ == want) {
      CPU_ZERO(&one);
      CPU_SET(c, &one);
      pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
      return;
    }
  }
}

/* the latency probes pin the CALLING thread; its mask (which threads created later inherit) is put back afterwards */
static int  save_affinity(cpu_set_t* m) { return pthread_getaffinity_np(pthread_self(), sizeof(*m), m); }
static void restore_affinity(const cpu_set_t* m, int have) { if (have == 0) pthread_setaffinity_np(pthread_self(), sizeof(*m), m); }

int ref_init(void)
{
  srslte_rm_turbo_gentables();
  return 0;
}

/* ------------------------------------------------------------------ tables */
int ref_cbsegm(uint32_t tbs, uint32_t out[9])
{
  srslte_cbsegm_t s;
  memset(&s, 0, sizeof(s));
  int r  = srslte_cbsegm(&s, tbs);
  out[0] = s.F;
  out[1] = s.C;
  out[2] = s.K1;
  out[3] = s.K2;
  out[4] = s.K1_idx;
  out[5] = s.K2_idx;
  out[6] = s.C1;
  out[7] = s.C2;
  out[8] = s.tbs;
  return r;
}
int      ref_cbsize(uint32_t idx) { return srslte_cbsegm_cbsize(idx); }
int      ref_cbindex(uint32_t K) { return srslte_cbsegm_cbindex(K); }
uint32_t ref_subblocks16(uint32_t K) { return srslte_tdec_autoimp_get_subblocks(K); }
uint32_t ref_subblocks8(uint32_t K) { return srslte_tdec_autoimp_get_subblocks_8bit(K); }

int ref_qpp(uint32_t K, uint32_t nsb, uint16_t* fwd, uint16_t* rev)
{
  srslte_tc_interl_t t;
  if (srslte_tc_interl_init(&t, K) < 0)
    return -1;
  int r = srslte_tc_interl_LTE_gen_interl(&t, K, nsb);
  if (r == 0) {
    memcpy(fwd, t.forward, K * sizeof(uint16_t));
    memcpy(rev, t.reverse, K * sizeof(uint16_t));
  }
  srslte_tc_interl_free(&t);
  return r;
}

/* --------------------------------------------------------------------- crc */
uint32_t ref_crc_byte(uint32_t poly, int order, uint8_t* data, int len_bits)
{
  srslte_crc_t c;
  srslte_crc_init(&c, poly, order);
  return srslte_crc_checksum_byte(&c, data, len_bits);
}
uint32_t ref_crc_bits(uint32_t poly, int order, uint8_t* bits, int len_bits)
{
  srslte_crc_t c;
  srslte_crc_init(&c, poly, order);
  return srslte_crc_checksum(&c, bits, len_bits);
}

/* --------------------------------------------------------- rate (de)matching */
int ref_rm_rx_lut16(int16_t* in, int16_t* out, uint32_t E, uint32_t cb_idx, uint32_t rv, int enable_sb)
{
  return srslte_rm_turbo_rx_lut_(in, out, E, cb_idx, rv, enable_sb ? true : false);
}
int ref_rm_rx_lut8(int8_t* in, int8_t* out, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  return srslte_rm_turbo_rx_lut_8bit(in, out, E, cb_idx, rv);
}
/* procedural transmitter rate matcher on unpacked bits (rm_turbo.c:870) */
int ref_rm_tx(uint8_t* in_bits, uint32_t in_len, uint8_t* out_bits, uint32_t E, uint32_t rv)
{
  uint32_t buff_len = 3 * 6176 + 64;
  uint8_t* w        = calloc(buff_len, 1);
  int      r        = srslte_rm_turbo_tx(w, buff_len, in_bits, in_len, out_bits, E, rv);
  free(w);
  return r;
}
/* procedural float receiver (rm_turbo.c:950), the function rm_turbo_test.c:171-190 compares against */
int ref_rm_rx_float(float* in, uint32_t E, float* out, uint32_t out_len, uint32_t rv)
{
  uint32_t buff_len = 3 * 6176 + 64;
  float*   w        = calloc(buff_len, sizeof(float));
  for (uint32_t i = 0; i < buff_len; i++)
    w[i] = SRSLTE_RX_NULL;
  int r = srslte_rm_turbo_rx(w, buff_len, in, E, out, out_len, rv, 0);
  free(w);
  return r;
}

/* ----------------------------------------------------------------- encoder */
int ref_tcod_encode(uint8_t* bits, uint8_t* out_bits, uint32_t K)
{
  srslte_tcod_t t;
  if (srslte_tcod_init(&t, 6144))
    return -1;
  int r = srslte_tcod_encode(&t, bits, out_bits, K);
  free(t.temp); /* srslte_tcod_free() would also free the process-global interleaver tables other objects use */
  return r;
}

/* ----------------------------------------------------------------- decoder */
typedef struct {
  srslte_tdec_t td;
  int16_t*      conv; /* patched 8->16 conversion buffer */
  int           use_patched;
} ref_tdec_t;

void* ref_tdec_new(uint32_t max_k, int dec_type, int force_not_sb)
{
  ref_tdec_t* h = calloc(1, sizeof(ref_tdec_t));
  if (srslte_tdec_init_manual(&h->td, max_k, (srslte_tdec_impl_type_t)dec_type)) {
    free(h);
    return NULL;
  }
  if (force_not_sb)
    srslte_tdec_force_not_sb(&h->td);
  h->conv = calloc(3 * (6144 + 32) + 12 + 64, sizeof(int16_t));
  return h;
}
void ref_tdec_del(void* hh)
{
  ref_tdec_t* h = hh;
  if (!h)
    return;
  srslte_tdec_free(&h->td);
  free(h->conv);
  free(h);
}
int ref_tdec_new_cb(void* hh, uint32_t K) { return srslte_tdec_new_cb(&((ref_tdec_t*)hh)->td, K); }
int ref_tdec_n_iter(void* hh) { return srslte_tdec_get_nof_iterations(&((ref_tdec_t*)hh)->td); }
int ref_tdec_current(void* hh, int out[3])
{
  ref_tdec_t* h = hh;
  out[0]        = (int)h->td.current_llr_type;
  out[1]        = (int)h->td.current_dec;
  out[2]        = (int)h->td.current_inter_idx;
  return 0;
}

void ref_tdec_iteration16(void* hh, int16_t* input, uint8_t* out)
{
  srslte_tdec_iteration(&((ref_tdec_t*)hh)->td, input, out);
}

/* patched!=0: for 400 < K <= 800 the reference's int8 AUTO path reads uninitialised memory
 * (turbodecoder.c:478 converts 3K+12 of the 3(K+32)+12 SB-8 elements); the patched variant converts
 * the whole SB buffer to int16 and then runs the reference's own 16-bit path. */
void ref_tdec_iteration8(void* hh, int8_t* input, uint8_t* out, int patched)
{
  ref_tdec_t* h = hh;
  uint32_t    K = h->td.current_long_cb;
  if (patched && K > 400 && K <= 800 && h->td.dec_type == SRSLTE_TDEC_AUTO) {
    if (h->td.n_iter == 0) {
      uint32_t n = 3 * (K + 32) + 12;
      for (uint32_t i = 0; i < n; i++)
        h->conv[i] = (int16_t)input[i];
    }
    h->use_patched = 1;
    srslte_tdec_iteration(&h->td, h->conv, out);
  } else {
    h->use_patched = 0;
    srslte_tdec_iteration_8bit(&h->td, input, out);
  }
}

/* copy an internal LLR array widened to int16: which = 0 app1, 1 app2, 2 ext1, 3 ext2 */
int ref_tdec_get_llr(void* hh, int which, int16_t* dst, uint32_t n)
{
  ref_tdec_t* h   = hh;
  void*       src = which == 0 ? h->td.app1 : which == 1 ? h->td.app2 : which == 2 ? h->td.ext1 : h->td.ext2;
  if (h->td.current_llr_type == SRSLTE_TDEC_16) {
    memcpy(dst, src, n * sizeof(int16_t));
  } else {
    for (uint32_t i = 0; i < n; i++)
      dst[i] = ((int8_t*)src)[i];
  }
  return (int)h->td.current_llr_type;
}

int ref_tdec_run_all16(void* hh, int16_t* input, uint8_t* out, uint32_t nof_iter, uint32_t K)
{
  return srslte_tdec_run_all(&((ref_tdec_t*)hh)->td, input, out, nof_iter, K);
}
int ref_tdec_run_all8(void* hh, int8_t* input, uint8_t* out, uint32_t nof_iter, uint32_t K)
{
  return srslte_tdec_run_all_8bit(&((ref_tdec_t*)hh)->td, input, out, nof_iter, K);
}

/* ------------------------------------------------- TB level (the real sch.c) */
typedef struct {
  srslte_sch_t           sch;
  srslte_softbuffer_rx_t rx;
  srslte_softbuffer_tx_t tx;
} ref_sch_t;

static srslte_mod_t mod_from_qm(uint32_t Qm)
{
  switch (Qm) {
    case 1:
      return SRSLTE_MOD_BPSK;
    case 2:
      return SRSLTE_MOD_QPSK;
    case 4:
      return SRSLTE_MOD_16QAM;
    case 6:
      return SRSLTE_MOD_64QAM;
    default:
      return SRSLTE_MOD_256QAM;
  }
}

void* ref_sch_new(int llr_is_8bit, uint32_t max_iter, uint32_t nof_prb)
{
  ref_sch_t* s = calloc(1, sizeof(ref_sch_t));
  if (srslte_sch_init(&s->sch))
    return NULL;
  s->sch.llr_is_8bit = llr_is_8bit ? true : false;
  srslte_sch_set_max_noi(&s->sch, max_iter);
  if (srslte_softbuffer_rx_init(&s->rx, nof_prb))
    return NULL;
  if (srslte_softbuffer_tx_init(&s->tx, nof_prb))
    return NULL;
  srslte_softbuffer_rx_reset(&s->rx);
  srslte_softbuffer_tx_reset(&s->tx);
  return s;
}
void ref_sch_del(void* ss)
{
  ref_sch_t* s = ss;
  if (!s)
    return;
  srslte_softbuffer_rx_free(&s->rx);
  srslte_softbuffer_tx_free(&s->tx);
  /* srslte_sch_free() also tears down the process-global rm tables; regenerate them for other users */
  srslte_sch_free(&s->sch);
  srslte_rm_turbo_gentables();
  free(s);
}
void ref_sch_set_max_noi(void* ss, uint32_t n) { srslte_sch_set_max_noi(&((ref_sch_t*)ss)->sch, n); }
void ref_sch_reset_rx(void* ss, uint32_t tbs) { srslte_softbuffer_rx_reset_tbs(&((ref_sch_t*)ss)->rx, tbs); }

static void fill_cfg(ref_sch_t* s, srslte_pdsch_cfg_t* cfg, uint32_t tbs, uint32_t Qm, uint32_t G, uint32_t rv, int tx)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->grant.nof_tb         = 1;
  cfg->grant.nof_layers     = 1;
  cfg->grant.nof_re         = G / Qm;
  cfg->grant.tb[0].enabled  = true;
  cfg->grant.tb[0].tbs      = (int)tbs;
  cfg->grant.tb[0].rv       = (int)rv;
  cfg->grant.tb[0].mod      = mod_from_qm(Qm);
  cfg->grant.tb[0].nof_bits = G;
  if (tx)
    cfg->softbuffers.tx[0] = &s->tx;
  else
    cfg->softbuffers.rx[0] = &s->rx;
}

/* data: tbs/8 bytes in; e_bytes: ceil(G/8) packed bytes out */
int ref_sch_encode(void* ss, uint32_t tbs, uint32_t Qm, uint32_t G, uint32_t rv, uint8_t* data, uint8_t* e_bytes)
{
  ref_sch_t*         s = ss;
  srslte_pdsch_cfg_t cfg;
  fill_cfg(s, &cfg, tbs, Qm, G, rv, 1);
  if (rv == 0)
    srslte_softbuffer_tx_reset_tbs(&s->tx, tbs);
  return srslte_dlsch_encode2(&s->sch, &cfg, data, e_bytes, 0, 1);
}

/* llr: int16[G] or int8[G]; data_out: >= tbs/8+6 bytes; cb_crc_out: C flags; returns the reference's return code */
int ref_sch_decode(void*    ss,
                   uint32_t tbs,
                   uint32_t Qm,
                   uint32_t G,
                   uint32_t rv,
                   void*    llr,
                   uint8_t* data_out,
                   float*   avg_iter,
                   uint8_t* cb_crc_out,
                   uint32_t max_cb_out)
{
  ref_sch_t*         s = ss;
  srslte_pdsch_cfg_t cfg;
  fill_cfg(s, &cfg, tbs, Qm, G, rv, 0);
  int r = srslte_dlsch_decode2(&s->sch, &cfg, (int16_t*)llr, data_out, 0, 1);
  if (avg_iter)
    *avg_iter = srslte_sch_last_noi(&s->sch);
  if (cb_crc_out) {
    for (uint32_t i = 0; i < max_cb_out && i < s->rx.max_cb; i++)
      cb_crc_out[i] = s->rx.cb_crc[i] ? 1 : 0;
  }
  return r;
}
/* peek at the HARQ soft buffer of one CB (int16 view of SOFTBUFFER_SIZE elements) */
int ref_sch_get_softbuffer(void* ss, uint32_t cb, int16_t* dst, uint32_t n)
{
  ref_sch_t* s = ss;
  if (cb >= s->rx.max_cb)
    return -1;
  memcpy(dst, s->rx.buffer_f[cb], n * sizeof(int16_t));
  return 0;
}

/* --------------------------------------------- CPU baseline runners (pthreads) */
typedef struct {
  int       tid, nthreads;
  uint32_t  ncb, K, nof_iter, repeat;
  int       is8, layout_sb;
  void*     llr; /* ncb x stride elements */
  uint32_t  stride;
  uint8_t*  out; /* ncb x K/8 */
  pthread_barrier_t* bar;
  double    t0, t1;
} c1_arg_t;

static ref_tdec_t* g_c1_dec[256][2];
static void* c1_worker(void* p)
{
  c1_arg_t*   a = p;
  ref_tdec_t* h = g_c1_dec[a->tid][a->layout_sb ? 1 : 0];
  pin_to_cpu(a->tid);
  pthread_barrier_wait(a->bar);
  a->t0 = now_s();
  for (uint32_t r = 0; r < a->repeat; r++) {
    for (uint32_t cb = a->tid; cb < a->ncb; cb += a->nthreads) {
      uint8_t* o = a->out + (size_t)cb * (a->K / 8);
      if (a->is8)
        srslte_tdec_run_all_8bit(&h->td, (int8_t*)a->llr + (size_t)cb * a->stride, o, a->nof_iter, a->K);
      else
        srslte_tdec_run_all(&h->td, (int16_t*)a->llr + (size_t)cb * a->stride, o, a->nof_iter, a->K);
    }
  }
  a->t1 = now_s();
  return NULL;
}

/* Decode ncb code blocks of size K (fixed nof_iter half-iterations, srslte_tdec_run_all[_8bit]) on nthreads
 * threads; returns wall seconds of the decode phase (max over threads). */
double ref_bench_c1(int       nthreads,
                    void*     llr,
                    uint32_t  stride,
                    uint32_t  ncb,
                    uint32_t  K,
                    uint32_t  nof_iter,
                    int       is8,
                    int       layout_sb,
                    uint8_t*  out,
                    uint32_t  repeat)
{
  if (nthreads > 256)
    nthreads = 256;
  for (int t = 0; t < nthreads; t++)
    if (!g_c1_dec[t][layout_sb ? 1 : 0])
      g_c1_dec[t][layout_sb ? 1 : 0] = ref_tdec_new(6144, SRSLTE_TDEC_AUTO, layout_sb ? 0 : 1);
  pthread_t*        th = calloc(nthreads, sizeof(pthread_t));
  c1_arg_t*         a  = calloc(nthreads, sizeof(c1_arg_t));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, nthreads);
  for (int t = 0; t < nthreads; t++) {
    a[t] = (c1_arg_t){t, nthreads, ncb, K, nof_iter, repeat ? repeat : 1, is8, layout_sb, llr, stride, out, &bar, 0, 0};
    pthread_create(&th[t], NULL, c1_worker, &a[t]);
  }
  double t0 = 1e300, t1 = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (a[t].t0 < t0)
      t0 = a[t].t0;
    if (a[t].t1 > t1)
      t1 = a[t].t1;
  }
  pthread_barrier_destroy(&bar);
  free(th);
  free(a);
  return t1 - t0;
}

typedef struct {
  int       tid, nthreads;
  uint32_t  ntb, tbs, Qm, G, rv, max_iter, repeat;
  int       is8;
  void*     llr; /* ntb x G */
  uint8_t*  out; /* ntb x out_stride */
  uint32_t  out_stride;
  int*      rc;
  float*    avg_iter;
  pthread_barrier_t* bar;
  double    t0, t1;
} tb_arg_t;

/* the reference keeps process-global tables behind plain bools (rm_turbo.c:81, turbocoder.c table_initiated), so
 * srslte_sch_init() must not run concurrently: worker objects are created once, sequentially, and cached */
#define MAX_WORKERS 256
static ref_sch_t* g_workers[MAX_WORKERS];
static int        g_workers_is8[MAX_WORKERS];
static ref_sch_t* get_worker(int tid, int is8, uint32_t max_iter)
{
  if (!g_workers[tid]) {
    ref_sch_t* s = calloc(1, sizeof(ref_sch_t));
    srslte_sch_init(&s->sch);
    srslte_softbuffer_rx_init(&s->rx, 100);
    g_workers[tid] = s;
  }
  g_workers[tid]->sch.llr_is_8bit = is8 ? true : false;
  srslte_sch_set_max_noi(&g_workers[tid]->sch, max_iter);
  return g_workers[tid];
}

static void* tb_worker(void* p)
{
  tb_arg_t*  a = p;
  ref_sch_t* s = g_workers[a->tid];
  pin_to_cpu(a->tid);
  pthread_barrier_wait(a->bar);
  a->t0 = now_s();
  for (uint32_t r = 0; r < a->repeat; r++) {
    for (uint32_t tb = a->tid; tb < a->ntb; tb += a->nthreads) {
      srslte_softbuffer_rx_reset_tbs(&s->rx, a->tbs);
      void* llr = a->is8 ? (void*)((int8_t*)a->llr + (size_t)tb * a->G) : (void*)((int16_t*)a->llr + (size_t)tb * a->G);
      srslte_pdsch_cfg_t cfg;
      fill_cfg(s, &cfg, a->tbs, a->Qm, a->G, a->rv, 0);
      a->rc[tb]       = srslte_dlsch_decode2(&s->sch, &cfg, (int16_t*)llr, a->out + (size_t)tb * a->out_stride, 0, 1);
      a->avg_iter[tb] = srslte_sch_last_noi(&s->sch);
    }
  }
  a->t1 = now_s();
  return NULL;
}

/* Decode ntb transport blocks (new transmissions: soft buffer reset before each) with the reference's real
 * srslte_dlsch_decode2 on nthreads threads; returns wall seconds. */
double ref_bench_tb(int      nthreads,
                    void*    llr,
                    uint32_t ntb,
                    uint32_t tbs,
                    uint32_t Qm,
                    uint32_t G,
                    uint32_t rv,
                    uint32_t max_iter,
                    int      is8,
                    uint8_t* out,
                    uint32_t out_stride,
                    int*     rc,
                    float*   avg_iter,
                    uint32_t repeat)
{
  srslte_rm_turbo_gentables();
  if (nthreads > MAX_WORKERS)
    nthreads = MAX_WORKERS;
  for (int t = 0; t < nthreads; t++)
    get_worker(t, is8, max_iter);
  pthread_t*        th = calloc(nthreads, sizeof(pthread_t));
  tb_arg_t*         a  = calloc(nthreads, sizeof(tb_arg_t));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, nthreads);
  for (int t = 0; t < nthreads; t++) {
    a[t] = (tb_arg_t){t, nthreads, ntb, tbs, Qm, G, rv, max_iter, repeat ? repeat : 1, is8, llr, out, out_stride, rc, avg_iter, &bar, 0, 0};
    pthread_create(&th[t], NULL, tb_worker, &a[t]);
  }
  double t0 = 1e300, t1 = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (a[t].t0 < t0)
      t0 = a[t].t0;
    if (a[t].t1 > t1)
      t1 = a[t].t1;
  }
  pthread_barrier_destroy(&bar);
  free(th);
  free(a);
  return t1 - t0;
}


/* --------------------------------------------- mixed-size transport blocks (BASELINE config 3) */
typedef struct {
  int             tid, nthreads;
  uint32_t        ntb, max_iter, repeat;
  int             is8;
  void*           llr;     /* all e-bits, one array */
  const uint64_t* off;     /* element offset of TB i in llr */
  const uint32_t *tbs, *Qm, *G;
  uint8_t*        out;     /* ntb x out_stride */
  uint32_t        out_stride;
  int*            rc;
  float*          avg_iter;
  pthread_barrier_t* bar;
  double          t0, t1;
} mix_arg_t;

static void* mix_worker(void* p)
{
  mix_arg_t* a = p;
  ref_sch_t* s = g_workers[a->tid];
  pin_to_cpu(a->tid);
  pthread_barrier_wait(a->bar);
  a->t0 = now_s();
  for (uint32_t r = 0; r < a->repeat; r++) {
    for (uint32_t tb = a->tid; tb < a->ntb; tb += a->nthreads) {
      srslte_softbuffer_rx_reset_tbs(&s->rx, a->tbs[tb]);
      void* llr = a->is8 ? (void*)((int8_t*)a->llr + a->off[tb]) : (void*)((int16_t*)a->llr + a->off[tb]);
      srslte_pdsch_cfg_t cfg;
      fill_cfg(s, &cfg, a->tbs[tb], a->Qm[tb], a->G[tb], 0, 0);
      a->rc[tb]       = srslte_dlsch_decode2(&s->sch, &cfg, (int16_t*)llr, a->out + (size_t)tb * a->out_stride, 0, 1);
      a->avg_iter[tb] = srslte_sch_last_noi(&s->sch);
    }
  }
  a->t1 = now_s();
  return NULL;
}

/* Decode ntb transport blocks of DIFFERENT sizes (new transmissions, rv 0) with srslte_dlsch_decode2 on nthreads pinned
 * threads; returns wall seconds. */
double ref_bench_tb_mixed(int nthreads, void* llr, const uint64_t* off, const uint32_t* tbs, const uint32_t* Qm, const uint32_t* G, uint32_t ntb,
                          uint32_t max_iter, int is8, uint8_t* out, uint32_t out_stride, int* rc, float* avg_iter, uint32_t repeat)
{
  srslte_rm_turbo_gentables();
  if (nthreads > MAX_WORKERS)
    nthreads = MAX_WORKERS;
  for (int t = 0; t < nthreads; t++)
    get_worker(t, is8, max_iter);
  pthread_t*        th = calloc(nthreads, sizeof(pthread_t));
  mix_arg_t*        a  = calloc(nthreads, sizeof(mix_arg_t));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, nthreads);
  for (int t = 0; t < nthreads; t++) {
    a[t] = (mix_arg_t){t, nthreads, ntb, max_iter, repeat ? repeat : 1, is8, llr, off, tbs, Qm, G, out, out_stride, rc, avg_iter, &bar, 0, 0};
    pthread_create(&th[t], NULL, mix_worker, &a[t]);
  }
  double t0 = 1e300, t1 = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (a[t].t0 < t0)
      t0 = a[t].t0;
    if (a[t].t1 > t1)
      t1 = a[t].t1;
  }
  pthread_barrier_destroy(&bar);
  free(th);
  free(a);
  return t1 - t0;
}

/* --------------------------------------------- symbols in: soft demodulation + descrambling + decode per transport block
 * (the chain of lib/src/phy/phch/pdsch.c:832-859 without the equaliser), all host threads, pinned */
typedef struct {
  int          tid, nthreads;
  uint32_t     ntb, tbs, mod, Qm, nsym, max_iter, repeat, c_init;
  int          is8;
  const float* sym; /* ntb x nsym complex */
  uint8_t*     out;
  uint32_t     out_stride;
  int*         rc;
  float*       avg_iter;
  pthread_barrier_t* bar;
  double       t0, t1;
} symb_arg_t;

static void* symb_worker(void* p)
{
  symb_arg_t* a = p;
  ref_sch_t*  s = g_workers[a->tid];
  const uint32_t G = a->nsym * a->Qm;
  srslte_sequence_t seq;
  memset(&seq, 0, sizeof(seq));
  srslte_sequence_LTE_pr(&seq, G, a->c_init);
  void* llr = NULL;
  if (posix_memalign(&llr, 64, (size_t)G * 2 + 64))
    return NULL;
  pin_to_cpu(a->tid);
  pthread_barrier_wait(a->bar);
  a->t0 = now_s();
  for (uint32_t r = 0; r < a->repeat; r++) {
    for (uint32_t tb = a->tid; tb < a->ntb; tb += a->nthreads) {
      const cf_t* sy = (const cf_t*)(a->sym + (size_t)tb * a->nsym * 2);
      if (a->is8) {
        srslte_demod_soft_demodulate_b((srslte_mod_t)a->mod, sy, (int8_t*)llr, (int)a->nsym);
        srslte_scrambling_sb_offset(&seq, (int8_t*)llr, 0, (int)G);
      } else {
        srslte_demod_soft_demodulate_s((srslte_mod_t)a->mod, sy, (int16_t*)llr, (int)a->nsym);
        srslte_scrambling_s_offset(&seq, (int16_t*)llr, 0, (int)G);
      }
      srslte_softbuffer_rx_reset_tbs(&s->rx, a->tbs);
      srslte_pdsch_cfg_t cfg;
      fill_cfg(s, &cfg, a->tbs, a->Qm, G, 0, 0);
      a->rc[tb]       = srslte_dlsch_decode2(&s->sch, &cfg, (int16_t*)llr, a->out + (size_t)tb * a->out_stride, 0, 1);
      a->avg_iter[tb] = srslte_sch_last_noi(&s->sch);
    }
  }
  a->t1 = now_s();
  free(llr);
  srslte_sequence_free(&seq);
  return NULL;
}

double ref_bench_tb_symbols(int nthreads, const float* sym, uint32_t ntb, uint32_t nsym, uint32_t mod, uint32_t tbs, uint32_t c_init, uint32_t max_iter,
                            int is8, uint8_t* out, uint32_t out_stride, int* rc, float* avg_iter, uint32_t repeat)
{
  srslte_rm_turbo_gentables();
  if (nthreads > MAX_WORKERS)
    nthreads = MAX_WORKERS;
  for (int t = 0; t < nthreads; t++)
    get_worker(t, is8, max_iter);
  const uint32_t Qm = mod == 0 ? 1 : 2 * mod;
  pthread_t*        th = calloc(nthreads, sizeof(pthread_t));
  symb_arg_t*       a  = calloc(nthreads, sizeof(symb_arg_t));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, nthreads);
  for (int t = 0; t < nthreads; t++) {
    a[t] = (symb_arg_t){t, nthreads, ntb, tbs, mod, Qm, nsym, max_iter, repeat ? repeat : 1, c_init, is8, sym, out, out_stride, rc, avg_iter, &bar, 0, 0};
    pthread_create(&th[t], NULL, symb_worker, &a[t]);
  }
  double t0 = 1e300, t1 = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (a[t].t0 < t0)
      t0 = a[t].t0;
    if (a[t].t1 > t1)
      t1 = a[t].t1;
  }
  pthread_barrier_destroy(&bar);
  free(th);
  free(a);
  return t1 - t0;
}

/* --------------------------------------------- per-TTI latency on ONE pinned core (us per TTI, p50 / p99;
 * the reference times the same call, lib/src/phy/phch/pdsch.c:921-924, 1061-1065).  lat_us[i] = wall time of call i. */
int ref_latency_tb(void* llr, uint32_t ntb, uint32_t tbs, uint32_t Qm, uint32_t G, uint32_t max_iter, int is8, uint32_t n_calls, double* lat_us)
{
  srslte_rm_turbo_gentables();
  ref_sch_t* s = get_worker(0, is8, max_iter);
  uint8_t*   out = calloc(tbs / 8 + 8 + 768, 1);
  cpu_set_t  saved;
  const int  have = save_affinity(&saved);
  pin_to_cpu(0);
  for (uint32_t i = 0; i < n_calls; i++) {
    uint32_t tb = i % ntb;
    void*    l  = is8 ? (void*)((int8_t*)llr + (size_t)tb * G) : (void*)((int16_t*)llr + (size_t)tb * G);
    double   t0 = now_s();
    srslte_softbuffer_rx_reset_tbs(&s->rx, tbs);
    srslte_pdsch_cfg_t cfg;
    fill_cfg(s, &cfg, tbs, Qm, G, 0, 0);
    srslte_dlsch_decode2(&s->sch, &cfg, (int16_t*)l, out, 0, 1);
    lat_us[i] = 1e6 * (now_s() - t0);
  }
  restore_affinity(&saved, have);
  free(out);
  return 0;
}
/* one "TTI" of the code-block workload: cbs_per_tti blocks of size K through srslte_tdec_run_all, one pinned core */
int ref_latency_c1(void* llr, uint32_t stride, uint32_t ncb, uint32_t K, uint32_t nof_iter, uint32_t cbs_per_tti, uint32_t n_calls, double* lat_us)
{
  if (!g_c1_dec[0][0])
    g_c1_dec[0][0] = ref_tdec_new(6144, SRSLTE_TDEC_AUTO, 1);
  ref_tdec_t* h   = g_c1_dec[0][0];
  uint8_t*    out = calloc(K / 8 + 16, 1);
  cpu_set_t   saved;
  const int   have = save_affinity(&saved);
  pin_to_cpu(0);
  for (uint32_t i = 0; i < n_calls; i++) {
    double t0 = now_s();
    for (uint32_t c = 0; c < cbs_per_tti; c++) {
      uint32_t cb = (i * cbs_per_tti + c) % ncb;
      srslte_tdec_run_all(&h->td, (int16_t*)llr + (size_t)cb * stride, out, nof_iter, K);
    }
    lat_us[i] = 1e6 * (now_s() - t0);
  }
  restore_affinity(&saved, have);
  free(out);
  return 0;
}

/* ------------------------------------------------------------------ soft demodulation + descrambling */
int ref_demod_s(int mod, const float* symbols, int16_t* llr, int nsymbols)
{
  return srslte_demod_soft_demodulate_s((srslte_mod_t)mod, (const cf_t*)symbols, llr, nsymbols);
}
int ref_demod_b(int mod, const float* symbols, int8_t* llr, int nsymbols)
{
  return srslte_demod_soft_demodulate_b((srslte_mod_t)mod, (const cf_t*)symbols, llr, nsymbols);
}
/* pseudo-random sequence of TS 36.211 7.2 as the reference stores it: packed bytes (c_bytes, MSB first) */
int ref_sequence_bytes(uint32_t c_init, uint32_t len, uint8_t* out)
{
  srslte_sequence_t seq;
  memset(&seq, 0, sizeof(seq));
  if (srslte_sequence_LTE_pr(&seq, len, c_init))
    return -1;
  memcpy(out, seq.c_bytes, (len + 7) / 8);
  srslte_sequence_free(&seq);
  return 0;
}
int ref_descramble_s(uint32_t c_init, int16_t* data, int len)
{
  srslte_sequence_t seq;
  memset(&seq, 0, sizeof(seq));
  if (srslte_sequence_LTE_pr(&seq, len, c_init))
    return -1;
  srslte_scrambling_s_offset(&seq, data, 0, len);
  srslte_sequence_free(&seq);
  return 0;
}
int ref_descramble_b(uint32_t c_init, int8_t* data, int len)
{
  srslte_sequence_t seq;
  memset(&seq, 0, sizeof(seq));
  if (srslte_sequence_LTE_pr(&seq, len, c_init))
    return -1;
  srslte_scrambling_sb_offset(&seq, data, 0, len);
  srslte_sequence_free(&seq);
  return 0;
}

/* ------------------------------------------------------------------ PUSCH: srslte_ulsch_encode / srslte_ulsch_decode
 * p = {tbs, Qm, L_prb, nof_symb, rv, nof_ack, ri_len, cqi (0 none, 1 wideband 4 bit, 2 higher-layer subband N=9 22 bit),
 *      I_offset_ack, I_offset_ri, I_offset_cqi}; nb_q = L_prb * 12 * nof_symb * Qm */
#include "srslte/phy/phch/pusch_cfg.h"
static void fill_ul_cfg(ref_sch_t* s, srslte_pusch_cfg_t* cfg, const uint32_t* p, int tx)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->grant.L_prb       = p[2];
  cfg->grant.nof_symb    = p[3];
  cfg->grant.nof_re      = p[2] * 12 * p[3];
  cfg->grant.tb.enabled  = true;
  cfg->grant.tb.tbs      = (int)p[0];
  cfg->grant.tb.rv       = (int)p[4];
  cfg->grant.tb.mod      = mod_from_qm(p[1]);
  cfg->grant.tb.nof_bits = cfg->grant.nof_re * p[1];
  cfg->uci_cfg.ack[0].nof_acks = p[5];
  cfg->uci_cfg.cqi.ri_len      = p[6];
  if (p[7]) {
    cfg->uci_cfg.cqi.data_enable = true;
    cfg->uci_cfg.cqi.type        = p[7] == 1 ? SRSLTE_CQI_TYPE_WIDEBAND : SRSLTE_CQI_TYPE_SUBBAND_HL;
    cfg->uci_cfg.cqi.N           = 9;
  }
  cfg->uci_offset.I_offset_ack = p[8];
  cfg->uci_offset.I_offset_ri  = p[9];
  cfg->uci_offset.I_offset_cqi = p[10];
  if (tx)
    cfg->softbuffers.tx = &s->tx;
  else
    cfg->softbuffers.rx = &s->rx;
}
static void fill_uci_value(srslte_uci_value_t* v, const uint32_t* p, uint32_t seed)
{
  memset(v, 0, sizeof(*v));
  for (uint32_t i = 0; i < p[5] && i < SRSLTE_UCI_MAX_ACK_BITS; i++)
    v->ack.ack_value[i] = (seed >> i) & 1u;
  v->ri = (seed >> 7) & 1u;
  if (p[7] == 1) {
    v->cqi.wideband.wideband_cqi = (seed >> 8) & 15u;
  } else if (p[7] == 2) {
    v->cqi.subband_hl.wideband_cqi_cw0     = (seed >> 8) & 15u;
    v->cqi.subband_hl.subband_diff_cqi_cw0 = (seed * 2654435761u) & 0x3FFFFu;
  }
}
/* data: tbs/8 bytes; q_bytes out: nb_q/8 packed bits of the interleaved, UCI-multiplexed codeword (before scrambling) */
int ref_ulsch_encode(void* ss, const uint32_t* p, uint32_t uci_seed, uint8_t* data, uint8_t* q_bytes)
{
  ref_sch_t*         s = ss;
  srslte_pusch_cfg_t cfg;
  srslte_uci_value_t uci;
  fill_ul_cfg(s, &cfg, p, 1);
  fill_uci_value(&uci, p, uci_seed);
  if (p[4] == 0)
    srslte_softbuffer_tx_reset_tbs(&s->tx, p[0]);
  uint32_t nb_q   = cfg.grant.tb.nof_bits;
  uint8_t* g_bits = srslte_vec_u8_malloc(nb_q / 8 + 64);
  memset(g_bits, 0, nb_q / 8 + 64);
  memset(q_bytes, 0, nb_q / 8);
  int r = srslte_ulsch_encode(&s->sch, &cfg, data, &uci, g_bits, q_bytes);
  free(g_bits);
  return r;
}
/* q_llr: int16[nb_q] descrambled LLRs, modified in place as the reference does; c_seq: nb_q unpacked scrambling bits;
 * g_bits out: int16[nb_q] (pre-filled by the caller; the reference leaves the last Q'_ri*Qm untouched);
 * out = {ret, ack_value[0..3], ack.valid, ri, cqi.data_crc, first cqi field} */
int ref_ulsch_decode(void* ss, const uint32_t* p, int16_t* q_llr, uint8_t* c_seq, int16_t* g_bits, uint8_t* data, int32_t* out)
{
  ref_sch_t*         s = ss;
  srslte_pusch_cfg_t cfg;
  srslte_uci_value_t uci;
  fill_ul_cfg(s, &cfg, p, 0);
  memset(&uci, 0, sizeof(uci));
  int r  = srslte_ulsch_decode(&s->sch, &cfg, q_llr, g_bits, c_seq, data, &uci);
  out[0] = r;
  for (int i = 0; i < 4; i++)
    out[1 + i] = uci.ack.ack_value[i];
  out[5] = uci.ack.valid;
  out[6] = uci.ri;
  out[7] = uci.cqi.data_crc;
  out[8] = p[7] == 2 ? uci.cqi.subband_hl.wideband_cqi_cw0 : uci.cqi.wideband.wideband_cqi;
  return r;
}
