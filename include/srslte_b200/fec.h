/*
 * srslte_b200/fec.h -- drop-in C declarations for the LTE turbo-decode hot path, served by the B200 engine.
 *
 * Every symbol below keeps the NAME, ARGUMENT MEANING and RETURN CONVENTION of the srsLTE 20.10.1 function it
 * replaces (file:line given per symbol, paths relative to the reference's lib/), so that the host C code of
 * lib/src/phy/phch/sch.c and pssch.c recompiles against this header unchanged.  All arithmetic runs on the GPU
 * (libsrslte_fec_b200.so); pointers are ordinary HOST pointers exactly as in the reference.  There is no CPU
 * fallback: without a usable sm_100 device the init functions return SRSLTE_ERROR.
 *
 * For throughput use the batched entry points in srslte_b200/batch.h; the per-code-block calls here are thin
 * wrappers that submit a batch of one and wait for it.
 */
#ifndef SRSLTE_B200_FEC_H
#define SRSLTE_B200_FEC_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SRSLTE_API
#define SRSLTE_API __attribute__((visibility("default"))) /* lib/include/srslte/config.h:32-46 */
#endif
#ifndef SRSLTE_SUCCESS
#define SRSLTE_SUCCESS 0 /* lib/include/srslte/config.h:57-59 */
#define SRSLTE_ERROR -1
#define SRSLTE_ERROR_INVALID_INPUTS -2
#endif

/* ---------------------------------------------------------------- cbsegm.h:30-52 (host-side parameter calc) */
#define SRSLTE_NOF_TC_CB_SIZES 188
typedef struct SRSLTE_API {
  uint32_t F;
  uint32_t C;
  uint32_t K1;
  uint32_t K2;
  uint32_t K1_idx;
  uint32_t K2_idx;
  uint32_t C1;
  uint32_t C2;
  uint32_t tbs;
} srslte_cbsegm_t;
SRSLTE_API int  srslte_cbsegm(srslte_cbsegm_t* s, uint32_t tbs);  /* cbsegm.c:48 */
SRSLTE_API int  srslte_cbsegm_cbsize(uint32_t index);            /* cbsegm.c:132 */
SRSLTE_API bool srslte_cbsegm_cbsize_isvalid(uint32_t size);     /* cbsegm.c:147 */
SRSLTE_API int  srslte_cbsegm_cbindex(uint32_t long_cb);         /* cbsegm.c:114 */

/* ---------------------------------------------------------------- crc.h:36-74 */
typedef struct SRSLTE_API {
  uint64_t table[256];
  int      polynom;
  int      order;
  uint64_t crcinit;
  uint64_t crcmask;
  uint64_t crchighbit;
  uint32_t srslte_crc_out;
} srslte_crc_t; /* field layout kept: crc.h:36-44 */
#ifndef SRSLTE_LTE_CRC24A
#define SRSLTE_LTE_CRC24A 0x1864CFB /* phy_common.h:71-74 */
#define SRSLTE_LTE_CRC24B 0X1800063
#define SRSLTE_LTE_CRC16 0x11021
#define SRSLTE_LTE_CRC8 0x19B
#endif
SRSLTE_API int      srslte_crc_init(srslte_crc_t* h, uint32_t srslte_crc_poly, int srslte_crc_order); /* crc.c:73 */
SRSLTE_API int      srslte_crc_set_init(srslte_crc_t* h, uint64_t init_value);                        /* crc.c:61 */
SRSLTE_API uint32_t srslte_crc_checksum_byte(srslte_crc_t* h, uint8_t* data, int len);                /* crc.c:143 */
SRSLTE_API uint32_t srslte_crc_checksum(srslte_crc_t* h, uint8_t* data, int len);                     /* crc.c:102 */
SRSLTE_API uint32_t srslte_crc_attach_byte(srslte_crc_t* h, uint8_t* data, int len);                  /* crc.c:159 */
SRSLTE_API uint32_t srslte_crc_attach(srslte_crc_t* h, uint8_t* data, int len);                       /* crc.c:172 */
/* crc.h:56-70: the two header-only helpers of the reference.  They stay host code here as they are there: their one caller
 * is the TRANSMIT-side encoder srslte_tcod_encode_lut (turbocoder.c:190-196), which is outside the replaced path and keeps
 * running on the CPU; they only touch the srslte_crc_t fields kept above (table, order, crcinit, crcmask). */
static inline void srslte_crc_checksum_put_byte(srslte_crc_t* h, uint8_t byte)
{
  const uint64_t c  = h->crcinit;
  const uint64_t ix = ((c >> (h->order - 8)) & 0xff) ^ byte; /* polynomial order 8, 16, 24 or 32 */
  h->crcinit        = (c << 8) ^ h->table[ix];
}
static inline uint64_t srslte_crc_checksum_get(srslte_crc_t* h)
{
  return h->crcinit & h->crcmask;
}

/* ---------------------------------------------------------------- turbodecoder_impl.h:28-38 */
typedef enum SRSLTE_API {
  SRSLTE_TDEC_AUTO = 0,
  SRSLTE_TDEC_GENERIC,
  SRSLTE_TDEC_SSE,         /* state-parallel un-windowed variant: NOT provided (manual-only in the reference) */
  SRSLTE_TDEC_SSE_WINDOW,  /* 8 lanes, int16  */
  SRSLTE_TDEC_NEON_WINDOW, /* NOT provided */
  SRSLTE_TDEC_AVX_WINDOW,  /* 16 lanes, int16 */
  SRSLTE_TDEC_SSE8_WINDOW, /* 16 lanes, int8  */
  SRSLTE_TDEC_AVX8_WINDOW, /* 32 lanes, int8  */
  SRSLTE_TDEC_NOF_IMP
} srslte_tdec_impl_type_t;

/* ---------------------------------------------------------------- turbodecoder.h:63-121
 * The reference embeds this struct by value (sch.h:68, pssch.h:85) and no caller touches its fields; the fields
 * that describe host scratch memory are replaced by an opaque engine handle. */
#define SRSLTE_TCOD_RATE 3        /* turbodecoder.h:41-47 */
#define SRSLTE_TCOD_TOTALTAIL 12
#define SRSLTE_TCOD_MAX_LEN_CB 6144
#define SRSLTE_TDEC_EXPECT_INPUT_SB 1
typedef struct SRSLTE_API {
  uint32_t                max_long_cb;
  void*                   b200_engine; /* device engine owning tables, stream and the decoder state */
  bool                    force_not_sb;
  srslte_tdec_impl_type_t dec_type;
  uint32_t                current_long_cb;
  int                     current_cbidx;
  int                     n_iter;
} srslte_tdec_t;

SRSLTE_API int      srslte_tdec_init(srslte_tdec_t* h, uint32_t max_long_cb);                                        /* turbodecoder.c:129 */
SRSLTE_API int      srslte_tdec_init_manual(srslte_tdec_t* h, uint32_t max_long_cb, srslte_tdec_impl_type_t dec_type); /* :151 */
SRSLTE_API void     srslte_tdec_free(srslte_tdec_t* h);                                                              /* :319 */
SRSLTE_API void     srslte_tdec_force_not_sb(srslte_tdec_t* h);                                                      /* :365 */
SRSLTE_API int      srslte_tdec_new_cb(srslte_tdec_t* h, uint32_t long_cb);                                          /* :511 */
SRSLTE_API int      srslte_tdec_get_nof_iterations(srslte_tdec_t* h);                                                /* :580 */
SRSLTE_API uint32_t srslte_tdec_autoimp_get_subblocks(uint32_t long_cb);                                             /* :381 */
SRSLTE_API uint32_t srslte_tdec_autoimp_get_subblocks_8bit(uint32_t long_cb);                                        /* :410 */
SRSLTE_API void     srslte_tdec_iteration(srslte_tdec_t* h, int16_t* input, uint8_t* output);                        /* :528 */
SRSLTE_API int      srslte_tdec_run_all(srslte_tdec_t* h, int16_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb); /* :537 */
SRSLTE_API void     srslte_tdec_iteration_8bit(srslte_tdec_t* h, int8_t* input, uint8_t* output);                    /* :552 */
SRSLTE_API int      srslte_tdec_run_all_8bit(srslte_tdec_t* h, int8_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb); /* :560 */

/* ---------------------------------------------------------------- rm_turbo.h:57-89 (receive side) */
SRSLTE_API void srslte_rm_turbo_gentables(void);   /* rm_turbo.c:280: tables are generated when an engine is created */
SRSLTE_API void srslte_rm_turbo_free_tables(void); /* rm_turbo.c:323 */
SRSLTE_API int  srslte_rm_turbo_rx_lut(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx);       /* :397 */
SRSLTE_API int  srslte_rm_turbo_rx_lut_(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx, bool enable_input_tdec); /* :410 */
SRSLTE_API int  srslte_rm_turbo_rx_lut_8bit(int8_t* input, int8_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx);    /* :456 */

/* ---------------------------------------------------------------- softbuffer.h:37-60
 * Field names and order of the reference are kept, so decode_tb_cb (sch.c:363-488), which dereferences buffer_f[] and
 * data[] (sch.c:404,409,422,424,466,481), recompiles unchanged: buffer_f[i] (SOFTBUFFER_SIZE int16) and data[i] (768 bytes)
 * are ordinary HOST arrays that the per-code-block symbols above read and write.  The batched path (srslte_b200_decode_tb,
 * batch.h) keeps the HARQ state of the same soft buffer in GPU memory behind b200_softbuffer and never touches the host
 * arrays; a soft buffer must be driven through ONE of the two paths between resets. */
#define SOFTBUFFER_SIZE 18600
typedef struct SRSLTE_API {
  uint32_t  max_cb;
  int16_t** buffer_f;
  uint8_t** data;
  bool*     cb_crc;
  bool      tb_crc;
  void*     b200_softbuffer; /* device-resident HARQ state of the batched path */
  bool      b200_host_dirty; /* the per-code-block path wrote into buffer_f[] since the last reset */
} srslte_softbuffer_rx_t;
SRSLTE_API int  srslte_softbuffer_rx_init(srslte_softbuffer_rx_t* q, uint32_t nof_prb);       /* softbuffer.c:41 */
SRSLTE_API void srslte_softbuffer_rx_reset(srslte_softbuffer_rx_t* p);                        /* :128 */
SRSLTE_API void srslte_softbuffer_rx_reset_tbs(srslte_softbuffer_rx_t* q, uint32_t tbs);      /* :122 */
SRSLTE_API void srslte_softbuffer_rx_reset_cb(srslte_softbuffer_rx_t* q, uint32_t nof_cb);    /* :133 */
SRSLTE_API void srslte_softbuffer_rx_free(srslte_softbuffer_rx_t* p);                         /* :97  */
/* transmit-side soft buffer (softbuffer.h:45-48, softbuffer.c:157-245): host memory management only, kept so that a build
 * that drops the reference's softbuffer.c still links its CPU encoder (encode_tb_off, sch.c:235-349) */
typedef struct SRSLTE_API {
  uint32_t  max_cb;
  uint8_t** buffer_b;
} srslte_softbuffer_tx_t;
SRSLTE_API int  srslte_softbuffer_tx_init(srslte_softbuffer_tx_t* q, uint32_t nof_prb);       /* softbuffer.c:157 */
SRSLTE_API void srslte_softbuffer_tx_reset(srslte_softbuffer_tx_t* p);                        /* :209 */
SRSLTE_API void srslte_softbuffer_tx_reset_tbs(srslte_softbuffer_tx_t* q, uint32_t tbs);      /* :203 */
SRSLTE_API void srslte_softbuffer_tx_reset_cb(srslte_softbuffer_tx_t* q, uint32_t nof_cb);    /* :214 */
SRSLTE_API void srslte_softbuffer_tx_free(srslte_softbuffer_tx_t* p);                         /* :190 */

/* ---------------------------------------------------------------- sch.c:503-570 decode_tb as one call
 * What decode_tb (static in sch.c) becomes: same arguments (q's relevant state passed explicitly), same return
 * values, same bytes in `data`.  avg_iterations receives what srslte_sch_last_noi() would report. */
SRSLTE_API int srslte_b200_decode_tb(srslte_softbuffer_rx_t* softbuffer,
                                     srslte_cbsegm_t*        cb_segm,
                                     uint32_t                Qm,
                                     uint32_t                rv,
                                     uint32_t                nof_e_bits,
                                     void*                   e_bits,
                                     bool                    llr_is_8bit,
                                     uint32_t                max_iterations,
                                     uint8_t*                data,
                                     float*                  avg_iterations);

#ifdef __cplusplus
}
#endif
#endif
