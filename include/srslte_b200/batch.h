/*
 * srslte_b200/batch.h -- the NEW batched C-ABI entry points of the B200 turbo-decode engine.
 *
 * These are what the existing C host code calls instead of looping over code blocks one at a time
 * (reference: the per-code-block loop decode_tb_cb, lib/src/phy/phch/sch.c:363-488, driven from
 * srslte_dlsch_decode2 sch.c:577-606 / srslte_ulsch_decode sch.c:1105-1180).  Plain C: pointers and sizes
 * only, no CUDA or torch types.  All work runs on the GPU; there is no CPU fallback -- every entry point
 * returns SRSLTE_B200_ERROR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef SRSLTE_B200_BATCH_H
#define SRSLTE_B200_BATCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRSLTE_B200_API __attribute__((visibility("default")))

#define SRSLTE_B200_SUCCESS 0
#define SRSLTE_B200_ERROR -1
#define SRSLTE_B200_ERROR_INVALID_INPUTS -2
#define SRSLTE_B200_ERROR_NO_DEVICE -3

#define SRSLTE_B200_MAX_CODEBLOCKS 32 /* SRSLTE_MAX_CODEBLOCKS, lib/include/srslte/phy/common/phy_common.h:63 */

/* flags for the batch calls */
#define SRSLTE_B200_IN_DEVICE 1u  /* LLR / e_bits pointers are device pointers (already resident in HBM) */
#define SRSLTE_B200_OUT_DEVICE 2u /* output byte pointers are device pointers */
#define SRSLTE_B200_SEQ_DEVICE 4u /* srslte_b200_demod_descramble: the scrambling sequences are device pointers (a receiver uploads
                                   * the sequences of its RNTIs once) while the symbols still come from the host */

#define SRSLTE_B200_UCI_DEFERRED 8u /* srslte_b200_ulsch_deinterleave with SRSLTE_B200_OUT_DEVICE: do not wait for the UCI LLRs --
                                     * ack_llr / ri_llr / cqi_llr are filled by the next srslte_b200_wait() on the context (which a
                                     * srslte_b200_decode_tbs_submit of the subframes' data normally precedes), so the call only enqueues */

typedef struct srslte_b200_ctx srslte_b200_ctx_t;             /* one engine instance = one GPU + one stream */
typedef struct srslte_b200_softbuffer srslte_b200_softbuffer_t; /* device-resident srslte_softbuffer_rx_t */

/* Create / destroy an engine on CUDA device `device` (-1: current device).  Generates and uploads all tables
 * (QPP: tc_interl_lte.c:69-109 for 188 K x {1,8,16,32} lanes; rate de-matching: rm_turbo.c:280-321). */
SRSLTE_B200_API int  srslte_b200_ctx_create(srslte_b200_ctx_t** ctx, int device);
SRSLTE_B200_API void srslte_b200_ctx_destroy(srslte_b200_ctx_t* ctx);
SRSLTE_B200_API const char* srslte_b200_last_error(void);

/* Pinned host memory helpers (page-locked buffers make the H2D/D2H copies asynchronous). */
SRSLTE_B200_API void* srslte_b200_host_alloc(uint64_t bytes);
/* write-combined page-locked memory for INPUT staging (LLRs the host only writes, sequentially): the device reads it
 * without snooping the CPU caches; CPU reads from it are very slow.  Free with srslte_b200_host_free. */
SRSLTE_B200_API void* srslte_b200_host_alloc_wc(uint64_t bytes);
SRSLTE_B200_API void  srslte_b200_host_free(void* p);
SRSLTE_B200_API void* srslte_b200_device_alloc(uint64_t bytes);
SRSLTE_B200_API void  srslte_b200_device_free(void* p);
SRSLTE_B200_API int   srslte_b200_memcpy_h2d(srslte_b200_ctx_t* ctx, void* dst, const void* src, uint64_t bytes);
SRSLTE_B200_API int   srslte_b200_memcpy_d2h(srslte_b200_ctx_t* ctx, void* dst, const void* src, uint64_t bytes);

/* ---- batch of independent code blocks, fixed number of half-iterations, no CRC:
 * the batched form of srslte_tdec_run_all[_8bit] (turbodecoder.c:537-578).
 *   llr        nof_cb blocks, llr_stride elements apart, int16 (llr_bits 16) or int8 (llr_bits 8)
 *   input_sb   0: standard order 3k+s, 3K+12 values (what srslte_tdec_force_not_sb selects, turbodecoder.c:365);
 *              1: the lane layout written by srslte_rm_turbo_rx_lut (3(K+32)+12 values, rm_turbo.c:263-277)
 *   dec_type   srslte_tdec_impl_type_t value (0 = AUTO)
 *   out        nof_cb x K/8 bytes, MSB first
 */
typedef struct {
  uint32_t K;
  uint32_t nof_cb;
  uint32_t nof_iterations;
  uint32_t llr_bits;
  uint32_t llr_stride;
  uint32_t input_sb;
  uint32_t dec_type;
} srslte_b200_cb_batch_t;

SRSLTE_B200_API int
srslte_b200_tdec_batch(srslte_b200_ctx_t* ctx, const srslte_b200_cb_batch_t* cfg, const void* llr, uint8_t* out, uint32_t flags);

/* ---- batch of transport blocks: the batched form of decode_tb (sch.c:503-570) for many TBs / TTIs at once. */
typedef struct {
  /* inputs */
  const void*               e_bits;     /* int16[nof_e_bits] or int8[nof_e_bits] (llr_is_8bit) */
  uint32_t                  nof_e_bits; /* G */
  uint32_t                  tbs;        /* transport block size in bits */
  uint32_t                  Qm;         /* modulation order x layers, as decode_tb's Qm */
  uint32_t                  rv;
  srslte_b200_softbuffer_t* softbuffer; /* HARQ state, or NULL for a one-shot new transmission */
  uint8_t*                  data;       /* >= tbs/8 + 6 bytes */
  /* outputs */
  int32_t  ret;                                   /* decode_tb's return value: 0, -1 (CRC), -2 (invalid) */
  float    avg_iterations;                        /* srslte_sch_last_noi */
  uint8_t  cb_crc[SRSLTE_B200_MAX_CODEBLOCKS];    /* softbuffer->cb_crc after the call */
  uint8_t  cb_noi[SRSLTE_B200_MAX_CODEBLOCKS];    /* half-iterations spent per code block in this call */
  uint32_t nof_cb;
} srslte_b200_tb_t;

SRSLTE_B200_API int srslte_b200_decode_tbs(srslte_b200_ctx_t* ctx,
                                           srslte_b200_tb_t*  tbs,
                                           uint32_t           nof_tb,
                                           int                llr_is_8bit,
                                           uint32_t           max_iterations,
                                           uint32_t           flags);

/* asynchronous split of the two calls above: submit enqueues copies + kernels on the engine's stream and returns;
 * wait blocks until they are done and fills the outputs.  One submit may be outstanding per context. */
SRSLTE_B200_API int srslte_b200_tdec_batch_submit(srslte_b200_ctx_t* ctx, const srslte_b200_cb_batch_t* cfg, const void* llr, uint8_t* out, uint32_t flags);
SRSLTE_B200_API int srslte_b200_decode_tbs_submit(srslte_b200_ctx_t* ctx, srslte_b200_tb_t* tbs, uint32_t nof_tb, int llr_is_8bit, uint32_t max_iterations, uint32_t flags);
SRSLTE_B200_API int srslte_b200_wait(srslte_b200_ctx_t* ctx);

/* Optional scheduling hint for the NEXT srslte_b200_decode_tbs[_submit] on this context: one number per transport block of that
 * call, larger = expected to need more half-iterations (the noise estimate of the channel estimator, srslte_chest_dl_res_t::
 * noise_estimate / the inverse of snr_db, is what a receiver has at pdsch.c:859; the MCS works too).  Code blocks of equal size
 * are then grouped by hint, so that the blocks sharing a warp of the persistent kernel stop at about the same half-iteration
 * instead of one slow block keeping the finished ones aboard.  Results do not depend on the hints in any way; without the call
 * (or with nof_tb different from the batch's) the engine takes an estimate of its own from the e-bits of each block (option
 * "auto_group", default 1; sizes whose neighbouring blocks belong to one transport block anyway keep submission order). */
SRSLTE_B200_API int srslte_b200_set_tb_hints(srslte_b200_ctx_t* ctx, const float* hints, uint32_t nof_tb);

/* ---- soft demodulation + descrambling on the device: the batched form of
 * srslte_demod_soft_demodulate_s / _b (lib/src/phy/modem/demod_soft.c:896-945) followed by
 * srslte_scrambling_s_offset / _sb_offset (lib/src/phy/scrambling/scrambling.c:43-53) -- the two steps between the
 * equaliser and decode_tb in lib/src/phy/phch/pdsch.c:832-852.  Integer results are those of the reference's x86 (SSE)
 * build, including the position-dependent rounding of its SIMD bodies and scalar tails.
 *   symbols         cf_t[nof_symbols] (interleaved re, im floats) of one codeword
 *   mod             srslte_mod_t: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM
 *   scramble_bytes  srslte_sequence_t::c_bytes of the codeword (bit i of the sequence = bit 7-(i%8) of byte i/8,
 *                   >= nof_symbols*Qm bits), or NULL for no descrambling
 *   e_bits          out: int16 or int8 [nof_symbols * Qm], 4-byte aligned for full store width
 * flags: SRSLTE_B200_IN_DEVICE  -> symbols and scramble_bytes are device pointers
 *        SRSLTE_B200_OUT_DEVICE -> e_bits are device pointers and the call returns once the work is enqueued on the
 *                                  context's stream (a following srslte_b200_decode_tbs*(..., SRSLTE_B200_IN_DEVICE) on the
 *                                  same context consumes them in order); otherwise the call returns with e_bits filled. */
typedef struct {
  const void*    symbols;
  uint32_t       nof_symbols;
  uint32_t       mod;
  const uint8_t* scramble_bytes;
  void*          e_bits;
  const float*   csi; /* channel state information per symbol (q->csi[cw] of pdsch.c) or NULL: when given, csi_correction
                       * (lib/src/phy/phch/pdsch.c:628-741, what cfg->csi_enable turns on) runs between the demodulator and the
                       * descrambler, with the integer results of the reference's x86 build; a host pointer unless
                       * SRSLTE_B200_IN_DEVICE */
} srslte_b200_demod_t;

SRSLTE_B200_API int srslte_b200_demod_descramble(srslte_b200_ctx_t* ctx, const srslte_b200_demod_t* cws, uint32_t nof_cw, int llr_is_8bit, uint32_t flags);

/* pseudo-random sequence of TS 36.211 7.2 packed like srslte_sequence_t::c_bytes (srslte_sequence_LTE_pr,
 * lib/src/phy/common/sequence.c); out: (len + 7) / 8 bytes.  Host helper for callers that do not hold the sequence. */
SRSLTE_B200_API void srslte_b200_sequence_bytes(uint32_t c_init, uint32_t len, uint8_t* out);

/* ---- PUSCH pre-steps: the data movement srslte_ulsch_decode (lib/src/phy/phch/sch.c:1105-1180) performs
 * between the descrambler and decode_tb, batched over transport blocks:
 *   - the LLRs at the HARQ-ACK and RI resource elements (uci_ulsch_interleave_{ack,ri}_gen, lib/src/phy/phch/uci.c:551-605)
 *     are gathered for the host's UCI decoders, in the order srslte_uci_decode_ack_ri reads them (uci.c:843-857);
 *   - the ACK positions are zeroed (sch.c:1067-1070) -- here on the way through, q_bits itself is left untouched;
 *   - ulsch_deinterleave (sch.c:992-1019, table of ulsch_interleave_gen :658-679): the channel de-interleaver of
 *     TS 36.212 5.2.2.8 with the RI positions skipped, INCLUDING the reference's side effect that, when RI is present,
 *     g_bits[0] receives the LLR of the last RI position (every RI position is sent to index 0);
 *   - the CQI LLRs at the front of g_bits are handed back; UL-SCH data starts at g_bits + Q_prime_cqi * Qm and holds
 *     (H_prime_total - Q_prime_ri - Q_prime_cqi) * Qm e-bits -- the e_bits / nof_e_bits of srslte_b200_decode_tbs.
 * The Q' values come from the reference's own host code (Q_prime_ri_ack uci.c:606-630, Q_prime_cqi uci.c:329-345); the
 * UCI payload decoders stay there too and consume ack_llr / ri_llr / cqi_llr.
 *   q_bits          int16[H_prime_total * Qm], 4-byte aligned when a device pointer
 *   Qm              2, 4, 6 (8 accepted)       N_pusch_symbs  cfg->grant.nof_symb (12, 11 with SRS; 10, 9 extended CP)
 *   g_bits          out: int16[(H_prime_total - Q_prime_ri) * Qm] written, 4-byte aligned when a device pointer
 *   ack_llr, ri_llr, cqi_llr   HOST arrays of Q' * Qm int16 each, or NULL; when any is given the call returns after the
 *                   stream has drained (the values are needed before the UCI decoders can run)
 *                   -- unless SRSLTE_B200_UCI_DEFERRED is set (with SRSLTE_B200_OUT_DEVICE): then they are written by the next
 *                   srslte_b200_wait() on this context and the arrays must stay valid until then
 * flags: SRSLTE_B200_IN_DEVICE -> q_bits are device pointers; SRSLTE_B200_OUT_DEVICE -> g_bits are device pointers and stay
 * stream-ordered with a following srslte_b200_decode_tbs*(..., SRSLTE_B200_IN_DEVICE) on the same context.
 * Returns SRSLTE_B200_ERROR_INVALID_INPUTS for geometries the reference cannot index (H' not a multiple of N_pusch_symbs,
 * more than four ACK or RI symbols per row of the interleaver matrix, UCI columns beyond N_pusch_symbs). */
typedef struct {
  const int16_t* q_bits;
  uint32_t       Qm;
  uint32_t       H_prime_total;
  uint32_t       N_pusch_symbs;
  uint32_t       Q_prime_ack;
  uint32_t       Q_prime_ri;
  uint32_t       Q_prime_cqi;
  int16_t*       g_bits;
  int16_t*       ack_llr;
  int16_t*       ri_llr;
  int16_t*       cqi_llr;
} srslte_b200_ulsch_t;

SRSLTE_B200_API int srslte_b200_ulsch_deinterleave(srslte_b200_ctx_t* ctx, const srslte_b200_ulsch_t* tbs, uint32_t nof_tb, uint32_t flags);

/* ---- transmit mirror: the batched form of srslte_dlsch_encode2 / encode_tb_off (lib/src/phy/phch/sch.c:
 * 235-349): TB CRC24A, code block segmentation, CB CRC24B, srslte_tcod_encode_lut (lib/src/phy/fec/turbocoder.c:190-372),
 * srslte_rm_turbo_tx_lut (lib/src/phy/fec/rm_turbo.c:349-395).  Stateless: retransmissions (rv != 0) are re-encoded from
 * the payload instead of being read back from a srslte_softbuffer_tx_t; the bits are the same.
 *   data     tbs/8 payload bytes            e_bits  out: nof_e_bits packed bits (first bit = MSB of byte 0), 4-byte aligned,
 *                                                    (nof_e_bits + 31) / 32 * 4 bytes are written
 *   ret      0, or -1 (TBS needs filler bits / too many code blocks: the reference rejects those too), -2 (invalid)
 * flags as for srslte_b200_demod_descramble. */
typedef struct {
  const uint8_t* data;
  uint32_t       tbs;
  uint32_t       Qm;
  uint32_t       rv;
  uint32_t       nof_e_bits;
  uint8_t*       e_bits;
  int32_t        ret;
} srslte_b200_enc_t;

SRSLTE_B200_API int srslte_b200_encode_tbs(srslte_b200_ctx_t* ctx, srslte_b200_enc_t* tbs, uint32_t nof_tb, uint32_t flags);

/* device-resident HARQ soft buffers (srslte_softbuffer_rx_init / _reset / _free, softbuffer.c:41-155) */
SRSLTE_B200_API int  srslte_b200_softbuffer_create(srslte_b200_ctx_t* ctx, srslte_b200_softbuffer_t** sb, uint32_t max_cb);
SRSLTE_B200_API void srslte_b200_softbuffer_reset(srslte_b200_softbuffer_t* sb);
SRSLTE_B200_API void srslte_b200_softbuffer_free(srslte_b200_softbuffer_t* sb);

/* measurement hooks: device time of the kernels of the last completed batch, kernel launches it issued */
SRSLTE_B200_API float    srslte_b200_last_gpu_ms(srslte_b200_ctx_t* ctx);
SRSLTE_B200_API uint32_t srslte_b200_last_launches(srslte_b200_ctx_t* ctx);
/* device time (ms) and launch count of only the max-log-MAP kernel in the last batch */
SRSLTE_B200_API float    srslte_b200_last_map_ms(srslte_b200_ctx_t* ctx);
SRSLTE_B200_API uint32_t srslte_b200_last_map_launches(srslte_b200_ctx_t* ctx);

/* options: "fast16" (default 1): int16 windowed decoders first try the native packed-instruction path (VIADD.16x2 /
 * VIMNMX.S16x2 / VIADDMNMX.S16x2) under a range monitor and replay with the exact saturating arithmetic whenever
 * a saturation cannot be ruled out; 0 forces the exact arithmetic everywhere.  Results are identical either way. */
/*          "latency" (default 1): batches that leave at most one group of code blocks per SM (one subframe or a few) run
 * the latency-shaped MAP kernel (backward and forward recursions on two warps at once, LLRs spread over four); 0 keeps the
 * throughput kernel for every batch size.  Results are identical either way. */
/*          "gen_fused" (default 1): code blocks of the generic decoder (K <= 400 under SRSLTE_TDEC_AUTO, turbodecoder.c:381-393)
 * run all their half-iterations, decisions and CRC checks in one launch, one CTA per pair of blocks; 0 keeps one launch per
 * half-iteration.  "fused_spread" (default 1): the persistent MAP kernel uses smaller CTAs when a decoder class has fewer
 * groups of code blocks than the GPU has resident warps, so that every SM gets some.  Results are identical either way. */
SRSLTE_B200_API int srslte_b200_set_option(srslte_b200_ctx_t* ctx, const char* name, int value);
/* statistics of the last completed batch: (code block x half-iteration) units run, and how many were replayed */
SRSLTE_B200_API uint32_t srslte_b200_last_half_iterations(srslte_b200_ctx_t* ctx);
SRSLTE_B200_API uint32_t srslte_b200_last_replayed(srslte_b200_ctx_t* ctx);

/* test hook: one int16 LLR plane (in the decoder's lane layout, row = trellis step) of code block `cb` of the LAST completed
 * batch of this context -- plane 2: the a-priori input of constituent decoder 1 (what turbodecoder_iter.h:107-109 leaves in
 * app1 after its subtraction), plane 3: the input of decoder 2 (app2 of iter.h:117-121); 0 / 1 / 4: systematic / parity 0 /
 * parity 1 as extracted.  The parity tests compare them with the oracle's arrays after every half-iteration. */
SRSLTE_B200_API int srslte_b200_debug_read_plane(srslte_b200_ctx_t* ctx, uint32_t cb, uint32_t plane, int16_t* out, uint32_t n);

/* host-side tables as the engine uploads them: QPP permutation (tc_interl_lte.c:69-109) in the index space of a
 * `lanes`-lane layout (lanes <= 1: natural order), and the rate de-matching table of one rv
 * (rm_turbo.c:177-251 [+ :263-277 when lanes > 0]): table[i] = buffer position of the i-th received e-bit */
SRSLTE_B200_API int srslte_b200_qpp_table(uint32_t K, uint32_t lanes, uint16_t* fwd, uint16_t* rev);
SRSLTE_B200_API int srslte_b200_rm_table(uint32_t K, uint32_t rv, uint32_t lanes, uint16_t* table);

/* CUDA-event stopwatch on the engine's own stream (torch.cuda.Event would only see torch's stream) */
SRSLTE_B200_API int   srslte_b200_timer_start(srslte_b200_ctx_t* ctx);
SRSLTE_B200_API float srslte_b200_timer_stop_ms(srslte_b200_ctx_t* ctx);
/* integer-ALU roofline probe: packed int16x2 instructions per second (whole GPU, one instruction kind at a time, 8
 * independent chains per thread) -- op 0 VIADD.16x2, 1 VIMNMX.S16x2, 2 VIADDMNMX.S16x2, 3 VIMNMX3.S16x2,
 * 4 the saturating add __vaddss2 (multi-instruction emulation on sm_100a, counted per source operation) */
SRSLTE_B200_API double srslte_b200_alu_probe(srslte_b200_ctx_t* ctx, int op);

#ifdef __cplusplus
}
#endif
#endif
