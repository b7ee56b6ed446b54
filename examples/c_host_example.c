/*
 * C99 host example: the receive path of lib/src/phy/phch/sch.c written against the drop-in headers.
 *
 *   gcc -std=c99 -Iinclude examples/c_host_example.c -Lsrsran_b200 -lsrslte_fec_b200 -Wl,-rpath,$PWD/srsran_b200 -o c_host_example
 *
 * 1. encodes one transport block with the transmit mirror (srslte_dlsch_encode2 semantics),
 * 2. decodes it through the batched entry (one launch per TB: what decode_tb becomes),
 * 3. decodes its first code block the way decode_tb_cb (sch.c:363-488) does, call by call, with the drop-in symbols:
 *    srslte_rm_turbo_rx_lut_ -> srslte_tdec_new_cb -> srslte_tdec_iteration -> srslte_crc_checksum_byte.
 * Exit code 0 = every check passed.  Needs a B200 (there is no CPU fallback: the calls fail loudly without one).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "srslte_b200/batch.h"
#include "srslte_b200/fec.h"

#define CHECK(c)                                                                                                        \
  do {                                                                                                                 \
    if (!(c)) {                                                                                                        \
      fprintf(stderr, "check failed at line %d: %s (%s)\n", __LINE__, #c, srslte_b200_last_error());                   \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

int main(void)
{
  const uint32_t tbs = 15264, Qm = 4, G = 20000, max_it = 8;
  srslte_b200_ctx_t* ctx = NULL;
  CHECK(srslte_b200_ctx_create(&ctx, 0) == 0);

  /* ---- 1. transmit mirror */
  uint8_t* data = (uint8_t*)malloc(tbs / 8);
  srand(7);
  for (uint32_t i = 0; i < tbs / 8; i++)
    data[i] = (uint8_t)rand();
  uint8_t*          e_packed = (uint8_t*)calloc((G + 31) / 32 * 4, 1);
  srslte_b200_enc_t tx       = {data, tbs, Qm, 0, G, e_packed, 0};
  CHECK(srslte_b200_encode_tbs(ctx, &tx, 1, 0) == 0 && tx.ret == 0);

  /* noiseless channel: bit 1 -> +40, bit 0 -> -40 (the LLR convention of turbodecoder_test.c:246-253) */
  int16_t* llr = (int16_t*)malloc(G * sizeof(int16_t));
  for (uint32_t i = 0; i < G; i++)
    llr[i] = ((e_packed[i / 8] >> (7 - i % 8)) & 1) ? 40 : -40;

  /* ---- 2. one call per transport block */
  uint8_t*         out = (uint8_t*)calloc(tbs / 8 + 16, 1);
  srslte_b200_tb_t tb;
  memset(&tb, 0, sizeof(tb));
  tb.e_bits = llr, tb.nof_e_bits = G, tb.tbs = tbs, tb.Qm = Qm, tb.rv = 0, tb.softbuffer = NULL, tb.data = out;
  CHECK(srslte_b200_decode_tbs(ctx, &tb, 1, 0, max_it, 0) == 0);
  CHECK(tb.ret == 0 && memcmp(out, data, tbs / 8) == 0);
  printf("batched decode_tb: ok, %u code blocks, %.2f half-iterations on average\n", tb.nof_cb, tb.avg_iterations);

  /* ---- 3. decode_tb_cb, call by call, with the reference's own symbols */
  srslte_cbsegm_t seg;
  CHECK(srslte_cbsegm(&seg, tbs) == 0 && seg.F == 0);
  const uint32_t K = seg.K1, cb_idx = (uint32_t)srslte_cbsegm_cbindex(K);
  const uint32_t Gp = G / Qm, gamma = Gp % seg.C, n_e = Qm * (Gp / seg.C); /* code block 0: sch.c:391-397 */
  (void)gamma;
  int16_t* cb_llr = (int16_t*)calloc(3 * K + 12 + 64, sizeof(int16_t));
  CHECK(srslte_rm_turbo_rx_lut_(llr, cb_llr, n_e, cb_idx, 0, false) == 0); /* standard order (no sub-block layout) */
  srslte_tdec_t dec;
  srslte_crc_t  crc_cb;
  CHECK(srslte_tdec_init(&dec, 6144) == 0);
  srslte_tdec_force_not_sb(&dec);
  CHECK(srslte_crc_init(&crc_cb, 0x1800063, 24) == 0); /* SRSLTE_LTE_CRC24B */
  CHECK(srslte_tdec_new_cb(&dec, K) == 0);
  uint8_t* cb_bytes = (uint8_t*)calloc(K / 8 + 8, 1);
  uint32_t it = 0, crc = 1;
  do {
    srslte_tdec_iteration(&dec, cb_llr, cb_bytes);
    crc = srslte_crc_checksum_byte(&crc_cb, cb_bytes, (int)K);
    it++;
  } while (crc != 0 && it < max_it);
  CHECK(crc == 0 && memcmp(cb_bytes, data, (K - 24) / 8) == 0);
  printf("decode_tb_cb loop with srslte_* symbols: ok after %u half-iteration(s), K = %u\n", it, K);

  srslte_tdec_free(&dec);
  srslte_b200_ctx_destroy(ctx);
  free(data), free(e_packed), free(llr), free(out), free(cb_llr), free(cb_bytes);
  return 0;
}
