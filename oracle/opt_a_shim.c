/*
 * TEST INFRASTRUCTURE ONLY -- flat-C driver for oracle/build_opt_a.sh: the reference's unmodified sch.c running on the B200
 * library's drop-in symbols (INTEGRATION.md option A).  Forwards to srslte_sch_init / srslte_dlsch_encode2 /
 * srslte_dlsch_decode2 (lib/src/phy/phch/sch.c:116-233, 577-640); nothing here re-implements reference arithmetic.
 */
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "srslte/phy/phch/pdsch_cfg.h"
#include "srslte/phy/phch/sch.h"

typedef struct {
  srslte_sch_t           sch;
  srslte_softbuffer_rx_t rx;
  srslte_softbuffer_tx_t tx;
} opta_t;

void* opta_new(int llr_is_8bit, uint32_t max_iter)
{
  opta_t* s = calloc(1, sizeof(opta_t));
  if (srslte_sch_init(&s->sch) || srslte_softbuffer_rx_init(&s->rx, 100) || srslte_softbuffer_tx_init(&s->tx, 100)) {
    free(s);
    return NULL;
  }
  s->sch.llr_is_8bit = llr_is_8bit ? true : false;
  srslte_sch_set_max_noi(&s->sch, max_iter);
  return s;
}
void opta_del(void* ss)
{
  opta_t* s = ss;
  srslte_sch_free(&s->sch);
  srslte_softbuffer_rx_free(&s->rx);
  srslte_softbuffer_tx_free(&s->tx);
  free(s);
}
static void fill(opta_t* s, srslte_pdsch_cfg_t* cfg, uint32_t tbs, uint32_t Qm, uint32_t G, uint32_t rv, int tx)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->grant.nof_tb        = 1;
  cfg->grant.nof_layers    = 1;
  cfg->grant.nof_re        = G / Qm;
  cfg->grant.tb[0].enabled = true;
  cfg->grant.tb[0].tbs     = (int)tbs;
  cfg->grant.tb[0].rv      = (int)rv;
  cfg->grant.tb[0].mod     = Qm == 2 ? SRSLTE_MOD_QPSK : Qm == 4 ? SRSLTE_MOD_16QAM : Qm == 6 ? SRSLTE_MOD_64QAM : SRSLTE_MOD_256QAM;
  cfg->grant.tb[0].nof_bits = G;
  if (tx) /* (softbuffers is a union, pdsch_cfg.h:65-68) */
    cfg->softbuffers.tx[0] = &s->tx;
  else
    cfg->softbuffers.rx[0] = &s->rx;
}
void opta_reset_rx(void* ss, uint32_t tbs) { srslte_softbuffer_rx_reset_tbs(&((opta_t*)ss)->rx, tbs); }
int  opta_encode(void* ss, uint32_t tbs, uint32_t Qm, uint32_t G, uint32_t rv, uint8_t* data, uint8_t* e_bytes)
{
  opta_t*            s = ss;
  srslte_pdsch_cfg_t cfg;
  fill(s, &cfg, tbs, Qm, G, rv, 1);
  if (rv == 0)
    srslte_softbuffer_tx_reset_tbs(&s->tx, tbs);
  return srslte_dlsch_encode2(&s->sch, &cfg, data, e_bytes, 0, 1);
}
int opta_decode(void* ss, uint32_t tbs, uint32_t Qm, uint32_t G, uint32_t rv, void* llr, uint8_t* data_out, float* avg_iter, uint8_t* cb_crc_out,
                uint32_t max_cb_out)
{
  opta_t*            s = ss;
  srslte_pdsch_cfg_t cfg;
  fill(s, &cfg, tbs, Qm, G, rv, 0);
  int r = srslte_dlsch_decode2(&s->sch, &cfg, (int16_t*)llr, data_out, 0, 1);
  if (avg_iter)
    *avg_iter = srslte_sch_last_noi(&s->sch);
  for (uint32_t i = 0; cb_crc_out && i < max_cb_out && i < s->rx.max_cb; i++)
    cb_crc_out[i] = s->rx.cb_crc[i] ? 1 : 0;
  return r;
}
