/*
 * TEST INFRASTRUCTURE ONLY.
 * csi_correction() is a STATIC function of the reference's lib/src/phy/phch/pdsch.c (:628-741).  To run the unmodified
 * function, this translation unit includes that source file where it lies (REF_PDSCH_C is passed by oracle/build_ref.sh) and
 * exports one wrapper; the link uses --gc-sections and a version script, so every other function of pdsch.c -- and its
 * dependencies on the rest of the library -- is dropped.  Nothing here re-implements reference arithmetic.
 */
#include REF_PDSCH_C

__attribute__((visibility("default"))) void ref_csi_correction(float* csi, void* e, uint32_t nof_bits, int mod, int is8)
{
  static srslte_pdsch_t q;
  srslte_pdsch_cfg_t    cfg;
  memset(&q, 0, sizeof(q));
  memset(&cfg, 0, sizeof(cfg));
  q.csi[0]                 = csi;
  q.llr_is_8bit            = is8 ? true : false;
  cfg.grant.tb[0].mod      = (srslte_mod_t)mod;
  cfg.grant.tb[0].nof_bits = nof_bits;
  csi_correction(&q, &cfg, 0, 0, e);
}
