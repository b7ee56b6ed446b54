/* Shim for the drop-in compile test (tests/test_c_abi.py::test_reference_call_sites_compile_against_fec_h): placed BEFORE the
 * reference's include directory, it replaces srslte/phy/fec/turbodecoder.h by the declarations of the B200 library, so that the
 * reference's unmodified sch.c / pssch.c are type-checked against include/srslte_b200/fec.h. */
#include "srslte_b200/fec.h"
