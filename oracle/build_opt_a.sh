#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# "Option A" of INTEGRATION.md as a build: the reference's UNMODIFIED lib/src/phy/phch/sch.c (decode_tb / decode_tb_cb, the
# per-code-block loop) and its non-FEC dependencies, compiled against include/srslte_b200/fec.h through the shim include
# directory tests/shim_include (which replaces srslte/phy/fec/{crc,turbodecoder,softbuffer,cbsegm}.h) and linked against
# srsran_b200/libsrslte_fec_b200.so INSTEAD of the reference's turbodecoder*.c, crc.c, cbsegm.c, softbuffer.c and the receive
# half of rm_turbo.c.  tests/test_gpu_parity.py::test_unmodified_sch_c_on_the_b200_library drives it.
# Output: oracle/_ref/libsch_on_b200.so (git-ignored, travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${SRSLTE_REFERENCE:-/root/reference}"
R="$REF/lib"
OUT="$HERE/_ref"
if [ ! -d "$R/src/phy/fec" ]; then
  echo "build_opt_a.sh: reference tree not found at $REF (fine on the GPU box: prebuilt .so is used)" >&2
  exit 0
fi
if [ ! -f "$ROOT/srsran_b200/libsrslte_fec_b200.so" ]; then
  echo "build_opt_a.sh: build srsran_b200/libsrslte_fec_b200.so first" >&2
  exit 1
fi
mkdir -p "$OUT/inc/srslte" "$OUT/obja"
sed -e 's/@SRSLTE_VERSION_MAJOR@/20/;s/@SRSLTE_VERSION_MINOR@/10/;s/@SRSLTE_VERSION_PATCH@/1/;s/@SRSLTE_VERSION_STRING@/20.10.1/' \
  "$R/include/srslte/version.h.in" > "$OUT/inc/srslte/version.h"
CFLAGS="-O2 -std=gnu99 -fPIC -mavx2 -mfma -msse4.1 -DLV_HAVE_SSE -DLV_HAVE_AVX -DLV_HAVE_AVX2 -DLV_HAVE_FMA \
 -I$ROOT/tests/shim_include -I$ROOT/include -I$OUT/inc -I$R/include -fno-strict-aliasing -w \
 -Werror=implicit-function-declaration -Werror=incompatible-pointer-types -Werror=int-conversion"
# everything sch.c needs EXCEPT the files whose symbols the B200 library provides
CFILES="
src/phy/phch/sch.c src/phy/phch/uci.c src/phy/phch/cqi.c src/phy/phch/ra.c
src/phy/common/phy_common.c src/phy/common/sequence.c
src/phy/fec/turbocoder.c src/phy/fec/tc_interl_lte.c src/phy/fec/tc_interl_umts.c src/phy/fec/rm_turbo.c
src/phy/utils/vector.c src/phy/utils/vector_simd.c src/phy/utils/bit.c src/phy/utils/debug.c src/phy/utils/phy_logger.c
src/phy/fec/viterbi.c src/phy/fec/viterbi37_port.c src/phy/fec/viterbi37_sse.c src/phy/fec/viterbi37_avx2.c
src/phy/fec/viterbi37_avx2_16bit.c src/phy/fec/parity.c src/phy/fec/rm_conv.c src/phy/fec/convcoder.c
"
OBJS=""
for f in $CFILES; do
  o="$OUT/obja/$(echo "$f" | tr '/' '_' | sed 's/\.c$/.o/')"
  gcc $CFLAGS -c "$R/$f" -o "$o"
  OBJS="$OBJS $o"
done
# rm_turbo.c holds the transmit side (kept: the reference's CPU encoder) AND the receive side (replaced): make the receive
# symbols of the reference object local so that sch.c's calls bind to the B200 library
# (srslte_rm_turbo_gentables / _free_tables stay the reference's: they also build the transmit tables)
objcopy -L srslte_rm_turbo_rx_lut -L srslte_rm_turbo_rx_lut_ -L srslte_rm_turbo_rx_lut_8bit "$OUT/obja/src_phy_fec_rm_turbo.o"
gcc $CFLAGS -c "$HERE/opt_a_shim.c" -o "$OUT/obja/opt_a_shim.o"
g++ -shared -o "$OUT/libsch_on_b200.so" $OBJS "$OUT/obj/random.o" "$OUT/obja/opt_a_shim.o" \
  -L"$ROOT/srsran_b200" -lsrslte_fec_b200 -Wl,-rpath,'$ORIGIN/../../srsran_b200' -Wl,--no-undefined -lm -lpthread
echo "built $OUT/libsch_on_b200.so"
