#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference implementation of the turbo-decode hot path
# (srsLTE 20.10.1, /root/reference/lib) straight from where its sources lie,
# plus our thin C shim (oracle/ref_shim.c), into oracle/_ref/libsrslte_ref.so.
# No reference source is copied into this repo; only build outputs land in
# oracle/_ref/ (git-ignored, but NOT gpurun-ignored so the .so travels to the
# GPU box).  The reference's own cmake build is not used (it needs fftw3 etc.).
#
# Flags follow SURVEY.md §8c: AVX2 build WITHOUT -DLV_HAVE_AVX512 so that
# SRSLTE_SIMD_B_SIZE == 32 (the documented srslte_vec_sub_bbb tail behaviour).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SRSLTE_REFERENCE:-/root/reference}"
R="$REF/lib"
OUT="$HERE/_ref"
if [ ! -d "$R/src/phy/fec" ]; then
  echo "build_ref.sh: reference tree not found at $REF (fine on the GPU box: prebuilt .so is used)" >&2
  exit 0
fi
mkdir -p "$OUT/inc/srslte" "$OUT/obj"
sed -e 's/@SRSLTE_VERSION_MAJOR@/20/;s/@SRSLTE_VERSION_MINOR@/10/;s/@SRSLTE_VERSION_PATCH@/1/;s/@SRSLTE_VERSION_STRING@/20.10.1/' \
  "$R/include/srslte/version.h.in" > "$OUT/inc/srslte/version.h"

CFLAGS="-O3 -std=gnu99 -fPIC -mavx2 -mfma -msse4.1 -DLV_HAVE_SSE -DLV_HAVE_AVX -DLV_HAVE_AVX2 -DLV_HAVE_FMA \
 -I$OUT/inc -I$R/include -fno-strict-aliasing -w"

CFILES="
src/phy/fec/turbodecoder.c src/phy/fec/turbodecoder_gen.c src/phy/fec/turbodecoder_sse.c
src/phy/fec/tc_interl_lte.c src/phy/fec/tc_interl_umts.c src/phy/fec/cbsegm.c src/phy/fec/crc.c
src/phy/fec/rm_turbo.c src/phy/fec/turbocoder.c src/phy/fec/softbuffer.c
src/phy/utils/vector.c src/phy/utils/vector_simd.c src/phy/utils/bit.c src/phy/utils/debug.c src/phy/utils/phy_logger.c
src/phy/channel/ch_awgn.c src/phy/channel/gauss.c
src/phy/phch/sch.c src/phy/phch/uci.c src/phy/phch/cqi.c src/phy/phch/ra.c
src/phy/common/phy_common.c src/phy/common/sequence.c
src/phy/modem/demod_soft.c src/phy/scrambling/scrambling.c
src/phy/fec/viterbi.c src/phy/fec/viterbi37_port.c src/phy/fec/viterbi37_sse.c src/phy/fec/viterbi37_avx2.c
src/phy/fec/viterbi37_avx2_16bit.c src/phy/fec/parity.c src/phy/fec/rm_conv.c src/phy/fec/convcoder.c
"
OBJS=""
for f in $CFILES; do
  o="$OUT/obj/$(echo "$f" | tr '/' '_' | sed 's/\.c$/.o/')"
  if [ ! -f "$o" ] || [ "$R/$f" -nt "$o" ]; then
    gcc $CFLAGS -c "$R/$f" -o "$o"
  fi
  OBJS="$OBJS $o"
done
o="$OUT/obj/random.o"
if [ ! -f "$o" ]; then
  g++ -O2 -fPIC -std=c++11 -I"$OUT/inc" -I"$R/include" -w -c "$R/src/phy/utils/random.cpp" -o "$o"
fi
OBJS="$OBJS $o"
gcc $CFLAGS -c "$HERE/ref_shim.c" -o "$OUT/obj/ref_shim.o"
g++ -shared -o "$OUT/libsrslte_ref.so" $OBJS "$OUT/obj/ref_shim.o" -lm -lpthread
echo "built $OUT/libsrslte_ref.so"

# csi_correction is static in pdsch.c: a shim that includes the source file, everything else garbage-collected at link time
echo '{ global: ref_csi_correction; local: *; };' > "$OUT/obj/csi.map"
gcc $CFLAGS -ffunction-sections -fdata-sections -DREF_PDSCH_C="\"$R/src/phy/phch/pdsch.c\"" -c "$HERE/ref_csi_shim.c" -o "$OUT/obj/ref_csi_shim.o"
gcc -shared -o "$OUT/libsrslte_ref_csi.so" "$OUT/obj/ref_csi_shim.o" "$OUT/obj/src_phy_utils_vector.o" "$OUT/obj/src_phy_utils_vector_simd.o" \
  "$OUT/obj/src_phy_utils_bit.o" "$OUT/obj/src_phy_utils_debug.o" "$OUT/obj/src_phy_utils_phy_logger.o" \
  -Wl,--gc-sections -Wl,--version-script="$OUT/obj/csi.map" -Wl,--no-undefined -lm
echo "built $OUT/libsrslte_ref_csi.so"

# Second build for bench.py's CPU baseline only: the reference's RELEASE optimisation flags (CMakeLists.txt:394-417:
# -Ofast -funroll-loops -mfpmath=sse on top of the ISA flags; -march=native is left out because the library is built here and
# runs on the GPU box's host CPU).  The parity tests keep using the -O3 build above.
FFLAGS="-Ofast -funroll-loops -mfpmath=sse -std=gnu99 -fPIC -mavx2 -mfma -msse4.1 -DLV_HAVE_SSE -DLV_HAVE_AVX -DLV_HAVE_AVX2 -DLV_HAVE_FMA \
 -I$OUT/inc -I$R/include -fno-strict-aliasing -w"
mkdir -p "$OUT/objf"
FOBJS=""
for f in $CFILES; do
  o="$OUT/objf/$(echo "$f" | tr '/' '_' | sed 's/\.c$/.o/')"
  if [ ! -f "$o" ] || [ "$R/$f" -nt "$o" ]; then
    gcc $FFLAGS -c "$R/$f" -o "$o"
  fi
  FOBJS="$FOBJS $o"
done
gcc $FFLAGS -c "$HERE/ref_shim.c" -o "$OUT/objf/ref_shim.o"
g++ -shared -o "$OUT/libsrslte_ref_fast.so" $FOBJS "$OUT/obj/random.o" "$OUT/objf/ref_shim.o" -lm -lpthread
echo "built $OUT/libsrslte_ref_fast.so"
