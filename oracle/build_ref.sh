#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Builds the reference decoder (harutel/hls-jpeg-decoder) as plain C++ from the sources
# WHERE THEY LIE under $REF (default /root/reference), into oracle/_ref/ (git-ignored,
# travels to the GPU box with the snapshot).  Nothing of the reference is copied into the
# tracked tree.  The only adaptation is the three capacity macros of loadjpg.h:55-57
# (unguarded #defines, cannot be overridden with -D): a sed-patched header is written to
# a mktemp directory, used for this one compile and deleted.
#
# Variants (IMG_MAX_WIDTH x IMG_MAX_HEIGHT, JPG_FILE_SIZE in KB):
#   std : 512 x 512,   105   -- exactly as shipped (config 1, Lenna)
#   hd  : 1920 x 1088, 2000  -- configs 2/3/5
#   big : 8192 x 8192, 64000 -- config 4
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/loadjpg.cpp" ]; then
    echo "build_ref: $REF not present; keeping whatever is in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT/data"
CXX="${CXX:-g++}"
FLAGS="-O2 -ffp-contract=off -fno-fast-math -w -shared -fPIC"

build_variant() {
    local name="$1" w="$2" h="$3" kb="$4"
    local tmp; tmp="$(mktemp -d)"
    sed -E \
        -e "s/^#define[[:space:]]+JPG_FILE_SIZE[[:space:]]+[0-9]+/#define JPG_FILE_SIZE ${kb}/" \
        -e "s/^#define[[:space:]]+IMG_MAX_WIDTH[[:space:]]+[0-9]+/#define IMG_MAX_WIDTH ${w}/" \
        -e "s/^#define[[:space:]]+IMG_MAX_HEIGHT[[:space:]]+[0-9]+/#define IMG_MAX_HEIGHT ${h}/" \
        "$REF/src/loadjpg.h" > "$tmp/loadjpg.h"
    # -I$tmp first: ref_harness.cpp includes "loadjpg.h" before the reference .cpp files, so the
    # patched header wins and the reference's own copy is skipped by its include guard.
    $CXX $FLAGS -I"$tmp" -I"$REF/src" "$HERE/ref_harness.cpp" -o "$OUT/libhjdref_${name}.so"
    rm -rf "$tmp"
    echo "build_ref: built $OUT/libhjdref_${name}.so (${w}x${h}, ${kb} KB)"
}

build_variant std 512 512 105
build_variant hd 1920 1088 2000
build_variant big 8192 8192 64000

# The reference's only fixture (config 1).  Kept out of git history on purpose: it lives in the
# ignored oracle/_ref/ so that GPU-box tests can read it without touching /root/reference.
cp -f "$REF/data/Lenna.jpg" "$OUT/data/Lenna.jpg"
echo "build_ref: done"

# Drop-in demonstration at the HLS-top boundary: the reference's UNCHANGED src/main.cpp and
# src/openjpg.cpp, linked against this repository's JpegDecodeHW (csrc/ref_shim_hw.cpp + libhjd.so)
# instead of the reference's src/loadjpg.cpp.  tests/test_gpu_dropin.py runs it on the GPU box.
LIB_DIR="$HERE/../hls_jpeg_decoder_b200"
if [ -f "$LIB_DIR/libhjd.so" ]; then
    $CXX -O2 -w -I"$REF/src" "$REF/src/main.cpp" "$REF/src/openjpg.cpp" "$LIB_DIR/csrc/ref_shim_hw.cpp" \
        -L"$LIB_DIR" -lhjd -Wl,-rpath,'$ORIGIN/../../hls_jpeg_decoder_b200' -o "$OUT/ref_main_on_gpu"
    echo "build_ref: built $OUT/ref_main_on_gpu (reference main.cpp + openjpg.cpp on the GPU decode core)"
fi
