#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own alignment cores straight from the read-only reference
# checkout into oracle/_ref/ (git-ignored, binaries only - no reference source is copied).
#   libssw_ref.so   <- dysgu/scikitbio/ssw.c      (C99, -O2; mirrors meson.build:186-193)
#   libedlib_ref.so <- dysgu/edlib/src/edlib.cpp  (-O3 -std=c++17; mirrors meson.build:177-184)
# Used by tests/ (golden-vector generation, oracle pinning) and by bench.py's
# `--impl reference` / cpu_baseline legs.  Never loaded by the product path.
set -euo pipefail
REF="${DYSGU_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/dysgu/scikitbio" ]; then
    echo "build_ref: reference tree not found at $REF (ok on the GPU box: prebuilt files are used)" >&2
    exit 0
fi
mkdir -p "$OUT"
gcc -O2 -std=c99 -fPIC -shared -w -I"$REF/dysgu/scikitbio" "$REF/dysgu/scikitbio/ssw.c" -o "$OUT/libssw_ref.so" -lm
g++ -O3 -std=c++17 -fPIC -shared -w -I"$REF/dysgu/edlib/src" "$REF/dysgu/edlib/src/edlib.cpp" -o "$OUT/libedlib_ref.so"
echo "build_ref: built $OUT/libssw_ref.so $OUT/libedlib_ref.so"
