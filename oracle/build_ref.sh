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
g++ -O2 -std=c++17 -fPIC -shared -w -I"$REF/dysgu" "$HERE/xxh_ref_wrap.cpp" -o "$OUT/libxxh_ref.so"   # the reference's XXHash64 (graph.pyx minimizers)
echo "build_ref: built $OUT/libssw_ref.so $OUT/libedlib_ref.so $OUT/libxxh_ref.so"

# Optional: the reference's two Cython wrappers (used only to generate wrapper-level golden vectors).
# Built in a scratch directory; only the extension modules land in oracle/_ref/pyref/dysgu/...
if python -c "import Cython" 2>/dev/null; then
    TMP="$(mktemp -d)"
    PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
    NPINC="$(python -c 'import numpy; print(numpy.get_include())')"
    EXT="$(python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
    mkdir -p "$TMP/dysgu/scikitbio" "$TMP/dysgu/edlib" "$OUT/pyref/dysgu/scikitbio" "$OUT/pyref/dysgu/edlib"
    touch "$TMP/dysgu/__init__.py" "$TMP/dysgu/scikitbio/__init__.py" "$TMP/dysgu/edlib/__init__.py"
    cp "$REF/dysgu/scikitbio/_ssw_wrapper.pyx" "$TMP/dysgu/scikitbio/"
    cp "$REF/dysgu/edlib/edlib.pyx" "$REF/dysgu/edlib/edlib.pxd" "$TMP/dysgu/edlib/"
    ( cd "$TMP" &&
      cython -3 dysgu/scikitbio/_ssw_wrapper.pyx -o ssw_wrapper.c >/dev/null 2>&1 &&
      gcc -O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION -I"$PYINC" -I"$NPINC" -I"$REF/dysgu/scikitbio" \
          ssw_wrapper.c "$REF/dysgu/scikitbio/ssw.c" -o "$OUT/pyref/dysgu/scikitbio/_ssw_wrapper$EXT" -lm &&
      cython --cplus -3 -I . dysgu/edlib/edlib.pyx -o edlib_wrap.cpp >/dev/null 2>&1 &&
      g++ -O3 -std=c++17 -fPIC -shared -w -I"$PYINC" -I"$REF/dysgu/edlib" -I"$REF/dysgu/edlib/src" -I"$TMP/dysgu/edlib" -I. \
          edlib_wrap.cpp "$REF/dysgu/edlib/src/edlib.cpp" -o "$OUT/pyref/dysgu/edlib/edlib$EXT" ) \
      && touch "$OUT/pyref/dysgu/__init__.py" "$OUT/pyref/dysgu/scikitbio/__init__.py" "$OUT/pyref/dysgu/edlib/__init__.py" \
      && echo "build_ref: built the reference Cython wrappers under $OUT/pyref" \
      || echo "build_ref: Cython wrappers not built (optional)" >&2
    rm -rf "$TMP"
fi
