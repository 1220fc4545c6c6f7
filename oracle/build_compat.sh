#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY - the drop-in proof of SURVEY 8b.
# Builds the reference's own two Cython wrappers, UNMODIFIED and from where they lie in the reference checkout, but links
# them against dysgu_b200/libdysgu_b200_compat.so (the reference's symbol names over the GPU engine) instead of compiling
# ssw.c / edlib.cpp into them (meson.build:177-193 is the recipe being mirrored).  Only the two extension modules land in
# oracle/_ref/pycompat/dysgu/...; no reference source is copied into the repository (the .pyx files pass through a
# scratch directory for cython, as in build_ref.sh).  tests/test_compat_gpu.py runs the wrapper-level golden vectors through
# these modules on the GPU.
set -euo pipefail
REF="${DYSGU_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref/pycompat"
LIBDIR="$(cd "$HERE/../dysgu_b200" && pwd)"
if [ ! -d "$REF/dysgu/scikitbio" ]; then
    echo "build_compat: reference tree not found at $REF (ok on the GPU box: prebuilt files are used)" >&2
    exit 0
fi
if [ ! -f "$LIBDIR/libdysgu_b200_compat.so" ]; then
    echo "build_compat: $LIBDIR/libdysgu_b200_compat.so missing - run python -m dysgu_b200.build first" >&2
    exit 1
fi
python -c "import Cython" 2>/dev/null || { echo "build_compat: Cython not available" >&2; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
NPINC="$(python -c 'import numpy; print(numpy.get_include())')"
EXT="$(python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
mkdir -p "$TMP/dysgu/scikitbio" "$TMP/dysgu/edlib" "$OUT/dysgu/scikitbio" "$OUT/dysgu/edlib"
touch "$TMP/dysgu/__init__.py" "$TMP/dysgu/scikitbio/__init__.py" "$TMP/dysgu/edlib/__init__.py"
cp "$REF/dysgu/scikitbio/_ssw_wrapper.pyx" "$TMP/dysgu/scikitbio/"
cp "$REF/dysgu/edlib/edlib.pyx" "$REF/dysgu/edlib/edlib.pxd" "$TMP/dysgu/edlib/"
# the extension modules find the engine relative to their own location: oracle/_ref/pycompat/dysgu/<pkg>/ -> dysgu_b200/
RPATH='$ORIGIN/../../../../../dysgu_b200'
cd "$TMP"
cython -3 dysgu/scikitbio/_ssw_wrapper.pyx -o ssw_wrapper.c >/dev/null 2>&1
gcc -O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION -I"$PYINC" -I"$NPINC" -I"$REF/dysgu/scikitbio" \
    ssw_wrapper.c -o "$OUT/dysgu/scikitbio/_ssw_wrapper$EXT" -L"$LIBDIR" -ldysgu_b200_compat -Wl,-rpath,"$RPATH" -Wl,--no-undefined \
    $(python3-config --ldflags --embed 2>/dev/null || true) -lm
cython --cplus -3 -I . dysgu/edlib/edlib.pyx -o edlib_wrap.cpp >/dev/null 2>&1
g++ -O2 -std=c++17 -fPIC -shared -w -I"$PYINC" -I"$REF/dysgu/edlib" -I"$REF/dysgu/edlib/src" -I"$TMP/dysgu/edlib" -I. \
    edlib_wrap.cpp -o "$OUT/dysgu/edlib/edlib$EXT" -L"$LIBDIR" -ldysgu_b200_compat -Wl,-rpath,"$RPATH" \
    $(python3-config --ldflags --embed 2>/dev/null || true)
touch "$OUT/dysgu/__init__.py" "$OUT/dysgu/scikitbio/__init__.py" "$OUT/dysgu/edlib/__init__.py"
echo "build_compat: built the reference's Cython wrappers against libdysgu_b200_compat.so under $OUT"
