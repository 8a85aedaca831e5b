#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's own preprocessing C extension
# (/root/reference/src/geoconv/preprocessing/c_extension/c_extension.c) from the source where it lies into
# oracle/_ref/c_extension*.so.  Nothing is copied into the repository; oracle/_ref/ is git-ignored.
# The only thing the reference build needs that the image lacks is <cblas.h>: oracle/ref_build/cblas.h maps
# the five level-1 calls onto the OpenBLAS that scipy's wheel bundles (see that header).
# Usage: oracle/build_ref.sh [reference_root]   (exit 0 and a message when the reference is absent)
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$REF/src/geoconv/preprocessing/c_extension/c_extension.c"
if [ ! -f "$SRC" ]; then echo "oracle/build_ref.sh: $SRC not present - nothing to build"; exit 0; fi
PY="${PYTHON:-python}"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
NPINC="$($PY -c 'import numpy; print(numpy.get_include())')"
EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
BLASDIR="$($PY -c 'import scipy, os; print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))')"
BLAS="$(ls "$BLASDIR"/libscipy_openblas-*.so | head -1)"
mkdir -p "$HERE/_ref"
gcc -O2 -fPIC -shared -I"$HERE/ref_build" -I"$PYINC" -I"$NPINC" "$SRC" -o "$HERE/_ref/c_extension$EXT" \
    "$BLAS" -Wl,-rpath,"$BLASDIR" -lm
echo "built $HERE/_ref/c_extension$EXT against $BLAS"
