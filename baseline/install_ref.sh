#!/usr/bin/env bash
# Offline install of the UNMODIFIED reference package into baseline/_ref (git-ignored; travels to the GPU box with
# the repo snapshot) for `bench.py --impl reference`.
#
# The reference's setup.py also builds a C extension for the GPC preprocessing (src/geoconv/preprocessing/
# c_extension/c_extension.c) that includes <cblas.h>; that header is not in this image, so the stock build stops
# there.  The extension is not on the measured path (conv_intrinsic.py / angular_max_pooling.py import none of
# it), so the install runs from a scratch copy whose setup.py declares no ext_modules; every Python source is
# installed byte for byte.  /root/reference itself is read-only and is never written.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
cat > "$TMP/ref/setup.py" <<'PY'
from setuptools import setup
setup()   # Python packages only: the cblas-dependent preprocessing extension is left out (see install_ref.sh)
PY
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP/ref"
rm -rf "$TMP"
# the installed hot-path sources are the reference's own
for f in pytorch/layers/conv_intrinsic.py pytorch/layers/conv_geodesic.py pytorch/layers/angular_max_pooling.py \
         preprocessing/barycentric_coordinates.py; do
  cmp "$SRC/src/geoconv/$f" "$ROOT/baseline/_ref/geoconv/$f"
done
echo "installed: $ROOT/baseline/_ref"
