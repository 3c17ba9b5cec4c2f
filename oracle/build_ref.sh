#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (clberube/BISIP 1.2.1) forward /
# log-probability path into oracle/_ref/ (git-ignored, but shipped to the GPU box).
#
# * reads sources where they lie under /root/reference (read-only), writes ONLY oracle/_ref/
# * the single change is pyx line 19 `DTYPE = np.float_` -> `np.float64` (np.float_ was
#   removed in NumPy 2.0; semantically identical) — SURVEY.md App. D
# * the shipped cython_funcs.c (Cython 0.29.15) does not compile on Python 3.12, so the
#   .pyx is re-cythonized with the Cython in this image
# * emcee/matplotlib/corner are absent from the image: oracle/refload.py stubs them at import
set -euo pipefail
REF=${BISIP_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/bisip" ]; then
  echo "build_ref: $REF not present (GPU box?) — using prebuilt $OUT if any" >&2
  exit 0
fi
PY=${PYTHON:-python}
[ -d "$OUT" ] && chmod -R u+w "$OUT"
mkdir -p "$OUT/bisip"
cp "$REF"/src/bisip/{__init__,models,utils,plotlib,data}.py "$OUT/bisip/"
rm -rf "$OUT/bisip/data" "$OUT/bisip/tests"
cp -r "$REF/src/bisip/data" "$REF/src/bisip/tests" "$OUT/bisip/"
chmod -R u+w "$OUT"
sed 's/^DTYPE = np.float_$/DTYPE = np.float64/' "$REF/src/bisip/cython_funcs.pyx" > "$OUT/bisip/cython_funcs.pyx"
$PY -m cython -3 "$OUT/bisip/cython_funcs.pyx" -o "$OUT/cython_funcs.c"
INC_PY=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
INC_NP=$($PY -c "import numpy;print(numpy.get_include())")
SUF=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
gcc -O2 -fPIC -shared -w -I"$INC_PY" -I"$INC_NP" "$OUT/cython_funcs.c" -o "$OUT/bisip/cython_funcs$SUF"
echo "build_ref: built $OUT/bisip/cython_funcs$SUF"
