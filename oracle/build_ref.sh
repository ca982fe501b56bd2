#!/usr/bin/env bash
# Builds the UNMODIFIED reference extension (megastep/src/{wrappers.cpp,kernels.cu,common.h}) for sm_100a,
# from the sources where they lie under /root/reference, into oracle/_ref/ (git-ignored; travels to the GPU
# box with gpurun). Nothing is copied into the repo. TEST/BENCH INFRASTRUCTURE ONLY: the result is the
# bit-exactness oracle and the timed reference arm (`bench.py --impl reference`); the product never loads it.
#
# Flags follow the reference's own JIT build (megastep/__init__.py:14-15: --use_fast_math -lineinfo) with the
# one change torch>=2 forces: nvcc -std=c++14 -> -std=c++17 (ATen.h:5 "#error C++17 or later").
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
SRC="$REF/megastep/src"
if [ ! -f "$SRC/kernels.cu" ]; then
    echo "build_ref: $SRC not present; keeping any prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
PY=${PYTHON:-python}
# The reference's PYTHON package (megastep/, rebar/), pip-installed — not copied into the repo — into baseline/_ref (the
# git-ignored directory the bench contract names for the installed reference; it travels to the GPU box): the parity tests
# and `bench.py --impl reference` drive the reference's own unmodified modules.py / core.py / scene.py / demo envs through
# it (tests/common.py::reference_package). pip builds in the source tree, so install from a copy under /tmp.
SITE="$HERE/../baseline/_ref"
if [ ! -f "$SITE/megastep/modules.py" ] || [ "$REF/megastep/modules.py" -nt "$SITE/megastep/modules.py" ]; then
    TMP=$(mktemp -d)
    cp -r "$REF" "$TMP/reference"
    rm -rf "$SITE" "$OUT/site"
    mkdir -p "$SITE"
    $PY -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target "$SITE" "$TMP/reference" >/dev/null 2>&1 && echo "build_ref: installed the reference package into baseline/_ref" \
        || echo "build_ref: pip install of the reference package failed (tests that drive its Python will skip)" >&2
    rm -rf "$TMP"
fi
TORCH_INC=$($PY - <<'EOF'
import warnings, logging
logging.disable(logging.CRITICAL)
import torch.utils.cpp_extension as c
print(' '.join('-I' + p for p in c.include_paths()))
EOF
)
TORCH_LIB=$($PY -c "import logging; logging.disable(logging.CRITICAL); import torch.utils.cpp_extension as c; print(c.library_paths()[0])")
PY_INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
TARGET="$OUT/megastepcuda$EXT"
if [ "$TARGET" -nt "$SRC/kernels.cu" ] && [ "$TARGET" -nt "$SRC/wrappers.cpp" ] && [ "$TARGET" -nt "$SRC/common.h" ]; then
    echo "build_ref: $TARGET up to date"
    exit 0
fi
DEFS="-DTORCH_EXTENSION_NAME=megastepcuda -DTORCH_API_INCLUDE_EXTENSION_H"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -c "$SRC/kernels.cu" -o "$OUT/kernels.o" -std=c++17 --use_fast_math -lineinfo \
    -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC \
    $TORCH_INC -I"$PY_INC" $DEFS &
g++ -c "$SRC/wrappers.cpp" -o "$OUT/wrappers.o" -std=c++17 -fPIC -O2 \
    $TORCH_INC -I"$PY_INC" -I/usr/local/cuda/include $DEFS &
wait
g++ -shared "$OUT/wrappers.o" "$OUT/kernels.o" -o "$TARGET" \
    -L"$TORCH_LIB" -Wl,-rpath,"$TORCH_LIB" -L/usr/local/cuda/lib64 \
    -ltorch -ltorch_python -lc10_cuda -lc10 -ltorch_cpu -ltorch_cuda -lcudart
echo "build_ref: built $TARGET"
