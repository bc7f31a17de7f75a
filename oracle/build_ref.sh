#!/usr/bin/env bash
# Compiles the REFERENCE's own relabel op (csrc/relabel.cpp + csrc/cpu/relabel_cpu.cpp)
# from where it lies under /root/reference into oracle/_ref/ref_relabel.so.
# No reference source is copied into this repo; only the built .so lands in
# oracle/_ref/ (git-ignored, but it travels to the GPU box with gpurun).
# Test infrastructure only: nothing on the product path loads this file.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/csrc" ] || { echo "reference not present at $REF; keeping prebuilt files"; exit 0; }
mkdir -p "$OUT"
if [ -f "$OUT/ref_relabel.so" ] && [ "$OUT/ref_relabel.so" -nt "$REF/csrc/cpu/relabel_cpu.cpp" ]; then
  echo "oracle/_ref/ref_relabel.so up to date"; exit 0
fi
PY=${PYTHON:-python}
TORCH_INC=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
PY_INC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
ABI=$($PY -c "import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
g++ -O2 -std=c++17 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI=$ABI \
  -I"$REF/csrc" -I"$TORCH_INC" -I"$TORCH_INC/torch/csrc/api/include" -I"$PY_INC" \
  "$REF/csrc/relabel.cpp" "$REF/csrc/cpu/relabel_cpu.cpp" \
  -L"$TORCH_LIB" -ltorch -ltorch_cpu -lc10 -Wl,-rpath,"$TORCH_LIB" \
  -o "$OUT/ref_relabel.so"
echo "built $OUT/ref_relabel.so"
