#!/usr/bin/env bash
# Compiles the REFERENCE's own async transfer op (csrc/async.cpp + csrc/cuda/async_cuda.cu, -DWITH_CUDA)
# for sm_100a from where it lies under /root/reference into oracle/_ref/ref_async.so.  No reference
# source is copied; only the built .so lands in oracle/_ref/ (git-ignored, travels with gpurun).
# Test infrastructure only: it pins read_async / write_async semantics on the GPU box
# (tests/test_gpu_ref_async.py runs it in a subprocess: its operator names collide with the product's).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/csrc" ] || { echo "reference not present at $REF; keeping prebuilt files"; exit 0; }
mkdir -p "$OUT"
if [ -f "$OUT/ref_async.so" ] && [ "$OUT/ref_async.so" -nt "$REF/csrc/cuda/async_cuda.cu" ]; then
  echo "oracle/_ref/ref_async.so up to date"; exit 0
fi
PY=${PYTHON:-python}
TORCH_INC=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
PY_INC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
ABI=$($PY -c "import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
INC="-I$REF/csrc -I$TORCH_INC -I$TORCH_INC/torch/csrc/api/include -I$PY_INC -I/usr/local/cuda/include"
TMP=$(mktemp -d)
g++ -O2 -std=c++17 -fPIC -DWITH_CUDA -D_GLIBCXX_USE_CXX11_ABI=$ABI $INC -c "$REF/csrc/async.cpp" -o "$TMP/async.o"
/usr/local/cuda/bin/nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DWITH_CUDA \
  -D_GLIBCXX_USE_CXX11_ABI=$ABI $INC -c "$REF/csrc/cuda/async_cuda.cu" -o "$TMP/async_cuda.o"
g++ -shared -o "$OUT/ref_async.so" "$TMP/async.o" "$TMP/async_cuda.o" -L"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_cuda \
  -lc10 -lc10_cuda -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TORCH_LIB"
rm -rf "$TMP"
echo "built $OUT/ref_async.so"
