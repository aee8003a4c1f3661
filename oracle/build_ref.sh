#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference CUDA extension (ytgui/SPT-proto
# extension/*.cu, *.cpp) for sm_100a from the sources where they lie under /root/reference
# into oracle/_ref/ext_ref.so (git-ignored, travels to the GPU box with gpurun).
# It is the GPU-side bit-level oracle for PQ codes / lookup indices / fp32 values at the
# shapes the reference supports (fp32, S in {256,512,1024}, m in {8,10,16}).
# Only tests/ may load it.  Nothing is copied out of /root/reference.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SPT_REFERENCE_DIR:-/root/reference}/extension"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "reference sources not present at $REF; skipping"; exit 0; }
mkdir -p "$OUT/obj"
PY=python
TORCH_DIR="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))' 2>/dev/null)"
PY_INC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
CUSPARSE_LIB="$($PY -c 'import nvidia.cusparse,os;print(os.path.join(list(nvidia.cusparse.__path__)[0],"lib"))' 2>/dev/null || echo /usr/local/cuda/lib64)"
CUDA=/usr/local/cuda
COMMON=(-std=c++17 -O2 -DTORCH_EXTENSION_NAME=ext_ref -DTORCH_API_INCLUDE_EXTENSION_H
        -D_GLIBCXX_USE_CXX11_ABI=1
        -I"$HERE/ref_shim" -I"$REF" -I"$TORCH_DIR/include"
        -I"$TORCH_DIR/include/torch/csrc/api/include" -I"$PY_INC" -I"$CUDA/include")
pids=()
for f in cdist lookup softmax; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cu" -nt "$OUT/obj/$f.o" ]; then
    $CUDA/bin/nvcc -x cu -c "$REF/$f.cu" -o "$OUT/obj/$f.o" "${COMMON[@]}" \
      --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC &
    pids+=($!)
  fi
done
for f in entry sddmm spmm; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cpp" -nt "$OUT/obj/$f.o" ]; then
    g++ -c "$REF/$f.cpp" -o "$OUT/obj/$f.o" "${COMMON[@]}" -fPIC -w &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared "$OUT"/obj/{cdist,lookup,softmax,entry,sddmm,spmm}.o -o "$OUT/ext_ref.so" \
  -L"$TORCH_DIR/lib" -lc10 -lc10_cuda -ltorch -ltorch_cpu -ltorch_cuda -ltorch_python \
  -L"$CUDA/lib64" -lcudart -L"$CUSPARSE_LIB" -l:libcusparse.so.12 \
  -Wl,-rpath,"$TORCH_DIR/lib" -Wl,-rpath,"$CUSPARSE_LIB"
echo "built $OUT/ext_ref.so"
