#!/usr/bin/env bash
# Build libbayeslm_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/bayeslms_b200/csrc"
OUT="$ROOT/bayeslms_b200/lib"
mkdir -p "$OUT" "$ROOT/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -I"$ROOT/include" ${BLM_NVCC_EXTRA:-})
pids=()
for f in blm_runtime blm_gemm blm_gemm2 blm_gemm_ln blm_gemm_sampled blm_elementwise blm_attention blm_lstm blm_train blm_cells; do
  "$NVCC" "${FLAGS[@]}" -c "$SRC/$f.cu" -o "$ROOT/build/$f.o" &
  pids+=($!)
done
"$NVCC" "${FLAGS[@]}" -x cu -c "$SRC/blm_text.cpp" -o "$ROOT/build/blm_text.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libbayeslm_b200.so" \
  "$ROOT"/build/blm_runtime.o "$ROOT"/build/blm_gemm.o "$ROOT"/build/blm_gemm2.o "$ROOT"/build/blm_gemm_ln.o "$ROOT"/build/blm_gemm_sampled.o "$ROOT"/build/blm_elementwise.o \
  "$ROOT"/build/blm_attention.o "$ROOT"/build/blm_lstm.o "$ROOT"/build/blm_train.o "$ROOT"/build/blm_text.o "$ROOT"/build/blm_cells.o -cudart static
echo "built $OUT/libbayeslm_b200.so"
