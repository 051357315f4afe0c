#!/usr/bin/env bash
# Builds svol_b200/csrc/libsvol_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v ${SVOL_EXTRA_NVCC_FLAGS:-})
SRCS=(api gemm_tc ffn_tc attn_tc attn_small attn_bwd_tc rowwise train evaluate matcher criterion plain)
mkdir -p build
pids=()
for s in "${SRCS[@]}"; do
  if [[ ! -f build/$s.o || $s.cu -nt build/$s.o || common.cuh -nt build/$s.o || svol_internal.h -nt build/$s.o \
        || ../../include/svol_b200.h -nt build/$s.o ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$s.cu" -o "build/$s.o" > "build/$s.log" 2>&1 &
    pids+=("$!:$s")
  fi
done
fail=0
for p in "${pids[@]}"; do
  if ! wait "${p%%:*}"; then echo "nvcc failed: ${p##*:}"; cat "build/${p##*:}.log"; fail=1; fi
done
[[ $fail -eq 0 ]] || exit 1
objs=(); for s in "${SRCS[@]}"; do objs+=("build/$s.o"); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o libsvol_b200.so "${objs[@]}" -lcudart_static -ldl -lrt -lpthread
echo "built $(pwd)/libsvol_b200.so"
