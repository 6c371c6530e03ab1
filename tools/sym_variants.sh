#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants gpurun_out
if [ "$1" = "build" ]; then
  rm -f tools/variants/symv_*
  for m8 in 2 1; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -DGENPC_SYM_MINB8=$m8 \
      -o tools/variants/symq_${m8} tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A2 "nn_sym_kernelILi8" | grep -E "registers|spill" | head -2 &
  done; wait
else
  export GENPC_CHAMFER_MODE=sym
  ./tools/variants/symq_2 | sed "s/\"variant\": \"/\"variant\": \"SYM default(qt4 redux minb2) /"
  for m8 in 2 1; do GENPC_SYM_QT=8 ./tools/variants/symq_${m8} | sed "s/\"variant\": \"/\"variant\": \"SYM qt8 redux minb8=$m8 /"; done
  GENPC_SYM_SPAN=512 ./tools/variants/symq_2 | sed "s/\"variant\": \"/\"variant\": \"SYM default span512 /"
fi
