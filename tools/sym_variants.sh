#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants gpurun_out
VARS="8:0:3 8:1:3 16:0:3 16:1:3 8:1:2 16:0:2"
if [ "$1" = "build" ]; then
  for v in $VARS; do IFS=: read ch rx m4 <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -DGENPC_SYM_CHUNK=$ch -DGENPC_SYM_REDUX=$rx -DGENPC_SYM_MINB4=$m4 \
      -o tools/variants/symv_${ch}_${rx}_${m4} tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A2 "nn_sym_kernelILi4" | grep -E "registers|spill" | head -2 &
  done; wait
else
  export GENPC_CHAMFER_MODE=sym
  for v in $VARS; do IFS=: read ch rx m4 <<< "$v"
     ./tools/variants/symv_${ch}_${rx}_${m4} | sed "s/\"variant\": \"/\"variant\": \"SYM chunk$ch redux$rx minb$m4 /"
  done
fi
