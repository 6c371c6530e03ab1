#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants gpurun_out
VARS="8:1 16:1 32:1 8:2 16:2 32:2"
if [ "$1" = "build" ]; then
  rm -f tools/variants/sym*
  for v in $VARS; do IFS=: read ch rx <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -DGENPC_SYM_CHUNK=$ch -DGENPC_SYM_REDUX=$rx \
      -o tools/variants/syme_${ch}_${rx} tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A2 "nn_sym_kernelILi4" | grep -E "registers|spill" | head -2 &
  done; wait
else
  export GENPC_CHAMFER_MODE=sym
  for v in $VARS; do IFS=: read ch rx <<< "$v"
     ./tools/variants/syme_${ch}_${rx} | sed "s/\"variant\": \"/\"variant\": \"SYM chunk$ch redux$rx /"
  done
fi
