#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants gpurun_out
if [ "$1" = "build" ]; then
  rm -f tools/variants/sym*
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -o tools/variants/symd tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A2 "nn_sym_kernelILi6" | grep -E "registers|spill" | head -2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -DGENPC_SYM_MINB4=1 -o tools/variants/symd1 tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A2 "nn_sym_kernelILi6" | grep -E "registers|spill" | head -2
else
  export GENPC_CHAMFER_MODE=sym
  ./tools/variants/symd | sed "s/\"variant\": \"/\"variant\": \"SYM qt4 minb2 /"
  GENPC_SYM_QT=6 ./tools/variants/symd | sed "s/\"variant\": \"/\"variant\": \"SYM qt6 minb2 /"
  GENPC_SYM_QT=6 ./tools/variants/symd1 | sed "s/\"variant\": \"/\"variant\": \"SYM qt6 minb1 /"
  GENPC_SYM_QT=4 ./tools/variants/symd1 | sed "s/\"variant\": \"/\"variant\": \"SYM qt4 minb1 /"
fi
