#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants gpurun_out
if [ "$1" = "build" ]; then
  for v in "2:3" "1:2" "2:2"; do IFS=: read m8 m4 <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -DGENPC_SYM_MINB8=$m8 -DGENPC_SYM_MINB4=$m4 \
      -o tools/variants/sym_${m8}_${m4} tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A2 "nn_sym_kernel" | grep -E "registers|spill" | head -6 &
  done; wait
else
  export GENPC_CHAMFER_MODE=scan; ./tools/variants/sym_2_3 | sed 's/"variant": "/"variant": "SCAN /'
  export GENPC_CHAMFER_MODE=sym
  for bin in sym_2_3 sym_1_2 sym_2_2; do for qt in 8 4; do for span in 128 256 512 1024; do
     GENPC_SYM_QT=$qt GENPC_SYM_SPAN=$span ./tools/variants/$bin | sed "s/\"variant\": \"/\"variant\": \"SYM $bin qt$qt span$span /"
  done; done; done
fi
