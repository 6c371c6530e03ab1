#!/bin/bash
# build (here, no GPU needed) or run (on the GPU box) the NN work-item variants
set -e
cd "$(dirname "$0")/.."
VARIANTS="2048:8:2:4 2048:8:3:2 2048:4:2:4 4096:8:2:4 1024:4:2:4 2048:8:3:4 4096:8:3:2 4096:4:2:4"
mkdir -p tools/variants gpurun_out
if [ "$1" = "build" ]; then
  for v in $VARIANTS; do IFS=: read sp ch mb qt <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -DGENPC_NN_SPAN=$sp -DGENPC_NN_CHUNK=$ch -DGENPC_NN_MINBLOCKS=$mb -DGENPC_NN_QT_MAX=$qt \
      -o tools/variants/nn_${sp}_${ch}_${mb}_${qt} tools/nn_variants.cu -Xptxas -v 2>&1 | grep -A1 "nn_scan_kernelILi$qt" | grep -E "registers|spill" | head -2 &
  done; wait
else
  for v in $VARIANTS; do IFS=: read sp ch mb qt <<< "$v"; ./tools/variants/nn_${sp}_${ch}_${mb}_${qt}; done | tee gpurun_out/nn_variants.jsonl
fi
