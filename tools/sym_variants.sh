#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GENPC_CHAMFER_MODE=sym
for ps in 0 1; do for span in 0 256 512 1024; do
  if [ $span = 0 ]; then unset GENPC_SYM_SPAN; else export GENPC_SYM_SPAN=$span; fi
  GENPC_SYM_PERSIST=$ps ./tools/variants/symp | sed "s/\"variant\": \"/\"variant\": \"SYM persist$ps span$span /"
done; done
