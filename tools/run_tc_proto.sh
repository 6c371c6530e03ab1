#!/bin/bash
# r02: tensor-core filter prototype (tools/tc_filter_proto.cu) -- variants built here (nvcc cross-compiles), run on the GPU box.
#   tools/run_tc_proto.sh build      (CPU container)
#   tools/run_tc_proto.sh run        (GPU box; writes gpurun_out/tc_proto.txt)
set -u
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
if [ "${1:-build}" = build ]; then
  mkdir -p tools/bin
  $NV -o tools/bin/tc_proto_full tools/tc_filter_proto.cu || exit 1
  $NV -DPROTO_TRACK=0 -o tools/bin/tc_proto_notrack tools/tc_filter_proto.cu || exit 1
  $NV -DPROTO_MODE=1 -o tools/bin/tc_proto_ldonly tools/tc_filter_proto.cu || exit 1
  $NV -DPROTO_MODE=2 -o tools/bin/tc_proto_mmaonly tools/tc_filter_proto.cu || exit 1
  $NV -DPROTO_EPI_WARPS=16 -o tools/bin/tc_proto_full16 tools/tc_filter_proto.cu || exit 1
  $NV -DPROTO_EPI_WARPS=16 -DPROTO_MODE=1 -o tools/bin/tc_proto_ldonly16 tools/tc_filter_proto.cu || exit 1
  exit 0
fi
mkdir -p gpurun_out
OUT=gpurun_out/tc_proto.txt
: > $OUT
for v in mmaonly ldonly ldonly16 notrack full full16; do
  echo "== $v" >> $OUT
  timeout 120 tools/bin/tc_proto_$v 32 20 >> $OUT 2>&1 || echo "FAILED rc=$?" >> $OUT
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv >> $OUT
timeout 300 ncu --set full --clock-control none --import-source on -s 3 -c 1 -o gpurun_out/r02_tc_proto tools/bin/tc_proto_full 32 1 >> $OUT 2>&1
tail -5 $OUT
