#!/bin/bash
# 2-GPU check of the end-to-end (host-fed) step: the bench as the driver runs it, then short runs with knobs.
mkdir -p gpurun_out; nproc
run() { tag=$1; extra=$2; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $extra > gpurun_out/n2dbg_$tag.json 2> gpurun_out/n2dbg_$tag.err
python -c "
import json; j=json.loads(open('gpurun_out/n2dbg_$tag.json').read().strip().splitlines()[-1]); print('$tag', j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e_plain']['ms_per_step'])"; }
run full "" X=1
run short1 "--no-extras --no-cpu-baseline" X=1
run short2 "--no-extras --no-cpu-baseline" X=1
run gated "--no-extras --no-cpu-baseline" GENPC_HOST_PRUNE=0
