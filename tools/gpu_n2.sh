#!/bin/bash
# 2-GPU session: multi-rank sanity of the headline bench (both arms) and the sharded 1M x 1M Chamfer with phase times.
set +e
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n2.json 2> gpurun_out/sharded_n2.err
timeout 300 python tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n1.json 2> gpurun_out/sharded_n1.err
for f in bench_n2 bench_ref_n2 sharded_n2 sharded_n1; do echo "== $f"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$f.json | head -c 1500; echo; grep -v "^\*\|OMP_NUM\|^$\|Setting" gpurun_out/$f.err | tail -3; done
