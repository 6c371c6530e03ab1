#!/bin/bash
# 2-GPU session: bench ours/reference at N=1 and N=2, sharded chamfer 1M x 1M at N=2, NCCL merge check.
set +e
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt
timeout 600 python bench.py --impl reference --steps 30 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 20 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n2.json 2> gpurun_out/sharded_n2.err
timeout 600 python tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n1.json 2> gpurun_out/sharded_n1.err
for f in bench_ref bench bench_n2 bench_ref_n2 sharded_n2 sharded_n1; do echo "== $f"; head -c 2500 gpurun_out/$f.json; echo; tail -3 gpurun_out/$f.err; done
