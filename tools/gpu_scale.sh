#!/bin/bash
# 8-GPU session: scaling of the headline bench and of the sharded 1M x 1M Chamfer.
set +e
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpus.txt
timeout 300 python -m pytest tests/test_sharded.py -m gpu -x -q --timeout 200 > gpurun_out/pytest_sharded.log 2>&1; tail -3 gpurun_out/pytest_sharded.log
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n$n.json 2> gpurun_out/sharded_n$n.err
done
timeout 300 python tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n1.json 2> gpurun_out/sharded_n1.err
for f in bench_n8 bench_n4 sharded_n8 sharded_n4 sharded_n2 sharded_n1; do echo "== $f"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$f.json | head -c 1800; echo; grep -v "^\*\|OMP_NUM\|^$\|Setting" gpurun_out/$f.err | tail -3; done
