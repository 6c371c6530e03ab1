#!/bin/bash
# 8-GPU session: scaling of the headline bench (both arms at 8) and of the sharded 1M x 1M Chamfer (1 / 2 GPUs: tools/gpu_n2.sh).
set +e
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpus.txt
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --impl reference --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/sharded_bench.py --npoints 1000000 > gpurun_out/sharded_n$n.json 2> gpurun_out/sharded_n$n.err
done
for f in bench_n8 bench_n4 bench_ref_n8 sharded_n8 sharded_n4; do echo "== $f"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version\|Loaded compiled" gpurun_out/$f.json | head -c 1300; echo; grep -v "^\*\|OMP_NUM\|^$\|Setting\|Loaded compiled" gpurun_out/$f.err | tail -3; done
exit 0
