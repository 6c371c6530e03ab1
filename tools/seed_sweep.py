"""C2 fused loss step (CUDA graph) over eight different batches of the bench's generator: how much the pruned scan's time depends
on the data (the exhaustive scan's does not)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import _lib
from genpc_b200.synthetic import pcn_batch
from genpc_b200.utils.loss_util import Completionloss, GraphedLossStep
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cl = Completionloss("cd_l2")
for seed0 in (0, 1000, 2000, 3000, 4000, 5000, 6000, 7000):       # bench.py: rank r uses pcn_batch(1000 * r, ...)
    a, b = pcn_batch(seed0, 32, 2048, 16384)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    row = {}
    for name, knobs in (("pruned", {}), ("pruned_one_launch", {"GENPC_PRUNE_COOP": "0"}), ("exhaustive", {"GENPC_CHAMFER_PRUNE": "0"})):
        with _lib.tunable(**knobs):
            st = GraphedLossStep(cl, ta, tb)
            ts = []
            for r in range(25):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); st(); e1.record(); torch.cuda.synchronize()
                if r >= 5: ts.append(e0.elapsed_time(e1))
            row[name] = round(float(np.mean(ts)), 4)
    print(seed0, row)
