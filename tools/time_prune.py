"""Spatially pruned Chamfer scan (GENPC_CHAMFER_PRUNE=1) against the exhaustive symmetric scan: forward and fused loss step on
BASELINE C2 (and other shapes given as BxNxM), L2 flushed between timed calls, best and median of 20."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import _lib
if os.environ.get("GENPC_LIB"):   # A/B of two builds on the same box
    _lib.LIB_PATH = os.path.abspath(os.environ["GENPC_LIB"])
from genpc_b200.loss_functions import chamfer_3DDist
from genpc_b200.synthetic import pcn_batch
from genpc_b200.utils.loss_util import Completionloss

dev = torch.device("cuda:0")
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(32, 2048, 16384)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    ts = []
    for r in range(reps + 3):
        flush.fill_(r & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1))
    return {"best_ms": round(min(ts), 4), "median_ms": round(float(np.median(ts)), 4)}


out = {}
for (B, N, M) in shapes:
    a, b = pcn_batch(0, B, N, M)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    row = {}
    for prune in ("0", "1"):
        with _lib.tunable(GENPC_CHAMFER_PRUNE=prune):
            cd = chamfer_3DDist()
            fwd = timed(lambda: cd(ta, tb))
            crit = Completionloss("cd_l2")
            ga, gb = ta.clone().requires_grad_(True), tb.clone().requires_grad_(True)

            def step():
                ga.grad = gb.grad = None
                crit.get_loss(ga, gb).backward()
            st = timed(step)
            stats = torch.zeros(4, dtype=torch.int32, device=dev)
            _lib.lib().genpc_chamfer_prune_stats(_lib.ptr(stats))
            cd(ta, tb); torch.cuda.synchronize()
            _lib.lib().genpc_chamfer_prune_stats(None)
            row["pruned" if prune == "1" else "exhaustive"] = {"forward": fwd, "loss_step": st, "stats_blocks_ties_groups": stats.cpu().tolist()[:3]}
    g = lambda n: (n + 31) // 32
    k = lambda n: (n + 63) // 64
    row["group_block_pairs_total"] = B * (g(N) * k(M) + g(M) * k(N))
    out[f"{B}x{N}x{M}"] = row
print(json.dumps(out, indent=1))
