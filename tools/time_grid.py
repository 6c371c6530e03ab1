"""Large-cloud Chamfer forward: pruned grid scan (library default for these shapes) against the exhaustive symmetric scan
(GENPC_CHAMFER_PRUNE=0) on the BASELINE C5 generator (LiDAR-like scene pair) and the C1 shape; equality of all four outputs."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import _lib
from genpc_b200.loss_functions import chamfer_3DDist
from genpc_b200.synthetic import lidar_scene_pair, superquadric

dev = torch.device("cuda:0")
sizes = [int(v) for v in sys.argv[1:]] or [1000000]


def timed(fn, reps):
    ts = []
    for r in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    return out, {"best_ms": round(min(ts), 3), "median_ms": round(float(np.median(ts)), 3)}


out = {}
cases = [(f"lidar_{n}x{n}", *[t[None].to(dev) for t in lidar_scene_pair(n, 0)]) for n in sizes]
fix = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scan_01184_xyz.npz")
if os.path.exists(fix):
    cases.append(("C1_71372x16384", torch.from_numpy(np.load(fix)["xyz"])[None].to(dev), torch.from_numpy(superquadric(0, 16384))[None].to(dev)))
for name, a, b in cases:
    row = {}
    res = {}
    for knob in (None, "0"):
        with _lib.tunable(GENPC_CHAMFER_PRUNE=knob):
            cd = chamfer_3DDist()
            stats = torch.zeros(4, dtype=torch.int32, device=dev)
            _lib.lib().genpc_chamfer_prune_stats(_lib.ptr(stats))
            cd(a, b); torch.cuda.synchronize()
            _lib.lib().genpc_chamfer_prune_stats(None)
            r, t = timed(lambda: cd(a, b), 3 if knob == "0" and a.shape[1] > 300000 else 10)
            res[knob] = [x.clone() for x in r]
            row["exhaustive" if knob == "0" else "default"] = dict(t, stats_blocks_ties_groups_superblocks=stats.cpu().tolist())
    row["identical"] = all(bool(torch.equal(x, y)) for x, y in zip(res[None], res["0"]))
    out[name] = row
print(json.dumps(out, indent=1))
