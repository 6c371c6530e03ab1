import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200.fps import furthest_point_sample
dev = torch.device("cuda:0")
out = {}
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for B in (1, 8, 16, 17, 18, 19, 32):
    x = torch.rand(B, 16384, 3, device=dev)
    row = {}
    for mode in ("default", "cluster", "cta"):
        if mode == "default": os.environ.pop("GENPC_FPS_MODE", None)
        else: os.environ["GENPC_FPS_MODE"] = mode
        row[mode] = round(ev(lambda: furthest_point_sample(x, 2048, 0)), 3)
    out[f"B{B}_16384->2048_ms"] = row
os.environ.pop("GENPC_FPS_MODE", None)
print(json.dumps(out))
