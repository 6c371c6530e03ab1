import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200.fps import furthest_point_sample
from genpc_b200 import _lib
def setk(name, value):   # the library reads the environment once at load time: flip knobs through the C ABI
    _lib.check(_lib.lib().genpc_set_tunable(name.encode(), None if value is None else str(value).encode()), name)
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
        if mode == "default": setk("GENPC_FPS_MODE", None)
        else: setk("GENPC_FPS_MODE", mode)
        row[mode] = round(ev(lambda: furthest_point_sample(x, 2048, 0)), 3)
    out[f"B{B}_16384->2048_ms"] = row
setk("GENPC_FPS_MODE", None)
print(json.dumps(out))
