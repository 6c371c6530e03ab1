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
for (N, K) in [(16384, 2048), (32768, 4096), (45000, 10000), (71372, 10000), (139138, 16384)]:
    x = torch.rand(1, N, 3, device=dev)
    for c16 in ("0", "1"):
        setk("GENPC_FPS_MODE", "cluster"); setk("GENPC_FPS_CLUSTER16", c16)
        out[f"{N}->{K} cluster{'16' if c16 == '1' else '8'}_ms"] = round(ev(lambda: furthest_point_sample(x, K, 0)), 3)
print(json.dumps(out, indent=1))
