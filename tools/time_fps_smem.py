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
setk("GENPC_FPS_MODE", "cta")
for (B, N, K) in [(32, 16384, 2048), (1, 16384, 2048), (32, 8192, 1024), (32, 5000, 512)]:
    x = torch.rand(B, N, 3, device=dev)
    for smem in ("0", "1"):
        setk("GENPC_FPS_SMEM", smem)
        out[f"B{B}_{N}->{K}_{'smem' if smem == '1' else 'l1'}_ms"] = round(ev(lambda: furthest_point_sample(x, K, 0)), 3)
print(json.dumps(out, indent=1))
