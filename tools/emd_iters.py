"""EMD cost per auction iteration: time of genpc_emd_forward for iters = 1..50 (ours), final unassigned count.
args: BxN shapes; EMD_ITERS=1,2,50 picks the iteration counts; GENPC_EMD_PRUNE=0/1 selects the Bid form."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import _lib
if os.environ.get("GENPC_LIB"):   # A/B of two builds on the same box
    _lib.LIB_PATH = os.path.abspath(os.environ["GENPC_LIB"])
from genpc_b200 import emd as ours
SHAPES = [(1, 8192), (32, 8192)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
ITERS = (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 50) if not os.environ.get("EMD_ITERS") else tuple(int(v) for v in os.environ["EMD_ITERS"].split(","))
dev = torch.device("cuda:0")
out = {}
for (B, n) in SHAPES:
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, n, 3, generator=g).to(dev)
    res = {}
    for iters in ITERS:
        ts = []
        for rep in range(3):
            dist = torch.zeros(B, n, device=dev); asg = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1
            asg_inv = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1; price = torch.zeros(B, n, device=dev)
            bid = torch.zeros(B, n, device=dev, dtype=torch.int32); binc = torch.zeros(B, n, device=dev)
            minc = torch.zeros(B, n, device=dev); uidx = torch.zeros(B * n, device=dev, dtype=torch.int32)
            midx = torch.zeros(B * n, device=dev, dtype=torch.int32)
            z = [torch.zeros(512, dtype=torch.int32, device=dev) for _ in range(3)]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ours.forward(x1, x2, dist, asg, price, asg_inv, bid, binc, minc, uidx, z[0], z[1], z[2], midx, 0.005, iters)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[str(iters)] = {"ms": round(min(ts[1:]), 4), "unassigned_before_last_iter_mean": float(z[0][:B].float().mean())}
    out[f"B{B}_n{n}"] = res
print(json.dumps(out, indent=1))
