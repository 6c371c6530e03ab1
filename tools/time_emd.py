"""EMD timing: ours vs the unmodified reference extension (oracle/_ref) on the same GPU."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from genpc_b200 import emd as ours
dev = torch.device("cuda:0")
ref = oracle.load_ref_ext("emd")
out = {}
for (B, n) in [(1, 8192), (32, 8192), (1, 16384), (20, 2048)]:
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, n, 3, generator=g).to(dev)
    res = {}
    for name, mod in (("ours", ours), ("reference_ext", ref)):
        if mod is None:
            continue
        ts = []
        for rep in range(4):
            dist = torch.zeros(B, n, device=dev); asg = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1
            asg_inv = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1; price = torch.zeros(B, n, device=dev)
            bid = torch.zeros(B, n, device=dev, dtype=torch.int32); binc = torch.zeros(B, n, device=dev)
            minc = torch.zeros(B, n, device=dev); uidx = torch.zeros(B * n, device=dev, dtype=torch.int32)
            midx = torch.zeros(B * n, device=dev, dtype=torch.int32)
            z = [torch.zeros(512, dtype=torch.int32, device=dev) for _ in range(3)]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mod.forward(x1, x2, dist, asg, price, asg_inv, bid, binc, minc, uidx, z[0], z[1], z[2], midx, 0.005, 50)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name + "_ms"] = min(ts[1:])
        res[name + "_emd"] = float(torch.sqrt(dist).mean())
    if "reference_ext_ms" in res:
        res["speedup"] = res["reference_ext_ms"] / res["ours_ms"]
    out[f"B{B}_n{n}_eps0.005_it50"] = res
print(json.dumps(out, indent=1))
