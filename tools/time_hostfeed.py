"""Host-fed Chamfer forward (genpc_chamfer_forward_host) vs copy-then-compute, C2 shape, by number of chunks."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import chamfer_3D  # noqa: E402
from genpc_b200.synthetic import pcn_batch  # noqa: E402

dev = torch.device("cuda:0")
B, N, M = 32, 2048, 16384
part, comp = pcn_batch(0, B, N, M)
ha, hb = torch.from_numpy(part).pin_memory(), torch.from_numpy(comp).pin_memory()
xa, xb = torch.empty(B, N, 3, device=dev), torch.empty(B, M, 3, device=dev)
d1, d2 = torch.empty(B, N, device=dev), torch.empty(B, M, device=dev)
i1, i2 = torch.empty(B, N, dtype=torch.int32, device=dev), torch.empty(B, M, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=30, warm=5):
    ts = []
    for r in range(reps + warm):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if r >= warm:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return {"min_ms": ts[0], "median_ms": ts[len(ts) // 2]}


def plain():
    xa.copy_(ha, non_blocking=True); xb.copy_(hb, non_blocking=True)
    chamfer_3D.forward(xa, xb, d1, d2, i1, i2)


out = {"shape": [B, N, M], "h2d_bytes": ha.numel() * 4 + hb.numel() * 4}
out["copy_only"] = timed(lambda: (xa.copy_(ha, non_blocking=True), xb.copy_(hb, non_blocking=True)))
out["resident_forward"] = timed(lambda: chamfer_3D.forward(xa, xb, d1, d2, i1, i2))
out["copy_then_forward"] = timed(plain)
for chunks in (2, 3, 4, 6, 8, 16):
    out[f"host_fed_chunks{chunks}"] = timed(lambda: chamfer_3D.forward_host(ha, hb, xa, xb, d1, d2, i1, i2, chunks))
from genpc_b200 import _lib  # noqa: E402
with _lib.tunable(GENPC_HOST_PRUNE="1"):   # sort + pruned scan per chunk behind the chunk's copy, one stream per chunk
    for chunks in (2, 3, 4, 6, 8):
        out[f"host_fed_pruned_chunks{chunks}"] = timed(lambda: chamfer_3D.forward_host(ha, hb, xa, xb, d1, d2, i1, i2, chunks))
print(json.dumps(out, indent=1))
