"""Phase timeline of the EMD auction (cloud 0): Bid / GetMax+Assign / compaction / flag hand-over per iteration.
Needs the tracing build:  nvcc <flags of csrc/build.py> -DGENPC_EMD_TRACE -shared -o tools/bin/libgenpc_trace.so genpc_b200/csrc/*.cu
usage: python tools/emd_trace.py B n [iters]"""
import ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genpc_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "bin", "libgenpc_trace.so")
from genpc_b200 import emd as ours
B, n = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x1, x2 = torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, n, 3, generator=g).to(dev)
for rep in range(3):
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    cnt = z(512, dt=torch.int32)
    ours.forward(x1, x2, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32), z(B, n), z(B, n),
                 z(B * n, dt=torch.int32), cnt, z(512, dt=torch.int32), z(512, dt=torch.int32), z(B * n, dt=torch.int32), 0.005, iters)
    torch.cuda.synchronize()
L = _lib.lib()
L.genpc_emd_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(8 * iters, np.uint64)
assert L.genpc_emd_trace_read(buf.ctypes.data, 8 * iters) == 0
t = buf.reshape(iters, 8).astype(np.int64)
rows = []
for it in range(iters):
    nxt = t[it + 1, 0] if it + 1 < iters else t[it, 3]
    rows.append(dict(it=it, bid_us=(t[it, 1] - t[it, 0]) / 1e3, getmax_assign_us=(t[it, 2] - t[it, 1]) / 1e3,
                     compact_release_us=(t[it, 3] - t[it, 2]) / 1e3, handover_us=(nxt - t[it, 3]) / 1e3,
                     b0_setup_us=(t[it, 4] - t[it, 0]) / 1e3, b0_coords_us=(t[it, 5] - t[it, 4]) / 1e3,
                     b0_first_threshold_us=(t[it, 6] - t[it, 5]) / 1e3, b0_rest_us=(t[it, 7] - t[it, 6]) / 1e3,
                     b0_end_to_all_done_us=(t[it, 1] - t[it, 7]) / 1e3))
late = rows[len(rows) // 2:]
mean = {k: round(float(np.mean([r[k] for r in late])), 2) for k in rows[0] if k != "it"}
print(json.dumps({"shape": f"B{B}_n{n}", "first_iterations": rows[:4], "late_mean_us": mean}, indent=1))
