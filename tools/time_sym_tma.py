"""r02 experiment: TMA bulk staging of the column span (nn_sym_tma_kernel, GENPC_SYM_TMA=1) against the
shipped LDG -> STS staging (nn_sym_kernel) on the C2 forward: bit-exactness + CUDA-event time, L2 flushed between runs."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import _lib, chamfer_3D
from genpc_b200.synthetic import pcn_batch
dev = torch.device("cuda:0")
part, comp = pcn_batch(0, 32, 2048, 16384)
a, b = torch.from_numpy(part).to(dev), torch.from_numpy(comp).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
ref = None
for mode in ("ldg_sts", "tma_bulk", "ldg_sts", "tma_bulk"):
    with _lib.tunable(GENPC_SYM_TMA="1" if mode == "tma_bulk" else None):
        d1 = torch.empty(32, 2048, device=dev); d2 = torch.empty(32, 16384, device=dev)
        i1 = torch.empty(32, 2048, dtype=torch.int32, device=dev); i2 = torch.empty(32, 16384, dtype=torch.int32, device=dev)
        ts = []
        for it in range(25):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); chamfer_3D.forward(a, b, d1, d2, i1, i2); e1.record(); torch.cuda.synchronize()
            if it >= 5: ts.append(e0.elapsed_time(e1))
        res = (d1.clone(), d2.clone(), i1.clone(), i2.clone())
        if ref is None: ref = res
        same = all(torch.equal(x, y) for x, y in zip(ref, res))
        out.setdefault(mode, []).append({"fwd_ms_mean": round(sum(ts) / len(ts), 4), "fwd_ms_min": round(min(ts), 4), "bit_identical": same})
print(json.dumps(out, indent=1))
