"""Scale search of reg_xyz (1000 anisotropic candidates x batched ICP + scoring): kernel path vs the torch formulation."""
import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200.reg_xyz import iterative_scale_search
from genpc_b200.synthetic import superquadric
from genpc_b200.utils.dataUtils import voxel_down_sample
dev = torch.device("cuda:0")
tc = torch.from_numpy(superquadric(0, 16384)).to(dev)
tgt_s = voxel_down_sample(tc, 0.03); src_s = voxel_down_sample(tc / torch.tensor([1.1, 1.0, 0.9], device=dev), 0.03)
out = {"src_pts": int(src_s.shape[0]), "tgt_pts": int(tgt_s.shape[0])}
for mode in ("kernel", "torch"):
    if mode == "torch": os.environ["GENPC_ICP_TORCH"] = "1"
    iterative_scale_search(src_s, tgt_s, [(0.8, 1.2)] * 3, 4, None, 0.5); torch.cuda.synchronize()
    t0 = time.perf_counter(); S, loss, Tb = iterative_scale_search(src_s, tgt_s, [(0.8, 1.2)] * 3, 10, None, 0.5); torch.cuda.synchronize()
    out[mode] = {"ms": round((time.perf_counter() - t0) * 1e3, 1), "best_scales": [float(v) for v in S.diagonal()[:3]], "loss": loss}
print(json.dumps(out))
