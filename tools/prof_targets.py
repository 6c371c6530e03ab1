"""One small workload per kernel family, for ncu:  python tools/prof_targets.py {register|emd|fps|depth|knn|icp}"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
from genpc_b200 import _lib
def setk(name, value):   # the library reads the environment once at load time: flip knobs through the C ABI
    _lib.check(_lib.lib().genpc_set_tunable(name.encode(), None if value is None else str(value).encode()), name)
what = sys.argv[1]
g = torch.Generator().manual_seed(0)
if what == "chamfer":   # C2 forward (tensor-core filter by default), three calls
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.synthetic import pcn_batch
    part, comp = pcn_batch(0, 32, 2048, 16384)
    a, b = torch.from_numpy(part).to(dev), torch.from_numpy(comp).to(dev)
    for _ in range(3):
        chamfer_3DDist()(a, b)
elif what == "register":
    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch
    from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric
    import numpy as np
    comp = np.stack([superquadric(s, 16384) for s in range(16)])
    part = np.stack([rigid_perturb(partial_view(comp[s], s, 16384), s)[0] for s in range(16)])
    rb = RegistrationBatch(torch.from_numpy(comp).to(dev), torch.from_numpy(part).to(dev), n_starts=1, max_iters=8)
    rb.run(4)
elif what == "emd":
    from genpc_b200.loss_functions import emdModule
    x, y = torch.rand(8, 8192, 3, generator=g).to(dev), torch.rand(8, 8192, 3, generator=g).to(dev)
    emdModule()(x, y, 0.005, 50); emdModule()(x, y, 0.005, 50)
elif what == "fps":
    from genpc_b200.fps import furthest_point_sample
    x = torch.rand(1, 16384, 3, generator=g).to(dev)
    furthest_point_sample(x, 2048, 0); furthest_point_sample(x, 2048, 0)
    setk("GENPC_FPS_MODE", "cta")
    furthest_point_sample(x, 2048, 0)
elif what == "depth":
    from genpc_b200 import depth as D
    from genpc_b200.synthetic import superquadric
    pts = torch.from_numpy(superquadric(0, 71372)).to(dev)
    cams, _ = D.create_cameras(8, 1.6, 49.1, 512, dev)
    for _ in range(2):
        ndc, uv, b = D.project_uv(cams, pts, True, 0.15); r = D.zbuffer_render(uv, ndc, 512, 2); D.unproject(cams, b, r["zbuf"], ndc, True)
elif what == "knn":
    from genpc_b200.reg_xyz import knn_mean_distance
    from genpc_b200.synthetic import superquadric
    pts = torch.from_numpy(superquadric(0, 20000)).to(dev)
    knn_mean_distance(pts, 20, True); knn_mean_distance(pts, 20, True)
elif what == "icp":
    from genpc_b200.reg_xyz import icp_point_to_point
    from genpc_b200.synthetic import superquadric
    tgt = torch.from_numpy(superquadric(0, 1800)).to(dev)
    src = (tgt[None] * torch.linspace(0.8, 1.2, 250, device=dev)[:, None, None]).contiguous()
    icp_point_to_point(src, tgt[None], 0.075, max_iteration=4)
torch.cuda.synchronize()
print("done", what)
