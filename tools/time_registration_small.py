"""Small-cloud regime of the real pipeline (about 1-3 K points after voxel down-sampling, 4 starts): ms for 201 Adam iterations
of the single-launch registration path, for the queries-per-thread choices (GENPC_REGISTER_QT)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch
from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric
dev = torch.device("cuda:0")
out = {}
for (nc, nr) in [(2500, 1000), (1200, 800), (3000, 3000)]:
    comp = superquadric(0, nc)[None]; part = rigid_perturb(partial_view(superquadric(0, 16384), 0, nr), 0)[0][None]
    tc, tp = torch.from_numpy(comp).to(dev), torch.from_numpy(part).to(dev)
    for qt in ("auto", "1", "2", "4"):
        if qt == "auto": os.environ.pop("GENPC_REGISTER_QT", None)
        else: os.environ["GENPC_REGISTER_QT"] = qt
        rb = RegistrationBatch(tc, tp, n_starts=4, lr=0.01, max_iters=300)
        rb.run(10); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rb.run(201); e1.record(); torch.cuda.synchronize()
        out[f"{nc}x{nr}_qt{qt}"] = {"ms_201_iters": round(e0.elapsed_time(e1), 3), "final_loss_sum": float(rb.losses()[:, rb.t - 1].double().sum())}
os.environ.pop("GENPC_REGISTER_QT", None)
print(json.dumps(out, indent=1))
