"""Small-cloud regime of the real pipeline (about 1-3 K points after voxel down-sampling, 4 starts): ms for 201 Adam iterations,
launch-per-iteration (r01) against the persistent cooperative kernel (r02: all iterations in one launch)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genpc_b200 import _lib
from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch
from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric
dev = torch.device("cuda:0")
out = {}
for (nc, nr) in [(2500, 1000), (1200, 800), (3000, 3000), (924, 2500)]:
    comp = superquadric(0, nc)[None]; part = rigid_perturb(partial_view(superquadric(0, 16384), 0, nr), 0)[0][None]
    tc, tp = torch.from_numpy(comp).to(dev), torch.from_numpy(part).to(dev)
    for mode in ("persistent", "launch_per_iter"):
        with _lib.tunable(GENPC_REGISTER_PERSIST="1" if mode == "persistent" else "0"):
            rb = RegistrationBatch(tc, tp, n_starts=4, lr=0.01, max_iters=300)
            rb.run(10); torch.cuda.synchronize()
            ts = []
            for rep in range(3):
                rb2 = RegistrationBatch(tc, tp, n_starts=4, lr=0.01, max_iters=300)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); rb2.run(201); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        out[f"{nc}x{nr}_x4starts_{mode}"] = {"ms_201_iters": round(min(ts), 3), "us_per_iter": round(min(ts) / 201 * 1e3, 2),
                                            "final_loss_sum": float(rb2.losses()[:, rb2.t - 1].double().sum())}
print(json.dumps(out, indent=1))
