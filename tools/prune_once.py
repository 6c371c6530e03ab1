import sys, torch
sys.path.insert(0, "/root/repo")
from genpc_b200 import _lib
from genpc_b200.loss_functions import chamfer_3DDist
from genpc_b200.synthetic import pcn_batch
B, N, M = [int(v) for v in sys.argv[1].split("x")]
a, b = pcn_batch(int(sys.argv[2]) if len(sys.argv) > 2 else 0, B, N, M)
dev = torch.device("cuda:0")
ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
_lib.lib().genpc_set_tunable(b"GENPC_CHAMFER_PRUNE", b"1")
cd = chamfer_3DDist()
for r in range(3):
    cd(ta, tb)
torch.cuda.synchronize()
