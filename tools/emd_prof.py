import os, sys, torch
sys.path.insert(0, "/root/repo")
from genpc_b200 import emd as ours
dev = torch.device("cuda:0")
iters = int(sys.argv[1])
B, n = 32, 8192
g = torch.Generator().manual_seed(0)
x1, x2 = torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, n, 3, generator=g).to(dev)
for rep in range(2):
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
    ours.forward(x1, x2, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32), z(B, n), z(B, n),
                 z(B * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(B * n, dt=torch.int32), 0.005, iters)
torch.cuda.synchronize()
