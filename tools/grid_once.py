import sys, torch
sys.path.insert(0, "/root/repo")
from genpc_b200.loss_functions import chamfer_3DDist
from genpc_b200.synthetic import lidar_scene_pair
n = int(sys.argv[1])
a, b = [t[None].cuda() for t in lidar_scene_pair(n, 0)]
cd = chamfer_3DDist()
for r in range(2):
    cd(a, b)
torch.cuda.synchronize()
