"""Probe: launches of the viewer depth sort / hit test at 6 M Gaussians and of the region-growing
normals at 200 000 points, k = 2000 (run under `ncu --metrics gpu__time_duration.sum` for the launch list)."""
import importlib
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

gs = importlib.import_module("3d_gaussian_splatting_project_b200")
rg = importlib.import_module("3d_gaussian_splatting_project_b200.region_growing")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
pos = gs.scene.gaussian_cloud(n, 1.5, seed=4)
d = torch.from_numpy(pos).cuda()
lab = torch.zeros(n, dtype=torch.int32, device="cuda")
view = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 6.0, 1.0]
proj = [1.2, 0, 0, 0, 0, 1.2 * 16 / 9, 0, 0, 0, 0, 1.01, 1.0, 0, 0, -0.2, 0]
vp = gs.viewer.multiply4(proj, view)
for _ in range(2):
    gs.viewer.run_sort(d, vp)
    gs.viewer.perform_hit_testing(960, 540, view, proj, (1920, 1080), d, lab)
    rg.knn_pca(d[:200_000].contiguous(), 2000)
torch.cuda.synchronize()
print("probe done")
