"""Probe: one call of the region-growing normals (200 000 points, k = 2000) after a warm-up."""
import importlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

gs = importlib.import_module("3d_gaussian_splatting_project_b200")
rg = importlib.import_module("3d_gaussian_splatting_project_b200.region_growing")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
d = torch.from_numpy(gs.scene.gaussian_cloud(n, 1.5, seed=4)).cuda()
for _ in range(2):
    rg.knn_pca(d, k)
torch.cuda.synchronize()
print("probe done")
