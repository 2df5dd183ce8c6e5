"""Probe: the fused K-means step (assignment + per-cluster float64 sums) at the C5 shape."""
import importlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

gs = importlib.import_module("3d_gaussian_splatting_project_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
data = gs.scene.blob_features(n, 59, seed=5)
np.random.seed(0)
cen = data[np.random.choice(n, 64, replace=False)]
d, c = torch.from_numpy(data).cuda(), torch.from_numpy(cen).cuda()
for _ in range(3):
    lab, sums = gs.ops.kmeans_step(d, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gs.ops.kmeans_step(d, c)
e1.record()
torch.cuda.synchronize()
print("kmeans_step ms", e0.elapsed_time(e1) / 10)
