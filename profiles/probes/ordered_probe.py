"""Probe: the reference-order K-means update (ordered chain kernel) at the C5 shape."""
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
lab = gs.ops.kmeans_assign(d, c)
print("largest cluster", int(torch.bincount(lab.long(), minlength=64).max()))
for _ in range(2):
    gs.ops.kmeans_update_ordered(d, lab, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
gs.ops.kmeans_update_ordered(d, lab, c)
e1.record()
torch.cuda.synchronize()
print("update_ordered ms", e0.elapsed_time(e1))
