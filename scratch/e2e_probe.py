import importlib, time, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
gs = importlib.import_module("3d_gaussian_splatting_project_b200")
ops, dls = gs.ops, gs.deep_learning_segmentation
V, H, W, N = 300, 1080, 1920, 6_000_000
cams = gs.scene.lookat_cameras(V, width=W, height=H, seed=4)
pos = gs.scene.gaussian_cloud(N, 1.5, seed=4)
maps_t = torch.empty((V, H, W), dtype=torch.int32, pin_memory=True)
gs.scene.block_label_maps(V, H, W, block=32, seed=1000, out=maps_t.numpy())
pos_p = torch.from_numpy(pos).pin_memory()
seg = [maps_t[v] for v in range(V)]
dev = torch.device("cuda")
def sync(): torch.cuda.synchronize()
for rep in range(3):
    sync(); t0 = time.perf_counter()
    out = dls.lift_labels(pos_p, cams, seg, None, device=dev)
    sync(); t1 = time.perf_counter()
    print("lift_labels total ms", (t1 - t0) * 1e3)
# breakdown
sync(); t0 = time.perf_counter()
shapes = [tuple(m.shape) for m in seg]
views = ops.make_views(cams, shapes, None)
t1 = time.perf_counter()
staged = torch.empty(V * H * W, dtype=torch.int32, device=dev)
sync(); t2 = time.perf_counter()
off = 0
for m in seg:
    staged[off:off + H * W].copy_(m.reshape(-1), non_blocking=True); off += H * W
t3 = time.perf_counter(); sync(); t4 = time.perf_counter()
lo, hi = ops.label_range(staged); sync(); t5 = time.perf_counter()
packed = ops.pack_labels(staged, lo, hi - lo + 1); sync(); t6 = time.perf_counter()
d_pos = pos_p.to(dev, non_blocking=True); sync(); t7 = time.perf_counter()
lab = ops.lift_votes(d_pos, views, packed, lo, hi - lo + 1); sync(); t8 = time.perf_counter()
h = lab.cpu().numpy(); t9 = time.perf_counter()
one = maps_t.reshape(-1).to(dev, non_blocking=True); sync(); t10 = time.perf_counter()
print(f"make_views {1e3*(t1-t0):.1f} alloc {1e3*(t2-t1):.1f} enqueue300copies {1e3*(t3-t2):.1f} copies_done {1e3*(t4-t3):.1f} range {1e3*(t5-t4):.1f} pack {1e3*(t6-t5):.1f} pos {1e3*(t7-t6):.1f} lift {1e3*(t8-t7):.1f} d2h {1e3*(t9-t8):.1f} single_big_copy {1e3*(t10-t9):.1f}")
