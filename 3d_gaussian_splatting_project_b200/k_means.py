"""Drop-in for the reference's `3D_clustering/k_means.py`: same functions, prints and CLI, with
the per-point KDTree loop and the per-cluster boolean-mask means replaced by GPU kernels.

    get_vertex_info, k_means_kd_tree, k_means_with_color, add_label_proberty, COLORS

Semantics kept from the reference (line numbers are the reference's):
  * init draws `np.random.choice(N, k, replace=False)` from the GLOBAL NumPy stream (:63, :111),
    so `np.random.seed(s)` before the call gives the reference's start;
  * every iteration prints its index and the centroid shift (:66/:83, :114/:131);
  * on convergence it prints "Converged after i iterations." and breaks BEFORE adopting the new
    centroids (:84-88, :132-136), so the final assignment uses the pre-update centroids;
  * an empty cluster keeps its old centroid (:78, :126);
  * labels come back as int64, centroids float32, `colors` recoloured in place with
    COLORS[c % 8] (divided by 255.0 only in k_means_with_color, :100 vs :148).

`update=` selects how the mean is formed:
  "ordered" (default on one GPU)  float32 sequential sum in index order -- NumPy's arithmetic,
            bit-identical centroids, hence identical trajectories;
  "fast"    float64 segmented sums, shardable; centroids agree with the reference to its own
            float32 rounding (about 1e-5 relative on clusters of 1e5 members).
Under torch.distributed (world size > 1) each rank passes ITS rows and `update` is forced to
"fast"; the partial sums and counts are exchanged inside the reduction kernel over peer memory
(ops.KMeansExchange -> gsl_kmeans_step_exchange).  GSLIFT_KMEANS_EXCHANGE=nccl selects the
three-step form instead (reduce kernel, NCCL all-reduce of K x (D+1) float64, finalize kernel),
which is also what is used when torch cannot set up symmetric memory on the machine.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import ops, plyio

COLORS = [[252, 199, 55], [242, 107, 15], [231, 56, 121], [126, 24, 145],
          [247, 44, 91], [255, 116, 139], [167, 212, 119], [228, 241, 172]]


def get_vertex_info(plydata):
    """(points [N,3], colors [N,3]) from x/y/z and f_dc_0..2, printing both (reference :10-31)."""
    vertex = plydata["vertex"]
    points = np.column_stack((vertex["x"], vertex["y"], vertex["z"]))
    colors = np.column_stack((vertex["f_dc_0"], vertex["f_dc_1"], vertex["f_dc_2"]))
    print(points)
    print(colors)
    return points, colors


def _dist_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


_exchanges = {}


def _make_exchange(centroids, dist):
    """The fused exchange for this (K, D, device), or None when the three-step NCCL form is asked
    for (GSLIFT_KMEANS_EXCHANGE=nccl) or symmetric memory cannot be set up.  Cached: the buffers are
    a collective allocation."""
    if os.environ.get("GSLIFT_KMEANS_EXCHANGE", "peer") == "nccl":
        return None
    key = (centroids.shape[0], centroids.shape[1], str(centroids.device), dist is not None)
    if key not in _exchanges:
        try:
            _exchanges[key] = ops.KMeansExchange(centroids.shape[0], centroids.shape[1], centroids.device)
        except Exception as exc:       # symmetric memory unavailable (no peer access, old torch): NCCL
            import sys
            if dist is None:
                raise
            print(f"gslift: peer-memory exchange unavailable ({exc}); using NCCL all-reduce", file=sys.stderr)
            _exchanges[key] = None
    return _exchanges[key]


def lloyd(data, centroids, max_iter=100, tol=1e-4, update=None, verbose=True, device=None):
    """The iteration shared by both reference entry points, on device tensors.

    data float32 [N,D] (this rank's rows), centroids float32 [K,D] (replicated).
    Returns (centroids, labels int32 device tensor, iterations run)."""
    dist = _dist_world()
    if update is None:
        update = os.environ.get("GSLIFT_KMEANS_UPDATE", "ordered")
    if dist is not None:
        update = "fast"
    if update not in ("ordered", "fast"):
        raise ValueError(f"update must be 'ordered' or 'fast', not {update!r}")
    labels = torch.empty(data.shape[0], dtype=torch.int32, device=data.device)
    sums = torch.empty((centroids.shape[0], centroids.shape[1] + 1), dtype=torch.float64, device=data.device)
    exchange = _make_exchange(centroids, dist) if update == "fast" else None
    done = 0
    for iteration in range(max_iter):
        if verbose:
            print(iteration)
        if exchange is not None:
            new_centroids, shift = exchange.step(data, centroids, labels)
        elif update == "fast":
            ops.kmeans_step(data, centroids, labels, sums)
            if dist is not None:
                dist.all_reduce(sums)                       # K x (D+1) float64 over NCCL
            new_centroids, shift = ops.kmeans_finalize(sums, centroids)
        else:
            ops.kmeans_assign(data, centroids, labels)
            new_centroids, shift = ops.kmeans_update_ordered(data, labels, centroids)
        shift_value = np.float32(shift.item())
        if exchange is not None and exchange.world > 1 and np.isnan(shift_value):
            raise RuntimeError("K-means exchange returned a NaN shift: a rank did not reach the exchange within "
                               "2 s (every rank must call lloyd with the same max_iter/tol), or the data is not finite")
        done = iteration + 1
        if verbose:
            print(shift_value)
        if shift_value < tol:
            if verbose:
                print(f"Converged after {iteration + 1} iterations.")
            break
        centroids = new_centroids
    ops.kmeans_assign(data, centroids, labels)
    return centroids, labels, done


def _run(data_np, k, colors, max_iter, tol, palette_scale, update):
    data_np = np.ascontiguousarray(data_np, dtype=np.float32)
    N, _ = data_np.shape
    start = data_np[np.random.choice(N, k, replace=False)]
    device = torch.device("cuda")
    data = torch.from_numpy(data_np).to(device)
    centroids, labels, _ = lloyd(data, torch.from_numpy(start).to(device), max_iter, tol, update)
    palette = (np.array(COLORS) / 255.0 if palette_scale else np.array(COLORS)).astype(np.float32)
    if colors.dtype == np.float32:
        colors_dev = torch.from_numpy(np.ascontiguousarray(colors)).to(device)
        ops.recolor(labels, torch.from_numpy(palette).to(device), colors_dev)
    else:   # the reference assigns the float64 palette into whatever dtype `colors` has
        wide = np.array(COLORS) / 255.0 if palette_scale else np.array(COLORS)
        pal = torch.from_numpy(wide.astype(colors.dtype)).to(device)
        colors_dev = pal[labels.long() % len(COLORS)]
    colors[...] = colors_dev.cpu().numpy()
    return centroids.cpu().numpy(), labels.cpu().numpy().astype(np.int64), colors


def k_means_kd_tree(data, k, colors, max_iter=100, tol=1e-4, update=None):
    """K-means on an arbitrary feature block [N,D] (reference :46-103)."""
    return _run(data, k, colors, max_iter, tol, False, update)


def k_means_with_color(points, k, colors, max_iter=100, tol=1e-4, update=None):
    """K-means on concat(points, colors) (reference :107-151)."""
    data = np.concatenate((points, colors), axis=1)
    return _run(data, k, colors, max_iter, tol, True, update)


def add_label_proberty(ply_data, output_ply, label):
    """Append an int32 `label` property and write an ASCII PLY (reference :169-194)."""
    vertex = plyio.describe_with_label(ply_data["vertex"].data, label)
    plyio.write_ply(output_ply, [("vertex", vertex)], text=True)
    print(f"New PLY file with label added saved to {output_ply}")


def main(argv=None):
    parser = argparse.ArgumentParser(description="K-means clustering on a point cloud.")
    parser.add_argument("--file_path", type=str, required=True, help="Path to the input PLY file.")
    parser.add_argument("--save_path", type=str, required=True, help="Path to save the modified PLY file.")
    parser.add_argument("--k", type=int, default=10, help="Number of clusters for k-means.")
    args = parser.parse_args(argv)

    with open(args.file_path, "rb") as fh:
        plydata = plyio.read_ply(fh)
    points, colors = get_vertex_info(plydata)
    _, labels, colors = k_means_with_color(points, args.k, colors, max_iter=10)
    print(labels)
    add_label_proberty(plydata, args.save_path, labels)


if __name__ == "__main__":
    main()
