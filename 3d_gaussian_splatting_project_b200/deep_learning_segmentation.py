"""Drop-in for the reference's `deep_learning_segmentation.py` with the N x V vote loop moved to
the GPU library.  Same public names, signatures, prints, CLI flags and output file:

    load_cameras, load_gaussians, project_gaussian, assign_labels, save_labeled_ply, main

What changed: `assign_labels` no longer walks Gaussians in Python (reference :274-306).  It
still visits the cameras in order, still skips a camera whose `<input_dir>/<img_name>.png` is
missing with the same warning (:257-259), still asks the 2-D segmenter for one map per view
(:266) -- but each map is uploaded and packed as it arrives and one `lift_votes` call
(include/gslift.h) produces the labels.

The 2-D segmentation stage itself (SegFormer / Mask2Former / YOLO inference, reference
:85-238) is upstream of this path and is not re-implemented.  `assign_labels` finds a
segmenter in this order:
  1. the `segmenter=` argument: callable(image_path, output_dir, model_type) -> int array [H, W]
  2. a precomputed `<output_dir>/<img_name>_segmap.npy`, the file the reference's
     `segment_image` writes for every view (:165)
  3. the reference's own `initialize_model` / `segment_image`, imported from the checkout
     named by $GSLIFT_REFERENCE_DIR
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os

import numpy as np
import torch
from PIL import Image

from . import ops, plyio

GAUSSIAN_DTYPE = np.dtype([("position", np.float32, 3), ("scale", np.float32, 3), ("rotation", np.float32, 4)])


def load_cameras(camera_file):
    """cameras.json -> list of dicts (reference :17-22)."""
    with open(camera_file, "r") as fh:
        return json.load(fh)


def load_gaussians(ply_file):
    """PLY -> (structured array with `position` filled from x/y/z, parsed PLY) (reference :25-40).
    `scale` and `rotation` stay zero, as in the reference."""
    ply = plyio.read_ply(ply_file)
    vertex = ply["vertex"]
    gaussians = np.zeros(len(vertex), dtype=GAUSSIAN_DTYPE)
    for axis, name in enumerate(("x", "y", "z")):
        gaussians["position"][:, axis] = vertex[name]
    return gaussians, ply


def project_gaussian(position, camera):
    """Scalar projection of one Gaussian mean (reference :43-82): (int x, int y) or None.
    Kept for API parity; the fast path evaluates the same float64 expressions on the GPU."""
    R = np.array(camera["rotation"])
    cam = R @ position + (-R @ np.array(camera["position"]))
    if cam[2] <= 0:
        return None
    w, h = camera["width"], camera["height"]
    u = (camera["fx"] * cam[0] / cam[2]) + w / 2
    v = (camera["fy"] * cam[1] / cam[2]) + h / 2
    if 0 <= u < w and 0 <= v < h:
        return (int(u), int(v))
    return None


# ----------------------------------------------------------------------------------------
# segmenter plumbing (upstream stage, not part of the hot path)
# ----------------------------------------------------------------------------------------
_ref_module = None


def _reference_module():
    """Import the reference's deep_learning_segmentation.py by path (for its segmenter)."""
    global _ref_module
    if _ref_module is None:
        root = os.environ.get("GSLIFT_REFERENCE_DIR")
        path = os.path.join(root, "deep_learning_segmentation.py") if root else None
        if not path or not os.path.exists(path):
            raise RuntimeError(
                "no segmenter: pass segmenter=..., provide <output_dir>/<img_name>_segmap.npy, or set "
                "GSLIFT_REFERENCE_DIR to a checkout of the reference (its segment_image is used unchanged)")
        spec = importlib.util.spec_from_file_location("_gslift_reference_dls", path)
        _ref_module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ref_module)
    return _ref_module


class _UpstreamSegmenter:
    """Lazily builds the reference's model once and calls its segment_image per view."""

    def __init__(self, model_type):
        self.model_type = model_type
        self._state = None

    def __call__(self, image_path, output_dir, model_type):
        ref = _reference_module()
        if self._state is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
            processor, model = ref.initialize_model(self.model_type, device)
            model.to(device)
            self._state = (processor, model, device)
        processor, model, device = self._state
        return ref.segment_image(image_path, output_dir, processor, model, device, model_type)


def _precomputed_map(output_dir, img_name):
    path = os.path.join(output_dir, f"{img_name}_segmap.npy") if output_dir else None
    return np.load(path) if path and os.path.exists(path) else None


# ----------------------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------------------
def _stage_maps(seg_maps, shapes, device, label_min, n_classes):
    """Host int maps -> packed uint8 codes on the device (N2 of SURVEY 8f: label-map staging).

    One process: upload every map (int32), device min/max picks a tight code range, pack.
    Under torch.distributed every rank holds the same host maps but uploads and packs only its
    contiguous block of views; the packed blocks (1 byte per pixel) are then broadcast rank by
    rank over NCCL, so the PCIe cost per rank drops by the world size."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    sizes = [h * w for h, w in shapes]
    starts = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)     # pixels, row-major staging
    pstarts = ops.packed_offsets(shapes)                                   # bytes, packed layout
    total = int(pstarts[-1])
    V = len(seg_maps)
    per, extra = divmod(V, world)
    cuts = [r * per + min(r, extra) for r in range(world + 1)]
    v_lo, v_hi = cuts[rank], cuts[rank + 1]
    mine = int(starts[v_hi] - starts[v_lo])
    staged = torch.empty(mine, dtype=torch.int32, device=device)
    off = 0
    for v in range(v_lo, v_hi):
        m = seg_maps[v]
        src = m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m, np.int32))
        staged[off:off + sizes[v]].copy_(src.reshape(-1), non_blocking=True)
        off += sizes[v]
    if label_min is None or n_classes is None:
        # tight code range from the data (device min/max): fewer histogram rows per Gaussian
        lo, hi = ops.label_range(staged) if mine else (2**31 - 1, -2**31)
        if world > 1:
            mm = torch.tensor([lo, -hi], dtype=torch.int64, device=device)
            dist.all_reduce(mm, op=dist.ReduceOp.MIN)
            lo, hi = int(mm[0].item()), -int(mm[1].item())
        if int(starts[-1]) == 0:
            lo, hi = ops.DEFAULT_LABEL_MIN, ops.DEFAULT_LABEL_MIN
        if hi - lo + 1 > ops.DEFAULT_N_CLASSES:
            raise ValueError(f"label maps span {hi - lo + 1} values; at most {ops.DEFAULT_N_CLASSES} are supported")
        label_min, n_classes = lo, hi - lo + 1
    packed = torch.empty(total, dtype=torch.uint8, device=device)
    if mine:
        ops.pack_labels(staged, shapes[v_lo:v_hi], label_min, n_classes, out=packed[int(pstarts[v_lo]):int(pstarts[v_hi])])
    if world > 1:
        for r in range(world):
            lo_px, hi_px = int(pstarts[cuts[r]]), int(pstarts[cuts[r + 1]])
            if hi_px > lo_px:
                dist.broadcast(packed[lo_px:hi_px], src=r)
    return packed, label_min, n_classes


_copy_streams = {}
last_call_stats = {}               # bytes lift_labels moved over PCIe in its last call (this process)
_pinned_codes = {}                 # pinned uint8 staging for host-narrowed views, by size
_host_stage = {"px_per_s": None, "fixed_s": 0.008}   # calibrated by use: host narrowing rate (pixels per second, all
                                                      # threads, with the DMA engine running) and the other host time of a call
_PCIE_BYTES_PER_S = 54e9           # pinned host -> device copy rate of a PCIe 5 x16 link (profiles/probes/h2d_probe.py)


def _host_cores():
    """Host cores this process may run on (affinity / cgroup aware where the platform tells)."""
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def _host_fraction(seg_maps, n_px):
    """Share of the views whose maps are narrowed to 1-byte codes on the host before they cross
    the bus.  The DMA engine moves 4 bytes per pixel for the others meanwhile; with P pixels in all,
    tp = bus time per byte, th = host time per pixel and c = the other host work of the call
    (enqueueing, the view table), both sides finish together when
        c + f P th = P tp (4 - 3 f)   =>   f = (4 P tp - c) / (P (th + 3 tp)).
    GSLIFT_HOST_STAGE=<fraction> fixes it (0 = all maps cross as int32)."""
    if any(isinstance(m, torch.Tensor) and m.is_cuda for m in seg_maps):
        return 0.0
    env = os.environ.get("GSLIFT_HOST_STAGE")
    if env is not None:
        return min(max(float(env), 0.0), 1.0)
    cores = _host_cores()
    if cores < 8:
        return 0.0
    rate = _host_stage["px_per_s"] or 0.9e9 * cores
    tp, th, c = 1.0 / _PCIE_BYTES_PER_S, 1.0 / rate, _host_stage["fixed_s"]
    if n_px <= 0:
        return 0.0
    return min(max((4 * n_px * tp - c) / (n_px * (th + 3 * tp)), 0.0), 0.85)


def _host_ptr(m):
    """(address, keep-alive) of a host int32 map, converting only if it is not int32 / contiguous."""
    if isinstance(m, torch.Tensor):
        t = m if (m.dtype == torch.int32 and m.is_contiguous()) else m.to(torch.int32).contiguous()
        return t.data_ptr(), t
    arr = m if (isinstance(m, np.ndarray) and m.dtype == np.int32 and m.flags.c_contiguous) else np.ascontiguousarray(m, np.int32)
    return arr.ctypes.data, arr


def _lift_pipelined(positions, cameras, image_sizes, seg_maps, shapes, device):
    """One-process fast path of lift_labels from HOST maps, organised around the PCIe transfer,
    which is what such a call waits for:

      * the first windows of 16 views cross the bus as int32, two staging buffers in flight on a
        copy stream, and are packed on the device (gsl_pack_labels);
      * the remaining windows are narrowed to 1-byte codes by the host cores meanwhile
        (gsl_host_pack_labels), cross as uint8 behind the int32 windows, and are tiled on the
        device (gsl_tile_codes);
      * every window is swept as soon as its packed maps are resident (gsl_lift_prepare once, then
        gsl_lift_gather_range per window / batch), so only the last sweep and the majority are not
        hidden behind the transfer.

    The first uploads are enqueued before anything else happens on the host.  Codes are label + 2
    (label_min = -1); returns None when the labels do not fit that window and the caller must take
    the general path."""
    import ctypes
    import time
    from ._native import check, lib
    L = lib()
    V = len(seg_maps)
    sizes = [h * w for h, w in shapes]
    starts = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)     # pixels, row-major staging
    pstarts = ops.packed_offsets(shapes)                                   # bytes, packed layout
    main = torch.cuda.current_stream(device)
    key = (device.type, device.index)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device)
    copy = _copy_streams[key]
    CH = 16
    chunks = list(range(0, V, CH))
    n_host = int(round(_host_fraction(seg_maps, int(starts[-1])) * len(chunks)))
    n_dev = len(chunks) - n_host                                 # windows [0, n_dev) cross as int32
    chunk_px = max([int(starts[min(v0 + CH, V)] - starts[v0]) for v0 in chunks[:n_dev]] + [1])
    slots = [torch.empty(chunk_px, dtype=torch.int32, device=device) for _ in range(2)]
    slot_free = [None, None]
    ready = [None] * n_dev

    def upload(ci):
        v0 = chunks[ci]
        v1 = min(v0 + CH, V)
        buf = slots[ci % 2]
        with torch.cuda.stream(copy):
            if slot_free[ci % 2] is not None:
                copy.wait_event(slot_free[ci % 2])
            off, v = 0, v0
            while v < v1:
                m = seg_maps[v]
                src = m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m, np.int32))
                n, last = sizes[v], v
                # maps that lie back to back in one host allocation (slices of one big pinned tensor)
                # cross in a single copy
                if src.dtype == torch.int32 and src.is_contiguous() and not src.is_cuda:
                    while last + 1 < v1:
                        nxt = seg_maps[last + 1]
                        if not (isinstance(nxt, torch.Tensor) and nxt.dtype == torch.int32 and nxt.is_contiguous() and not nxt.is_cuda
                                and nxt.untyped_storage().data_ptr() == src.untyped_storage().data_ptr()
                                and nxt.storage_offset() == src.storage_offset() + n):
                            break
                        last += 1
                        n += sizes[last]
                    flat = torch.as_strided(src, (n,), (1,), src.storage_offset())
                else:
                    flat = src.reshape(-1)
                buf[off:off + n].copy_(flat, non_blocking=True)
                off += n
                v = last + 1
            ready[ci] = torch.cuda.Event()
            ready[ci].record(copy)
        return off

    def pack_runs(fn, src_ptr, elem, v0, v1):
        """fn = gsl_pack_labels / gsl_tile_codes over the runs of equal shapes of views [v0, v1)."""
        v = v0
        while v < v1:
            n = 1
            while v + n < v1 and shapes[v + n] == shapes[v]:
                n += 1
            yield fn, src_ptr + elem * int(starts[v] - starts[v0]), n, shapes[v][1], shapes[v][0], int(pstarts[v])
            v += n

    trace = os.environ.get("GSLIFT_TRACE") == "1"               # phase timings of this call on stderr
    t_call0 = time.perf_counter()
    with torch.cuda.device(device):
        if trace:
            import sys
            t_host0 = time.perf_counter()
            print(f"[gslift trace] pipeline starts at {t_host0:.6f}; {n_dev} windows as int32, {n_host} narrowed on the host", file=sys.stderr)
            ev0 = torch.cuda.Event(enable_timing=True); ev0.record(main)
        n_px = [0] * n_dev
        for ci in range(min(2, n_dev)):                          # the transfer starts now
            n_px[ci] = upload(ci)
        pos = positions if isinstance(positions, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(positions, np.float32))
        pos = pos.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        views = ops.make_views(cameras, shapes, image_sizes)
        N = pos.shape[0]
        packed = torch.empty(int(pstarts[-1]), dtype=torch.uint8, device=device)
        ws = ops._ws.get(device, L.gsl_lift_workspace_bytes(N, V))
        minmax = torch.tensor([2**31 - 1, -2**31], dtype=torch.int32, device=device)
        err = torch.zeros(1, dtype=torch.int32, device=device)
        vptr = views.ctypes.data
        check(L.gsl_lift_prepare(pos.data_ptr(), N, vptr, V, ws.data_ptr(), ws.numel(), main.cuda_stream))
        # ---- windows that cross as int32: everything below is enqueued without waiting
        for ci in range(n_dev):
            v0 = chunks[ci]
            v1 = min(v0 + CH, V)
            buf = slots[ci % 2]
            main.wait_event(ready[ci])
            check(L.gsl_label_range(buf.data_ptr(), n_px[ci], minmax.data_ptr(), main.cuda_stream))
            for _, src, n, w, h, dst in pack_runs(None, buf.data_ptr(), 4, v0, v1):
                check(L.gsl_pack_labels(src, n, w, h, packed.data_ptr() + dst, -1, 255, err.data_ptr(), main.cuda_stream))
            slot_free[ci % 2] = torch.cuda.Event()
            slot_free[ci % 2].record(main)
            if ci + 2 < n_dev:
                n_px[ci + 2] = upload(ci + 2)            # refills the slot just packed
            check(L.gsl_lift_gather_range(pos.data_ptr(), N, vptr, V, v0, v1, packed.data_ptr(), None, 0.0, 0,
                                          ws.data_ptr(), ws.numel(), main.cuda_stream))
        # ---- windows narrowed on the host, a few at a time, while the DMA engine is busy with the above
        host_mm = (ctypes.c_int * 2)(2**31 - 1, -2**31)
        if n_host:
            hv0 = chunks[n_dev]
            host_px = int(starts[V] - starts[hv0])
            if _pinned_codes.get("n", 0) < host_px:
                t_alloc = time.perf_counter()
                _pinned_codes["buf"] = torch.empty(host_px, dtype=torch.uint8, pin_memory=True)
                _pinned_codes["n"] = host_px
                t_call0 += time.perf_counter() - t_alloc         # a one-time cost, not part of the calibration
            pinned = _pinned_codes["buf"]
            codes = torch.empty(host_px, dtype=torch.uint8, device=device)
            bad = ctypes.c_int(0)
            # host batches of two windows; the last two windows go one at a time, so that little is
            # left to upload and sweep once the host is done
            batches, b0 = [], n_dev
            while b0 < len(chunks):
                nb = 2 if len(chunks) - b0 > 3 else 1
                batches.append((b0, min(b0 + nb, len(chunks))))
                b0 += nb
            t_narrow, px_narrow = 0.0, 0
            for b0, b1 in batches:
                vb0 = chunks[b0]
                vb1 = min(chunks[b1 - 1] + CH, V)
                keep = [_host_ptr(seg_maps[v]) for v in range(vb0, vb1)]
                ptrs = (ctypes.c_void_p * len(keep))(*[k[0] for k in keep])
                npx = (ctypes.c_int64 * len(keep))(*[sizes[v] for v in range(vb0, vb1)])
                off0, off1 = int(starts[vb0] - starts[hv0]), int(starts[vb1] - starts[hv0])
                t0 = time.perf_counter()
                check(L.gsl_host_pack_labels(ptrs, npx, len(keep), -1, 255, pinned.data_ptr() + off0, _host_cores(), host_mm, ctypes.byref(bad)))
                t_narrow += time.perf_counter() - t0
                px_narrow += off1 - off0
                with torch.cuda.stream(copy):
                    codes[off0:off1].copy_(pinned[off0:off1], non_blocking=True)
                    landed = torch.cuda.Event()
                    landed.record(copy)
                main.wait_event(landed)
                for _, src, n, w, h, dst in pack_runs(None, codes.data_ptr() + off0, 1, vb0, vb1):
                    check(L.gsl_tile_codes(src, n, w, h, packed.data_ptr() + dst, main.cuda_stream))
                check(L.gsl_lift_gather_range(pos.data_ptr(), N, vptr, V, vb0, vb1, packed.data_ptr(), None, 0.0, 0,
                                              ws.data_ptr(), ws.numel(), main.cuda_stream))
            if t_narrow > 0:
                _host_stage["px_per_s"] = px_narrow / t_narrow
                # the other host work of this call; first calls also pay allocations, hence the cap
                _host_stage["fixed_s"] = min(max(time.perf_counter() - t_call0 - t_narrow, 0.0), 0.015)
        dev_px = int(starts[chunks[n_dev]] if n_dev < len(chunks) else starts[V])
        last_call_stats.update(h2d_bytes=4 * dev_px + int(starts[V] - dev_px) + int(pos.numel()) * 4, d2h_bytes=N * 4,
                               views_as_int32=min(n_dev * CH, V), views_narrowed_on_host=V - min(n_dev * CH, V))
        if trace:
            t_host1 = time.perf_counter()
            ev_up = torch.cuda.Event(enable_timing=True); ev_up.record(copy)
            ev_sw = torch.cuda.Event(enable_timing=True); ev_sw.record(main)
        lo, hi = (int(x) for x in minmax.tolist())           # synchronises: every copy has been consumed
        lo, hi = min(lo, int(host_mm[0])), max(hi, int(host_mm[1]))
        if lo < -1 or hi > 253:
            return None
        labels = torch.empty(N, dtype=torch.int32, device=device)
        check(L.gsl_lift_majority(N, V, -1, max(hi + 2, 1), labels.data_ptr(), ws.data_ptr(), ws.numel(), main.cuda_stream))
        if trace:
            ev_mj = torch.cuda.Event(enable_timing=True); ev_mj.record(main)
            ev_mj.synchronize()
            rate = _host_stage["px_per_s"]
            print(f"[gslift trace] majority done at {time.perf_counter():.6f} | host enqueue + narrowing {1e3 * (t_host1 - t_host0):.2f} ms"
                  f" ({'%.1f' % (rate / 1e9) if rate else '-'} Gpx/s) | uploads done +{ev0.elapsed_time(ev_up):.2f} ms | "
                  f"last sweep done +{ev0.elapsed_time(ev_sw):.2f} ms | majority done +{ev0.elapsed_time(ev_mj):.2f} ms", file=sys.stderr)
    return labels


def lift_labels(positions, cameras, seg_maps, image_sizes=None, device=None, want_near=False,
                near_eps=1e-4, label_min=None, n_classes=None):
    """Majority-vote labels for `positions` given one segmentation map per camera.

    positions   float32 [N,3] array or tensor (under torch.distributed: THIS rank's Gaussians)
    cameras     camera dicts, in voting order
    seg_maps    list of int arrays [seg_h, seg_w] (NumPy or tensors), one per camera
    image_sizes per camera (orig_w, orig_h); default = the map's own size (scale 1.0)
    Returns int32 NumPy labels (and the near-boundary mask when want_near).
    """
    device = torch.device(device if device is not None else "cuda")
    if os.environ.get("GSLIFT_TRACE") == "1":
        import sys, time
        print(f"[gslift trace] lift_labels entered at {time.perf_counter():.6f}", file=sys.stderr)
    shapes = [tuple(m.shape) for m in seg_maps]
    n_pos = positions.shape[0]
    import torch.distributed as dist
    single = not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    if (single and not want_near and label_min is None and n_classes is None and len(cameras) and n_pos
            and all(len(sh) == 2 for sh in shapes)):
        fast = _lift_pipelined(positions, cameras, image_sizes, seg_maps, shapes, device)
        if fast is not None:
            return _to_host(fast)
    pos = positions if isinstance(positions, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(positions, np.float32))
    pos = pos.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    views = ops.make_views(cameras, shapes, image_sizes)
    packed, label_min, n_classes = _stage_maps(seg_maps, shapes, device, label_min, n_classes)
    world = dist.get_world_size() if not single else 1
    last_call_stats.update(h2d_bytes=4 * sum(h * w for h, w in shapes) // world + int(pos.numel()) * 4, d2h_bytes=int(pos.shape[0]) * 4,
                           views_as_int32=len(shapes), views_narrowed_on_host=0)
    res = ops.lift_votes(pos, views, packed, label_min, n_classes, want_near=want_near, near_eps=near_eps)
    if want_near:
        return _to_host(res[0]), _to_host(res[1])
    return _to_host(res)


def _to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> NumPy through a pinned staging tensor (a pageable `.cpu()` of 24 MB costs
    ~10 ms; pinned, it is PCIe speed).  The array owns its pinned storage."""
    trace = os.environ.get("GSLIFT_TRACE") == "1"
    if trace:
        import sys, time
        t0 = time.perf_counter()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    if trace:
        t1 = time.perf_counter()
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    if trace:
        print(f"[gslift trace] result to host: pinned alloc {1e3 * (t1 - t0):.2f} ms, copy {1e3 * (time.perf_counter() - t1):.2f} ms", file=sys.stderr)
    return host.numpy()


def assign_labels(gaussians, cameras, input_dir, output_dir, model_type="mask2former", segmenter=None):
    """Label every Gaussian by majority vote over its projections (reference :241-308).

    Returns np.int32 [N]; -1 for Gaussians no view sees."""
    upstream = None
    used, maps, sizes = [], [], []
    for camera in cameras:
        img_path = os.path.join(input_dir, camera["img_name"] + ".png")
        if not os.path.exists(img_path):
            print(f"Warning: Image {camera['img_name']} not found")
            continue
        print(f"Processing image {os.path.basename(img_path)}...")
        with Image.open(img_path) as image:
            size = (image.size[0], image.size[1])
        if segmenter is not None:
            seg_map = segmenter(img_path, output_dir, model_type)
        else:
            seg_map = _precomputed_map(output_dir, camera["img_name"])
            if seg_map is None:
                if upstream is None:
                    upstream = _UpstreamSegmenter(model_type)
                seg_map = upstream(img_path, output_dir, model_type)
        used.append(camera)
        maps.append(np.asarray(seg_map))
        sizes.append(size)
    positions = np.ascontiguousarray(gaussians["position"], np.float32)
    if not used:
        return np.full(len(positions), -1, dtype=np.int32)
    return lift_labels(positions, used, maps, sizes)


def save_labeled_ply(output_file, plydata, labels):
    """Write the input vertices plus an int32 `label` column as binary PLY (reference :311-332)."""
    vertex = plyio.describe_with_label(plydata["vertex"].data, labels)
    plyio.write_ply(output_file, [("vertex", vertex)], text=False)


def main():
    parser = argparse.ArgumentParser(description="Add labels to gaussians PLY file")
    parser.add_argument("--ply_file", help="Input PLY file with gaussians data")
    parser.add_argument("--camera_file", help="JSON file with camera data")
    parser.add_argument("--input_dir", help="Directory containing input images")
    parser.add_argument("--output_dir", help="Output directory to saved segmented input images")
    parser.add_argument("--output_file", help="Output PLY file with labels")
    parser.add_argument("--model", choices=["segformer", "mask2former", "yolo"], default="mask2former",
                        help="Choose segmentation model: mask2former or yolo")
    args = parser.parse_args()

    print("Loading cameras...")
    cameras = load_cameras(args.camera_file)
    print("Loading gaussians...")
    gaussians, plydata = load_gaussians(args.ply_file)
    print("Assigning labels...")
    labels = assign_labels(gaussians, cameras, args.input_dir, args.output_dir, model_type=args.model)
    print("Saving labeled PLY file...")
    save_labeled_ply(args.output_file, plydata, labels)
    print(f"Done! Labeled PLY file saved as {args.output_file}")

    values, counts = np.unique(labels, return_counts=True)
    print("\nLabel statistics:")
    print(f"Total gaussians: {len(labels)}")
    print(f"Number of unique labels: {len(values)}")
    print("Label counts:")
    for value, count in zip(values, counts):
        print(f"Label {value}: {count} gaussians ({100 * count / len(labels):.2f}%)")


if __name__ == "__main__":
    main()
