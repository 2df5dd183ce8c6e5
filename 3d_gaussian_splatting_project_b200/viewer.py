"""Viewer-side consumers of the lifted labels on the GPU.

Mirrors the two per-Gaussian loops of the reference viewer's worker
(Web_Viewer_Gaussians_Selection/gaussians_selection.js, `gs`), names in snake case:

    multiply4(a, b)                  gs:110-123
    run_sort(positions, view_proj)   gs:417-462  runSort -> depthIndex (Uint32Array)
    perform_hit_testing(...)         gs:361-395  performHitTesting -> selected label

Both run through libgslift.so (gsl_viewer_depth_sort / gsl_viewer_hit_test, csrc/viewer.cu) in
float64 with JavaScript's evaluation order; there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._native import check, lib

NO_SELECTION = -999999      # gs:6
ROW_FLOATS = 8              # the viewer's 32-byte rows: 3 position, 3 scale floats, rgba, rot (gs:237)


def multiply4(a, b) -> list:
    """multiply4 (gs:110-123): column-major 4x4 product as the viewer forms projection x view.
    Plain Python floats (binary64, like JavaScript numbers), same association order."""
    a = [float(v) for v in a]
    b = [float(v) for v in b]
    if len(a) != 16 or len(b) != 16:
        raise ValueError("multiply4 takes two 16-element matrices")

    def row_by_col(row, col):
        return b[row] * a[col] + b[row + 1] * a[col + 4] + b[row + 2] * a[col + 8] + b[row + 3] * a[col + 12]

    return [row_by_col(r, c) for r in (0, 4, 8, 12) for c in range(4)]


def _rows(positions) -> torch.Tensor:
    """Accepts the viewer's float rows [N][8] or plain positions [N][3] (any stride >= 3), float32,
    on a CUDA device (host arrays are uploaded to the current device)."""
    t = positions if isinstance(positions, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(positions, np.float32))
    if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] < 3:
        raise TypeError("positions must be float32 [N][stride >= 3]")
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("viewer operators need a CUDA device (there is no CPU path)")
        t = t.cuda()
    return t.contiguous()


def run_sort(positions, view_proj, out: torch.Tensor | None = None) -> torch.Tensor:
    """runSort (gs:427-457): depthIndex, the Gaussian indices in a stable order of increasing 16-bit
    depth bucket, as a device uint32 tensor (Uint32Array, gs:453).  The early-out of gs:421-425 (camera barely moved) is the
    caller's decision, as in the viewer's throttled loop (gs:588-600)."""
    pos = _rows(positions)
    vp = np.ascontiguousarray(np.asarray(view_proj, np.float64).reshape(16))
    N, stride = pos.shape
    if out is None:
        out = torch.empty(N, dtype=torch.int32, device=pos.device)
    if out.numel() != N or out.element_size() != 4 or not out.is_cuda:
        raise ValueError("out must be a 4-byte device tensor of N elements")
    L = lib()
    ws = ops._ws.get(pos.device, L.gsl_viewer_sort_workspace_bytes(N))
    with torch.cuda.device(pos.device):
        check(L.gsl_viewer_depth_sort(pos.data_ptr(), N, stride, vp.ctypes.data, out.data_ptr(),
                                      ws.data_ptr(), ws.numel(), ops._stream()))
    return out.view(torch.uint32) if out.dtype != torch.uint32 else out


def perform_hit_testing(x, y, view_matrix, projection_matrix, viewport, positions, label_data,
                        return_index: bool = False):
    """performHitTesting (gs:361-395).  `positions` / `label_data` stand for the worker's `buffer` and
    `labelData` globals.  Returns the selected label (NO_SELECTION when nothing lies within 10 px),
    or (label, index) with return_index."""
    pos = _rows(positions)
    labels = label_data if isinstance(label_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(label_data, np.int32))
    if labels.dtype != torch.int32:
        raise TypeError("label_data must be int32 (Int32Array, gs:298)")
    labels = labels.to(pos.device).contiguous()
    N, stride = pos.shape
    if labels.numel() != N:
        raise ValueError("label_data must hold one label per Gaussian")
    m = np.array(multiply4(projection_matrix, view_matrix), np.float64)          # gs:364
    L = lib()
    ws = ops._ws.get(pos.device, L.gsl_viewer_hit_workspace_bytes())
    res = torch.empty(4, dtype=torch.int32, device=pos.device)                   # [label, pad, index lo, index hi]
    with torch.cuda.device(pos.device):
        check(L.gsl_viewer_hit_test(pos.data_ptr(), labels.data_ptr(), N, stride, m.ctypes.data, float(x), float(y),
                                    float(viewport[0]), float(viewport[1]), NO_SELECTION, res.data_ptr(),
                                    res.data_ptr() + 8, ws.data_ptr(), ws.numel(), ops._stream()))
    host = res.cpu().numpy()
    label, index = int(host[0]), int(host[2:4].view(np.int64)[0])
    return (label, index) if return_index else label
