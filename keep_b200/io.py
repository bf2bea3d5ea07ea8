"""Slide feature files -> device tensors (the input side of the WSI scripts).

Reference: `WSI_evaluation/zeroshot_detection_WSI.py:29-31` (and the other two scripts) read a CLAM-style `.h5` with
datasets `features [N,768] float32` and `coords [N,2] int`; `WSI_evaluation/utils.py:50-60` reads the same `.h5` or a
`.pt` tensor of features. `load_slide` accepts those two formats plus `.npz`/`.npy` (h5py is an optional import: the
reference pins it, this image does not ship it), and `to_device` uploads through pinned staging buffers in chunks so
that the host->device copy of one chunk overlaps the page-locking memcpy of the next.
"""
from __future__ import annotations

import os

import numpy as np
import torch


def load_slide(path: str):
    """-> (features float32 [N,D] CPU tensor, coords int64 [N,2] CPU tensor or None)."""
    ext = os.path.splitext(path)[1].lower()
    if ext in (".h5", ".hdf5"):
        try:
            import h5py
        except ImportError as e:  # loud, not a silent fallback
            raise ImportError("reading .h5 feature files needs h5py (WSI_evaluation/zeroshot_detection_WSI.py:3)") from e
        with h5py.File(path, "r") as f:  # zeroshot_detection_WSI.py:29-31
            feats = torch.from_numpy(f["features"][:])
            coords = torch.from_numpy(np.asarray(f["coords"][:])) if "coords" in f else None
    elif ext == ".pt":  # utils.py:58-60: a bare feature tensor; a dict with 'features'/'coords' is accepted as well
        obj = torch.load(path, map_location="cpu")
        if isinstance(obj, dict):
            feats, coords = obj["features"], obj.get("coords")
        else:
            feats, coords = obj, None
    elif ext == ".npz":
        z = np.load(path)
        feats = torch.from_numpy(z["features"])
        coords = torch.from_numpy(z["coords"]) if "coords" in z.files else None
    elif ext == ".npy":
        feats, coords = torch.from_numpy(np.load(path)), None
    else:
        raise ValueError(f"unknown slide feature format {ext!r} ({path})")
    if feats.dim() != 2:
        raise ValueError(f"features must be [N,D], got {tuple(feats.shape)}")
    feats = feats.to(torch.float32).contiguous()
    if coords is not None:
        coords = torch.as_tensor(coords).to(torch.int64).reshape(-1, 2).contiguous()
        if coords.shape[0] != feats.shape[0]:
            raise ValueError(f"{coords.shape[0]} coords for {feats.shape[0]} feature rows")
    return feats, coords


def to_device(t: torch.Tensor, device, chunk_rows: int = 16384) -> torch.Tensor:
    """Chunked upload through two pinned staging buffers (double-buffered on a copy stream); returns a device tensor
    that is ready on the CURRENT stream."""
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("to_device uploads to a CUDA device")
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=t.dtype, device=device)
    n = t.shape[0]
    if n == 0:
        return out
    if t.is_pinned():
        out.copy_(t, non_blocking=True)
        return out
    rows = max(1, min(n, chunk_rows))
    stage = [torch.empty((rows,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True) for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    # `out` was allocated on the current stream: the caching allocator may have handed back a block that kernels already
    # queued on that stream still read (the previous slide's features), so the copies must be ordered after them
    copy_stream.wait_stream(torch.cuda.current_stream(device))
    out.record_stream(copy_stream)
    for i, r0 in enumerate(range(0, n, rows)):
        k = min(rows, n - r0)
        s = i & 1
        if i >= 2:
            done[s].synchronize()            # the staging buffer has left the host
        stage[s][:k].copy_(t[r0:r0 + k])      # pageable -> pinned (CPU memcpy), overlaps the previous chunk's DMA
        with torch.cuda.stream(copy_stream):
            out[r0:r0 + k].copy_(stage[s][:k], non_blocking=True)
            done[s].record(copy_stream)
    torch.cuda.current_stream(device).wait_stream(copy_stream)
    for e in done:
        e.synchronize()                       # the staging buffers are about to be freed
    return out
