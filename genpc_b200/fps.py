"""Farthest point sampling on the GPU -- the device-resident replacement for the reference's
`fpsample.fps_sampling(xyz_np, K)` CPU round trip (main.py:21-22, reg_xyz.py:215, DepthPrompting.py:88-91).

`fps_sampling(pc, n_samples, start_idx=None)` keeps fpsample's call shape.  Where fpsample picks a random
start when `start_idx` is None (so the reference output is not reproducible), this build pins start = 0
and breaks arg-max ties towards the lowest index.
"""
import numpy as np
import torch

from . import _lib


def furthest_point_sample(xyz, K, start=0, return_seq=False):
    """xyz [B,N,3] float32 CUDA tensor -> idx [B,K] int32 (and the selected-distance sequence)."""
    _lib.require_cuda(xyz)
    if xyz.dtype != torch.float32:
        raise _lib.GenpcError("xyz must be float32")
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    L = _lib.lib()
    idx = torch.empty(B, K, dtype=torch.int32, device=xyz.device)
    seq = torch.empty(B, K, dtype=torch.float32, device=xyz.device) if return_seq else None
    with torch.cuda.device(xyz.device):
        nbytes = L.genpc_fps_workspace_bytes(B, N, K)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=xyz.device)
        rc = L.genpc_fps(_lib.ptr(xyz), B, N, K, int(start), _lib.ptr(idx), _lib.ptr(seq), _lib.ptr(ws), nbytes,
                         _lib.current_stream(xyz.device))
    _lib.check(rc, "genpc_fps")
    return (idx, seq) if return_seq else idx


def fps_sampling(pc, n_samples, start_idx=None, device="cuda"):
    """fpsample-shaped entry point: pc [N,3] (numpy or tensor) -> indices [n_samples] (same kind as the input)."""
    is_np = isinstance(pc, np.ndarray)
    t = torch.as_tensor(pc, dtype=torch.float32)
    if not t.is_cuda:
        t = t.to(device)
    idx = furthest_point_sample(t[None], int(n_samples), 0 if start_idx is None else int(start_idx))[0]
    return idx.cpu().numpy().astype(np.uint64) if is_np else idx.long()
