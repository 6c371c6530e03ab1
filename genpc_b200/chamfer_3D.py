"""Mirror of the reference's pybind module `chamfer_3D` (chamfer_cuda.cpp:30-33).

    forward(xyz1, xyz2, dist1, dist2, idx1, idx2) -> int
    backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2) -> int

Same argument order, caller-allocated outputs written in place, return 1 on success (the reference
returns 1 / 0 and prints, chamfer3D.cu:145-151); failures raise instead of being silently ignored.
Work is enqueued on the CURRENT torch stream (the reference uses the legacy default stream).
"""
import torch

from . import _lib


def _chk(t, dtype, name):
    if t.dtype != dtype or not t.is_contiguous():
        raise _lib.GenpcError(f"{name} must be a contiguous {dtype} tensor")


def forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
    _lib.require_cuda(xyz1, xyz2, dist1, dist2, idx1, idx2)
    for t, n in ((xyz1, "xyz1"), (xyz2, "xyz2"), (dist1, "dist1"), (dist2, "dist2")):
        _chk(t, torch.float32, n)
    for t, n in ((idx1, "idx1"), (idx2, "idx2")):
        _chk(t, torch.int32, n)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    L = _lib.lib()
    with torch.cuda.device(xyz1.device):
        nbytes = L.genpc_chamfer_workspace_bytes(B, N, M)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=xyz1.device)
        rc = L.genpc_chamfer_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(dist1), _lib.ptr(dist2),
                                     _lib.ptr(idx1), _lib.ptr(idx2), B, N, M, _lib.ptr(ws), nbytes,
                                     _lib.current_stream(xyz1.device))
    _lib.check(rc, "genpc_chamfer_forward")
    return 1


def backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
    _lib.require_cuda(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
    for t, n in ((xyz1, "xyz1"), (xyz2, "xyz2"), (gradxyz1, "gradxyz1"), (gradxyz2, "gradxyz2"),
                 (graddist1, "graddist1"), (graddist2, "graddist2")):
        _chk(t, torch.float32, n)
    for t, n in ((idx1, "idx1"), (idx2, "idx2")):
        _chk(t, torch.int32, n)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    with torch.cuda.device(xyz1.device):
        rc = _lib.lib().genpc_chamfer_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(graddist1),
                                               _lib.ptr(graddist2), _lib.ptr(idx1), _lib.ptr(idx2),
                                               _lib.ptr(gradxyz1), _lib.ptr(gradxyz2), B, N, M,
                                               _lib.current_stream(xyz1.device))
    _lib.check(rc, "genpc_chamfer_backward")
    return 1
