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


_feeds = {}


def _feed(device):
    """One host-feed handle (copy stream + gate words) per device, created on first use."""
    import ctypes

    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _feeds:
        h = ctypes.c_void_p()
        with torch.cuda.device(idx):
            _lib.check(_lib.lib().genpc_host_feed_create(ctypes.byref(h)), "genpc_host_feed_create")
        _feeds[idx] = h
    return _feeds[idx]


def forward_host(h_xyz1, h_xyz2, xyz1, xyz2, dist1, dist2, idx1, idx2, chunks=6):
    """No pybind counterpart: `forward` fed from HOST tensors.  h_xyz1 / h_xyz2 (CPU, float32, contiguous; pinned for an
    asynchronous copy) are streamed into the caller-allocated CUDA tensors xyz1 / xyz2 chunk by chunk while the scan
    already consumes the chunks that have arrived (genpc_chamfer_forward_host).  Outputs as `forward`."""
    _lib.require_cuda(xyz1, xyz2, dist1, dist2, idx1, idx2)
    if h_xyz1.is_cuda or h_xyz2.is_cuda:
        raise _lib.GenpcError("forward_host takes CPU tensors as h_xyz1 / h_xyz2")
    for t, n in ((h_xyz1, "h_xyz1"), (h_xyz2, "h_xyz2"), (xyz1, "xyz1"), (xyz2, "xyz2"), (dist1, "dist1"), (dist2, "dist2")):
        _chk(t, torch.float32, n)
    for t, n in ((idx1, "idx1"), (idx2, "idx2")):
        _chk(t, torch.int32, n)
    if h_xyz1.shape != xyz1.shape or h_xyz2.shape != xyz2.shape:
        raise _lib.GenpcError("host and device clouds must have the same shape")
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    L = _lib.lib()
    with torch.cuda.device(xyz1.device):
        nbytes = L.genpc_chamfer_workspace_bytes(B, N, M)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=xyz1.device)
        rc = L.genpc_chamfer_forward_host(_feed(xyz1.device), _lib.ptr(h_xyz1), _lib.ptr(h_xyz2), _lib.ptr(xyz1),
                                          _lib.ptr(xyz2), _lib.ptr(dist1), _lib.ptr(dist2), _lib.ptr(idx1), _lib.ptr(idx2),
                                          B, N, M, int(chunks), _lib.ptr(ws), nbytes, _lib.current_stream(xyz1.device))
    _lib.check(rc, "genpc_chamfer_forward_host")
    return 1


def host_feed_error(device):
    """True if a host-fed launch on `device` ever gave up waiting for its data (synchronises the current stream)."""
    rc = _lib.lib().genpc_host_feed_error(_feed(device), _lib.current_stream(device))
    if rc not in (0, 1):
        _lib.check(rc, "genpc_host_feed_error")
    return bool(rc)


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
