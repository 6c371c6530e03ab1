"""Mirror of the reference's pybind module `emd` (emd.cpp:25-29).

    forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments,
            unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters) -> int
    backward(xyz1, xyz2, gradxyz, graddist, idx) -> int

Same 16 / 5 positional arguments and caller-allocated buffers (emd_module.py:43-54); `unass_cnt_sum` and
`cnt_tmp` are accepted and left untouched (the single persistent kernel does not need them).  Returns 1 on
success; the reference's shape violations (emd_cuda.cu:236-249, which it reports as -1 after a printf) raise.
"""
import torch

from . import _lib


def _chk(t, dtype, name):
    if t.dtype != dtype or not t.is_contiguous() or not t.is_cuda:
        raise _lib.GenpcError(f"{name} must be a contiguous CUDA {dtype} tensor")


def forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx,
            unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters):
    for t, n in ((xyz1, "xyz1"), (xyz2, "xyz2"), (dist, "dist"), (price, "price"), (bid_increments, "bid_increments"),
                 (max_increments, "max_increments")):
        _chk(t, torch.float32, n)
    for t, n in ((assignment, "assignment"), (assignment_inv, "assignment_inv"), (bid, "bid"), (unass_idx, "unass_idx"),
                 (unass_cnt, "unass_cnt"), (max_idx, "max_idx")):
        _chk(t, torch.int32, n)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if unass_cnt.numel() < B:
        raise _lib.GenpcError("unass_cnt must hold at least B ints")
    L = _lib.lib()
    with torch.cuda.device(xyz1.device):
        nbytes = L.genpc_emd_workspace_bytes_n(B, n)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=xyz1.device)
        rc = L.genpc_emd_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(dist), _lib.ptr(assignment), _lib.ptr(price),
                                 _lib.ptr(assignment_inv), _lib.ptr(bid), _lib.ptr(bid_increments),
                                 _lib.ptr(max_increments), _lib.ptr(unass_idx), _lib.ptr(unass_cnt), _lib.ptr(max_idx),
                                 B, n, m, float(eps), int(iters), _lib.ptr(ws), nbytes,
                                 _lib.current_stream(xyz1.device))
    _lib.check(rc, "genpc_emd_forward")
    return 1


def backward(xyz1, xyz2, gradxyz, graddist, idx):
    for t, n in ((xyz1, "xyz1"), (xyz2, "xyz2"), (gradxyz, "gradxyz"), (graddist, "graddist")):
        _chk(t, torch.float32, n)
    _chk(idx, torch.int32, "idx")
    B, n, _ = xyz1.shape
    with torch.cuda.device(xyz1.device):
        rc = _lib.lib().genpc_emd_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(gradxyz), _lib.ptr(graddist),
                                           _lib.ptr(idx), B, n, _lib.current_stream(xyz1.device))
    _lib.check(rc, "genpc_emd_backward")
    return 1
