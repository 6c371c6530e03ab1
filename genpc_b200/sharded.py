"""Multi-GPU forms of the path (one process per GPU, torch.distributed).

* Independent units (batched Chamfer / EMD / FPS / views / registration scans) shard by index with NO data-path
  collective: `shard_range` gives each rank its slice; results are gathered by the caller if it wants them.
* Million-point Chamfer (BASELINE config C5) shards the TARGETS: every rank holds both full clouds (12 MB per
  million points), scans all queries against its slice of the other cloud with global target indices, and the
  partial results are merged by ONE all-reduce-MIN over packed 64-bit words (dist_bits << 32 | idx) -- NCCL over
  NVLink (ncclInt64 / ncclMin).  dist >= 0 keeps the words non-negative as int64, and the low word makes the
  lowest global index win ties, so the merged result is bit-identical to the single-GPU kernel.
The host-side logic (slicing, merge, unpack) is device-agnostic and is covered on CPU with gloo (tests/).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib

EMPTY = -1  # all-ones word as int64


def shard_range(n, rank, world):
    """Contiguous near-even split of range(n): rank r owns [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_min_packed(packed, group=None):
    """In-place all-reduce-MIN of packed (dist, idx) words stored as int64.  Words are non-negative, except the
    'empty' marker -1 (all ones), which must lose against any real word: it is mapped to INT64_MAX around the
    collective."""
    big = torch.iinfo(torch.int64).max
    packed.masked_fill_(packed == EMPTY, big)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    packed.masked_fill_(packed == big, EMPTY)
    return packed


def unpack_packed(packed):
    """int64 words -> (dist float32, idx int32); pure torch so it also runs on CPU tensors (gloo tests)."""
    hi = (packed >> 32) & 0xFFFFFFFF
    lo = packed & 0xFFFFFFFF
    lo32 = torch.where(lo >= 2 ** 31, lo - 2 ** 32, lo).to(torch.int32)
    return _bits_to_float(hi), lo32


def _bits_to_float(hi):
    # hi holds 32-bit patterns in int64; values >= 2^31 only occur for the empty marker (NaN pattern)
    hi32 = torch.where(hi >= 2 ** 31, hi - 2 ** 32, hi).to(torch.int32)
    return hi32.view(torch.float32)


def nn_partial_packed(queries, targets_shard, idx_base, packed=None):
    """CUDA: scan all queries [B,Nq,3] against one target shard [B,Mt,3]; returns / updates packed int64 [B,Nq]."""
    _lib.require_cuda(queries, targets_shard)
    queries, targets_shard = queries.contiguous(), targets_shard.contiguous()
    B, Nq, _ = queries.shape
    Mt = targets_shard.shape[1]
    init = packed is None
    if init:
        packed = torch.empty(B, Nq, dtype=torch.int64, device=queries.device)
    with torch.cuda.device(queries.device):
        rc = _lib.lib().genpc_nn_partial_packed(_lib.ptr(queries), _lib.ptr(targets_shard), _lib.ptr(packed), B, Nq, Mt,
                                                int(idx_base), 1 if init else 0, _lib.current_stream(queries.device))
    _lib.check(rc, "genpc_nn_partial_packed")
    return packed


def nn_unpack(packed):
    """CUDA unpack kernel: packed int64 [..] -> (dist float32, idx int32)."""
    _lib.require_cuda(packed)
    d = torch.empty(packed.shape, dtype=torch.float32, device=packed.device)
    i = torch.empty(packed.shape, dtype=torch.int32, device=packed.device)
    with torch.cuda.device(packed.device):
        rc = _lib.lib().genpc_nn_unpack(_lib.ptr(packed), _lib.ptr(d), _lib.ptr(i), packed.numel(),
                                        _lib.current_stream(packed.device))
    _lib.check(rc, "genpc_nn_unpack")
    return d, i


def shard_range_aligned(n, rank, world, align=128):
    """Like shard_range but every boundary except the last is a multiple of `align` (row blocks of the symmetric scan)."""
    blocks = (n + align - 1) // align
    lo, hi = shard_range(blocks, rank, world)
    return min(lo * align, n), min(hi * align, n)


def sharded_chamfer_forward(xyz1, xyz2, group=None, phase_ms=None):
    """Sharded Chamfer forward, every point pair evaluated ONCE across the job.  xyz1 [1,N,3], xyz2 [1,M,3]: the FULL
    clouds, identical on every rank.  Rank r scans its 128-aligned row slice of xyz1 against all of xyz2 with the
    symmetric kernel: exact (dist1, idx1) for its rows, partial (dist, row block) minima for every point of xyz2.
    ONE all-reduce-MIN over [packed rows | packed cols] (16 MB for 1M + 1M points) completes both, a local fix-up
    resolves idx2.  Returns (dist1, dist2, idx1, idx2) identical on every rank and bit-identical to chamfer_3DDist.
    Batched inputs (B > 1) fall back to the target-sharded scan (sharded_chamfer_forward_targets).
    phase_ms: optional dict; filled with the device time of the three phases (scan / all-reduce / unpack + fix-up)
    measured with CUDA events -- synchronises, for measurements only."""
    if xyz1.shape[0] != 1:
        return sharded_chamfer_forward_targets(xyz1, xyz2, group)
    _lib.require_cuda(xyz1, xyz2)
    xyz1, xyz2 = xyz1.contiguous().float(), xyz2.contiguous().float()
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    N, M = xyz1.shape[1], xyz2.shape[1]
    dev = xyz1.device
    lo, hi = shard_range_aligned(N, rank, world)
    packed = torch.full((N + M,), EMPTY, dtype=torch.int64, device=dev)   # [rows of cloud 1 | cols = cloud 2]
    L = _lib.lib()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if phase_ms is not None else None
    with torch.cuda.device(dev):
        if ev: ev[0].record()
        rows = xyz1[0, lo:hi]
        rc = L.genpc_chamfer_sym_partial(_lib.ptr(rows) if hi > lo else None, _lib.ptr(xyz2),
                                         ctypes.c_void_p(packed.data_ptr() + lo * 8), ctypes.c_void_p(packed.data_ptr() + N * 8),
                                         1, hi - lo, M, lo, 0, _lib.current_stream(dev))
        _lib.check(rc, "genpc_chamfer_sym_partial")
        if ev: ev[1].record()
        allreduce_min_packed(packed, group)
        if ev: ev[2].record()
        d1, i1 = nn_unpack(packed[:N])
        d2 = torch.empty(M, dtype=torch.float32, device=dev)
        i2 = torch.empty(M, dtype=torch.int32, device=dev)
        rc = L.genpc_chamfer_sym_fixup(_lib.ptr(xyz1), _lib.ptr(xyz2), ctypes.c_void_p(packed.data_ptr() + N * 8), 1, N, M,
                                       _lib.ptr(d2), _lib.ptr(i2), _lib.current_stream(dev))
        _lib.check(rc, "genpc_chamfer_sym_fixup")
        if ev:
            ev[3].record()
            torch.cuda.synchronize(dev)
            phase_ms.update(scan=ev[0].elapsed_time(ev[1]), allreduce=ev[1].elapsed_time(ev[2]),
                            unpack_fixup=ev[2].elapsed_time(ev[3]))
    return d1[None], d2[None], i1[None], i2[None]


def sharded_chamfer_forward_targets(xyz1, xyz2, group=None):
    """Target-sharded Chamfer forward (one scan per direction and shard).  xyz1 [B,N,3], xyz2 [B,M,3]: the FULL clouds,
    identical on every rank.  Returns (dist1, dist2, idx1, idx2) identical on every rank, bit-identical to chamfer_3DDist."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    N, M = xyz1.shape[1], xyz2.shape[1]
    lo2, hi2 = shard_range(M, rank, world)
    lo1, hi1 = shard_range(N, rank, world)
    p1 = nn_partial_packed(xyz1, xyz2[:, lo2:hi2], lo2)
    p2 = nn_partial_packed(xyz2, xyz1[:, lo1:hi1], lo1)
    packed = torch.cat([p1.reshape(-1), p2.reshape(-1)])   # one collective for both directions
    allreduce_min_packed(packed, group)
    d, i = nn_unpack(packed)
    n1 = p1.numel()
    return d[:n1].view_as(p1), d[n1:].view_as(p2), i[:n1].view_as(p1), i[n1:].view_as(p2)


class ShardedChamferFunction(torch.autograd.Function):
    """chamfer_3DFunction (dist_chamfer_3D.py:26-64) for clouds that are REPLICATED on every rank: the forward is
    `sharded_chamfer_forward` (each point pair evaluated once across the job, one all-reduce-MIN); the backward needs no
    communication at all -- every rank holds both clouds and both index arrays after the forward, so it evaluates the
    full gradient locally (HBM bound, 44 bytes per point: 0.1 ms for 1M + 1M points) and all ranks end up with the
    same gradients, like the inputs they belong to."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, group=None):
        d1, d2, i1, i2 = sharded_chamfer_forward(xyz1, xyz2, group)
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2, i1, i2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        from . import chamfer_3D

        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        a, b = xyz1.contiguous().float(), xyz2.contiguous().float()
        g1, g2 = torch.zeros_like(a), torch.zeros_like(b)
        chamfer_3D.backward(a, b, g1, g2, graddist1.contiguous(), graddist2.contiguous(), i1.contiguous(), i2.contiguous())
        return g1, g2, None


class sharded_chamfer_3DDist(torch.nn.Module):
    """`chamfer_3DDist` (dist_chamfer_3D.py:67-74) across the GPUs of a process group: same call, same four outputs,
    bit-identical values, differentiable w.r.t. both (replicated) inputs."""

    def __init__(self, group=None):
        super().__init__()
        self.group = group

    def forward(self, input1, input2):
        return ShardedChamferFunction.apply(input1, input2, self.group)


# ---- independent units: batch sharding -----------------------------------------------------------
# The reference scatters a batch over its visible GPUs with nn.DataParallel(emdModule) inside ONE process
# (utils/loss_util.py:12).  Here every rank is one process with one GPU and the same (replicated) batch: rank r computes
# its contiguous batch slice, then ONE all-gather hands every rank the full result -- no data-path collective inside the
# compute, results bit-identical to the single-GPU call because batch entries never interact.

def _gather_batch(parts_shape, local, lo, hi, group):
    """All-gather of per-rank batch slices [lo:hi] of a [B, ...] tensor (uneven slices allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty(parts_shape, dtype=local.dtype, device=local.device)
    sizes = [shard_range(parts_shape[0], r, world) for r in range(world)]
    chunks = [out[a:b] for a, b in sizes]
    if all(b - a == sizes[0][1] - sizes[0][0] for a, b in sizes):
        dist.all_gather(chunks, local.contiguous(), group=group)       # equal slices: views into `out`, no extra copy
    else:
        for r, (a, b) in enumerate(sizes):                             # uneven: one broadcast per owner
            if a == b:
                continue
            buf = local.contiguous() if r == dist.get_rank(group) else chunks[r]
            dist.broadcast(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
            if r == dist.get_rank(group):
                chunks[r].copy_(buf)
    return out


def data_parallel_batch(fn, *tensors, group=None):
    """out = fn(*tensors) with the batch dimension (dim 0 of every tensor) sharded over the ranks of `group`: every rank
    runs fn on its slice, the outputs (a tensor or a tuple of tensors, batch first) are all-gathered.  Forward only."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    B = tensors[0].shape[0]
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        local = fn(*[t[lo:hi].contiguous() for t in tensors])
    else:   # more ranks than batch entries: run one entry to learn the output shapes, contribute nothing
        local = fn(*[t[:1].contiguous() for t in tensors])
        local = tuple(o[:0] for o in local) if isinstance(local, (tuple, list)) else local[:0]
    if isinstance(local, (tuple, list)):
        return tuple(_gather_batch((B,) + tuple(o.shape[1:]), o, lo, hi, group) for o in local)
    return _gather_batch((B,) + tuple(local.shape[1:]), local, lo, hi, group)


def data_parallel_emd(input1, input2, eps, iters, group=None):
    """emdModule()(input1, input2, eps, iters) with the batch sharded over the ranks: the one-process-per-GPU equivalent of
    the reference's nn.DataParallel(emdModule) (utils/loss_util.py:12).  -> (dist [B,n], assignment [B,n]) on every rank."""
    from .loss_functions import emdModule

    mod = emdModule()
    with torch.no_grad():
        return data_parallel_batch(lambda a, b: mod(a, b, eps, iters), input1, input2, group=group)


def data_parallel_fps(xyz, K, start=0, group=None):
    """furthest_point_sample over a batch sharded by cloud (FPS batches / views are independent units)."""
    from .fps import furthest_point_sample

    return data_parallel_batch(lambda x: furthest_point_sample(x, K, start), xyz, group=group)
