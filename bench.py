#!/usr/bin/env python
"""bench.py -- headline benchmark of the GenPC geometric hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "C2"): batched Chamfer fwd+bwd, B=32 per GPU, 2048 x 16384 points (PCN shape,
BASELINE.json configs[1]); a step = one forward + backward pass over one batch; metric = directed
point-pair evaluations per second (2*B*N*M per step), whole job over all ranks (weak scaling: every rank
owns its own 32 scans, no data-path collective).

Printed keys beyond the base contract:
  roofline      dominant kernel (nn_sym_kernel): FP32-pipe bound -> achieved/peak in TFLOP/s (8 FLOP/pair)
  roofline_bwd  gradient scatter kernel: HBM bound -> GB/s against MEASURED_PEAKS.json
  cpu_baseline  the oracle port (OpenMP, all host cores) and torch.cdist on the same workload sample
  ref_cuda_ext  the unmodified reference CUDA extension (oracle/_ref) on the same GPU and inputs
  e2e           same metric through the public API from pinned HOST buffers (H2D + D2H inside the timing)
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B, N, M = 32, 2048, 16384
FLOP_PER_PAIR = 8            # 3 sub, 3 mul, 2 add
FP32_NOMINAL_TFLOPS = 74.4   # 148 SM x 128 lanes x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)
L2_FLUSH_BYTES = 256 << 20


def read_ncu_profile(stem, kernel_substr):
    """Latest committed ncu --set full summary profiles/r*_<stem>_ncu.txt (written by tools/ncu_summary.py from the capture
    of THIS bench command): -> {metric: (value, unit)} of the first kernel whose name contains kernel_substr, plus 'file'.
    The bench line quotes DRAM traffic / pipe utilisation from that file at run time instead of pasting constants."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{stem}_ncu.txt")))
    for f in reversed(files):
        cur, out = None, {}
        for line in open(f):
            if line.startswith("== "):
                if out:
                    break
                cur = line if kernel_substr in line else None
                continue
            m = re.match(r"\s+(\S+)\s+([-0-9.eE+]+)\s*(\S*)", line)
            if cur and m:
                out[m.group(1)] = (float(m.group(2)), m.group(3))
        if out:
            out["file"] = os.path.relpath(f, ROOT)
            return out
    return None


def ncu_dram_bytes(prof):
    if not prof:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k not in prof:
            return None
        tot += prof[k][0] * mult.get(prof[k][1], 1.0)
    return tot


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def make_inputs(rank):
    from genpc_b200.synthetic import pcn_batch

    part, comp = pcn_batch(1000 * rank, B, N, M)
    return part, comp


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def measured_fp32_peak():
    p = os.path.join(ROOT, "profiles", "fp32_peak_b200.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------
# CPU baselines (rank 0, N=1 only) -- oracle/ is used here strictly as the thing that is timed beside us
# ---------------------------------------------------------------------------------------------------
def cpu_baseline(part, comp, budget_b=None, keep_forward=False):
    import oracle

    cores = os.cpu_count()
    nb = budget_b or B
    a, b = part[:nb], comp[:nb]
    t0 = time.perf_counter()
    d1, d2, i1, i2 = oracle.chamfer_forward(a, b)
    g1 = np.full(d1.shape, 1.0 / d1.size, np.float32)
    g2 = np.full(d2.shape, 1.0 / d2.size, np.float32)
    oracle.chamfer_backward(a, b, g1, g2, i1, i2)
    dt = time.perf_counter() - t0
    pairs = 2.0 * nb * N * M
    out = {"value": pairs / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
           "sample": f"oracle C port (OpenMP), fwd+bwd on {nb} of the {B} scans of C2, {dt:.2f} s"}
    if keep_forward:
        out["_oracle_forward"] = (d1, d2, i1, i2)   # popped by the caller: full-batch parity assert of the GPU arm
    # the reference's stated CPU path (BASELINE.json): torch.cdist (no-mm) + min both ways
    try:
        torch.set_num_threads(cores)
        nb2 = min(nb, 4)
        ta, tb = torch.from_numpy(part[:nb2]), torch.from_numpy(comp[:nb2])
        t0 = time.perf_counter()
        D = torch.cdist(ta, tb, compute_mode="donot_use_mm_for_euclid_dist")
        D.min(2), D.min(1)
        dt2 = time.perf_counter() - t0
        out["torch_cdist"] = {"value": 2.0 * nb2 * N * M / dt2, "unit": "pairs/s", "cores": torch.get_num_threads(),
                              "sample": f"torch.cdist(no-mm)+min both ways, forward only, {nb2} scans, {dt2:.2f} s"}
    except Exception as e:  # pragma: no cover
        out["torch_cdist"] = {"error": str(e)}
    # the strongest CPU ALGORITHM for the same result: KD-tree queries on all cores (float64 tree:
    # distances agree to rounding, ties may resolve differently -- a speed reference, not a parity arm)
    try:
        from scipy.spatial import cKDTree

        nb3 = min(nb, 8)
        t0 = time.perf_counter()
        for s in range(nb3):
            cKDTree(comp[s]).query(part[s], k=1, workers=-1)
            cKDTree(part[s]).query(comp[s], k=1, workers=-1)
        dt3 = time.perf_counter() - t0
        out["scipy_ckdtree"] = {"value": 2.0 * nb3 * N * M / dt3, "unit": "equivalent pairs/s", "cores": cores,
                                "sample": f"cKDTree build + query(k=1, workers=-1) both ways, forward only, {nb3} scans, {dt3:.2f} s"}
    except Exception as e:  # pragma: no cover
        out["scipy_ckdtree"] = {"error": str(e)}
    return out


# ---------------------------------------------------------------------------------------------------
# The reference's stock path: dist_chamfer_3D.py:26-64 restated around the UNMODIFIED extension
# (the file itself cannot be imported on Python 3.12: importlib.find_loader, :6)
# ---------------------------------------------------------------------------------------------------
def import_reference_api():
    """The UNMODIFIED reference Python API (`loss_functions.chamfer_3DDist`) from baseline/_ref (pip-installed copy of
    /root/reference) on top of its UNMODIFIED compiled extensions (oracle/_ref).  The only
    shim is for the interpreter: Python 3.12 removed importlib.find_loader, which dist_chamfer_3D.py:6 calls."""
    import importlib
    import importlib.util

    ref_pkg = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_pkg, "loss_functions")):
        return None
    for name in ("chamfer_3D", "emd"):
        d = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(os.path.join(d, name + ".so")):
            return None
        if d not in sys.path:
            sys.path.insert(0, d)
    if not hasattr(importlib, "find_loader"):
        importlib.find_loader = lambda name, path=None: importlib.util.find_spec(name)
    mine = [k for k in sys.modules if k == "loss_functions" or k.startswith("loss_functions.")]
    for k in mine:
        del sys.modules[k]
    sys.path.insert(0, ref_pkg)
    import contextlib

    try:
        with contextlib.redirect_stdout(sys.stderr):  # the reference prints at import; stdout carries ONE JSON line
            import loss_functions  # noqa: F401  (the reference's package, not genpc_b200.loss_functions)

            from utils.loss_util import Completionloss as RefCompletionloss  # the reference's own loss facade

        assert os.path.realpath(loss_functions.__file__).startswith(os.path.realpath(ref_pkg))
        ref_loss = RefCompletionloss("cd_l2")          # utils/loss_util.py:8-23 (also builds DataParallel(emdModule))
        return ref_loss.get_loss                        # == chamfer_l2: mean(d1) + mean(d2)
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"[bench] reference API import failed: {e}\n")
        return None
    finally:
        sys.path.remove(ref_pkg)


def make_ref_function(ext):
    class RefChamfer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, xyz1, xyz2):
            batchsize, n, _ = xyz1.size()
            _, m, _ = xyz2.size()
            device = xyz1.device
            dist1 = torch.zeros(batchsize, n).to(device)            # CPU alloc + H2D, as the reference (:33-42)
            dist2 = torch.zeros(batchsize, m).to(device)
            idx1 = torch.zeros(batchsize, n).type(torch.IntTensor).to(device)
            idx2 = torch.zeros(batchsize, m).type(torch.IntTensor).to(device)
            torch.cuda.set_device(device)
            ext.forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
            ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
            return dist1, dist2, idx1, idx2

        @staticmethod
        def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
            xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
            graddist1 = graddist1.contiguous()
            graddist2 = graddist2.contiguous()
            device = graddist1.device
            gradxyz1 = torch.zeros(xyz1.size()).to(device)           # (:56-60)
            gradxyz2 = torch.zeros(xyz2.size()).to(device)
            ext.backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
            return gradxyz1, gradxyz2

    return lambda a, b: RefChamfer.apply(a.contiguous(), b.contiguous())


def chamfer_l2_loss(fn, a, b):
    """utils/loss_util.py:31-33 chamfer_l2 = mean(d1) + mean(d2)."""
    d1, d2, _, _ = fn(a, b)
    return torch.mean(d1) + torch.mean(d2)


def run_gpu_arm(args, loss_fn, rank, world, dev, part, comp, tag, host_loss_fn=None, graphed_factory=None):
    """Times K steps device-resident (`value`) and K steps end-to-end from pinned host buffers (`e2e`).
    loss_fn(a, b) -> scalar loss is the public API call: Completionloss('cd_l2').get_loss of either implementation.
    host_loss_fn(ha, hb) -> (loss, a_cuda, b_cuda): our host-fed public call (H2D copy overlapped with the scan); when
    given it is what `e2e` measures, and the plain `.to(device)` + loss_fn flow is reported beside it as `e2e_plain`."""
    import torch.distributed as dist

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    a = torch.from_numpy(part).to(dev).requires_grad_(True)
    b = torch.from_numpy(comp).to(dev).requires_grad_(True)
    ha, hb = torch.from_numpy(part).pin_memory(), torch.from_numpy(comp).pin_memory()
    hloss = torch.zeros(1).pin_memory()

    def step_resident():
        a.grad = None
        b.grad = None
        loss = loss_fn(a, b)
        loss.backward()
        return loss

    def step_e2e():
        da = ha.to(dev, non_blocking=True).requires_grad_(True)
        db = hb.to(dev, non_blocking=True).requires_grad_(True)
        loss = loss_fn(da, db)
        loss.backward()
        hloss.copy_(loss.detach(), non_blocking=True)
        return loss

    def step_e2e_hostfed():
        loss, da, db = host_loss_fn(ha, hb)
        loss.backward()
        hloss.copy_(loss.detach(), non_blocking=True)
        return loss

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step, K, W):
        for _ in range(W):
            flush.zero_()
            step()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for i in range(K):
            flush.zero_()  # L2 flush between timed iterations, outside the event bracket
            ev[i][0].record()
            step()
            ev[i][1].record()
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_eager = timed(step_resident, args.steps, args.warmup)
    ms_res = ms_eager
    if graphed_factory is not None:
        # the same step (same kernels, same device-resident inputs) captured once into a CUDA graph: one graph launch per step
        gstep = graphed_factory(a.detach(), b.detach())
        ms_res = timed(lambda: gstep(), args.steps, args.warmup)
    clocks = sampler.stop()
    ms_plain = timed(step_e2e, args.steps, args.warmup)
    ms_e2e = timed(step_e2e_hostfed, args.steps, args.warmup) if host_loss_fn is not None else ms_plain
    pairs_per_step = 2.0 * B * N * M * world
    res = {
        "value": pairs_per_step * args.steps / (ms_res * 1e-3),
        "ms_per_step": ms_res / args.steps,
        "e2e": {"value": pairs_per_step * args.steps / (ms_e2e * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": int(ha.numel() * 4 + hb.numel() * 4), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "clocks": clocks,
    }
    if graphed_factory is not None:
        graph_api = "GraphedLossStep(Completionloss('cd_l2'), gen, gt)(): the fused step replayed as one CUDA graph"
        eager_api = "Completionloss('cd_l2').get_loss(gen, gt); loss.backward()  (Python + autograd per step)"
        res["api_value"] = graph_api
        res["eager"] = {"value": pairs_per_step * args.steps / (ms_eager * 1e-3), "unit": "pairs/s", "ms_per_step": ms_eager / args.steps,
                        "api": eager_api}
        if ms_eager < ms_res:
            # both are public calls on the same device-resident inputs; `value` is the faster one, the other stays beside it
            res["graphed"] = {"value": res["value"], "unit": "pairs/s", "ms_per_step": res["ms_per_step"], "api": graph_api}
            res["value"], res["ms_per_step"], res["api_value"] = res["eager"]["value"], res["eager"]["ms_per_step"], eager_api
    if host_loss_fn is not None:
        res["e2e"]["api"] = ("Completionloss('cd_l2').get_loss_from_host(gen_pinned, gt_pinned); loss.backward(); loss -> host "
                             "(genpc_chamfer_forward_host: H2D copy in 6 chunks, each chunk's sort + pruned exact scan queued behind its copy on its own stream)")
        res["e2e_plain"] = {"value": pairs_per_step * args.steps / (ms_plain * 1e-3), "unit": "pairs/s",
                            "ms_per_step": ms_plain / args.steps,
                            "api": "gen.to(device); gt.to(device); get_loss; backward; loss -> host (copy, then compute)"}
    return res, (a, b, flush)


def time_kernels_ours(dev, a, b, flush, iters=20):
    """CUDA-event time of the forward op (memset + nn_sym_kernel + unpack + fixup; the scan is ~95 % of it) and of
    the backward op (chamfer_grad_kernel) on the launching stream, L2 flushed before each."""
    from genpc_b200 import chamfer_3D

    a, b = a.detach(), b.detach()
    d1 = torch.zeros(B, N, device=dev); d2 = torch.zeros(B, M, device=dev)
    i1 = torch.zeros(B, N, dtype=torch.int32, device=dev); i2 = torch.zeros(B, M, dtype=torch.int32, device=dev)
    g1 = torch.full((B, N), 1.0 / (B * N), device=dev); g2 = torch.full((B, M), 1.0 / (B * M), device=dev)
    gx1 = torch.zeros_like(a); gx2 = torch.zeros_like(b)
    tf, tb = [], []
    for it in range(iters + 3):
        flush.zero_()
        e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e0.record()
        chamfer_3D.forward(a, b, d1, d2, i1, i2)
        e1.record()
        gx1.zero_(); gx2.zero_()
        flush.zero_()
        e2.record()
        chamfer_3D.backward(a, b, gx1, gx2, g1, g2, i1, i2)
        e3.record()
        torch.cuda.synchronize(dev)
        if it >= 3:
            tf.append(e0.elapsed_time(e1)); tb.append(e2.elapsed_time(e3))
    return sum(tf) / len(tf), sum(tb) / len(tb)


def registration_metric(rank, world, dev, iters=10, total_scans=64, pts=16384):
    """BASELINE config C3 (secondary metric): `total_scans` scans data-parallel over the ranks (strong scaling, no
    collective), pose+scale Adam iterations on the Chamfer loss at 16384 x 16384 points, one start per scan.
    Returns scan-iterations per second, whole job (max time over ranks)."""
    import torch.distributed as dist

    from genpc_b200 import _lib
    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch
    from genpc_b200.sharded import shard_range
    from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric

    lo, hi = shard_range(total_scans, rank, world)
    comp = np.stack([superquadric(5000 + s, pts) for s in range(lo, hi)])
    part = np.stack([rigid_perturb(partial_view(comp[s - lo], s, pts), s)[0] for s in range(lo, hi)])
    rb = RegistrationBatch(torch.from_numpy(comp).to(dev), torch.from_numpy(part).to(dev), n_starts=1, lr=0.01,
                           max_iters=iters + 4)
    rb.run(3)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rb.run(iters)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    L = rb.losses()
    return {"metric": "registration_scan_iters_per_sec", "value": total_scans * iters / (ms * 1e-3), "unit": "scan-iters/s",
            "scans": total_scans, "pts": pts, "iters_timed": iters, "ms_per_iter": ms / iters, "scaling": "strong",
            "pairs_per_s": total_scans * iters * 2.0 * pts * pts / (ms * 1e-3),
            "launches_per_iter": int(_lib.lib().genpc_register_launches_per_iter(hi - lo, pts, pts)),
            "loss_first_last_scan0": [float(L[0, 0]), float(L[0, rb.t - 1])],
            "workload": "C3: pose+scale Adam steps on 3*(CDp-L1(pts->ref)+0.5*CDp-L1(ref->pts)), 1 start per scan"}


def ev_best(fn, reps=3, warm=1):
    """Best-of CUDA-event time (ms) of fn() on the current stream."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def c5_sharded_metric(rank, world, dev, n=1_000_000, reps=2):
    """BASELINE config C5: n x n-point Chamfer forward, rows of cloud 1 sharded over the ranks (every point pair evaluated
    once across the job), ONE NCCL all-reduce-MIN over the packed (dist, idx) words, local fix-up.  Strong scaling.  All
    ranks call this (collective inside); device time, max over ranks; rank 0 checks 2000 sampled points bit-exactly against
    the oracle (a full CPU scan of 1e12 pairs is out of reach)."""
    import torch.distributed as dist

    from genpc_b200.sharded import sharded_chamfer_forward
    from genpc_b200.synthetic import lidar_scene_pair

    a, b = lidar_scene_pair(n, 0)
    ta, tb = a[None].to(dev), b[None].to(dev)
    out = sharded_chamfer_forward(ta, tb)
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sharded_chamfer_forward(ta, tb)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    ph = {}
    if world > 1:
        dist.barrier()
    sharded_chamfer_forward(ta, tb, phase_ms=ph)   # one more pass with per-phase CUDA events
    pt = torch.tensor([ph["scan"], ph["allreduce"], ph["unpack_fixup"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    ms = min(ts)
    res = {"workload": f"C5: {n} x {n} Chamfer forward (LiDAR-like scene), row-sharded symmetric scan + all-reduce-MIN of packed "
                       "(dist, idx) words + local fix-up", "n": n, "n_gpus": world, "scaling": "strong", "ms": ms,
           "pairs_per_s": 2.0 * n * n / (ms * 1e-3),
           "phase_ms_max_over_ranks": {"scan": float(pt[0]), "allreduce_min_packed": float(pt[1]), "unpack_fixup": float(pt[2])},
           "collective": "torch.distributed all_reduce(MIN, int64) over NCCL" if world > 1 else "none (1 rank)",
           "allreduce_bytes": 8 * 2 * n}
    if rank == 0:
        import oracle

        sel = np.random.default_rng(0).choice(n, 2000, replace=False)
        ed, ei = oracle.nn_distance(a[sel][None].numpy(), b[None].numpy())
        ed2, ei2 = oracle.nn_distance(b[sel][None].numpy(), a[None].numpy())
        d1, d2, i1, i2 = out
        ok = (np.array_equal(d1[0, sel].cpu().numpy(), ed[0]) and np.array_equal(i1[0, sel].cpu().numpy(), ei[0]) and
              np.array_equal(d2[0, sel].cpu().numpy(), ed2[0]) and np.array_equal(i2[0, sel].cpu().numpy(), ei2[0]))
        res["sample_check_bit_exact_vs_oracle"] = bool(ok)
        assert ok, "C5: sharded Chamfer differs from the oracle on the sampled points"
        # the same forward through the plain operator on ONE GPU: clouds of this size take the Hilbert-sorted, two-level
        # pruned exact scan (csrc/nn_grid.cuh) -- every output compared with the sharded exhaustive result above
        from genpc_b200 import _lib
        from genpc_b200.loss_functions import chamfer_3DDist

        cd = chamfer_3DDist()
        stats = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.lib().genpc_chamfer_prune_stats(_lib.ptr(stats))
        po = cd(ta, tb)
        torch.cuda.synchronize(dev)
        _lib.lib().genpc_chamfer_prune_stats(None)
        tp = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            po = cd(ta, tb)
            e1.record()
            torch.cuda.synchronize(dev)
            tp.append(e0.elapsed_time(e1))
        same = all(bool(torch.equal(x, y)) for x, y in zip(po, out))
        st = stats.cpu().tolist()
        tot_pairs = 2 * ((n + 31) // 32) * ((n + 63) // 64)
        res["pruned_single_gpu"] = {"ms": min(tp), "pairs_per_s_algorithmic": 2.0 * n * n / (min(tp) * 1e-3),
                                    "speedup_vs_exhaustive_same_n_gpus": ms / min(tp), "identical_to_exhaustive": same,
                                    "group_block_pairs_visited": st[0], "of": tot_pairs, "visited_fraction": st[0] / tot_pairs,
                                    "query_groups": st[2], "note": "chamfer_3DDist on one GPU; bit-identical outputs; the sharded "
                                    "exhaustive path above is what runs when the sampled probe finds clouds that do not overlap"}
        assert same, "C5: pruned scan differs from the exhaustive result"
    return res


def _emd_buffers(Bn, n, dev):
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)   # noqa: E731
    return dict(dist=z(Bn, n), asg=z(Bn, n, dt=torch.int32) - 1, price=z(Bn, n), asg_inv=z(Bn, n, dt=torch.int32) - 1,
                bid=z(Bn, n, dt=torch.int32), binc=z(Bn, n), minc=z(Bn, n), uidx=z(Bn * n, dt=torch.int32),
                ucnt=z(512, dt=torch.int32), ucs=z(512, dt=torch.int32), ctmp=z(512, dt=torch.int32), midx=z(Bn * n, dt=torch.int32))


def emd_c5_metric(dev, n=8192, eps=0.005, iters=50):
    """BASELINE config C5, EMD leg (emd_module.py:98-118 `test_emd` inputs: torch.rand pairs, eps 0.005, 50 iterations):
    ours through the mirror of the pybind module, the unmodified reference extension beside it on the same GPU."""
    import oracle
    from genpc_b200 import emd as ours

    ref = oracle.load_ref_ext("emd")
    out = {"workload": f"C5 EMD: auction matching of torch.rand(B,{n},3) pairs, eps {eps}, {iters} iterations", "n": n}
    for Bn in (1, 32):
        g = torch.Generator().manual_seed(0)
        x1, x2 = torch.rand(Bn, n, 3, generator=g).to(dev), torch.rand(Bn, n, 3, generator=g).to(dev)
        row = {}
        for name, mod in (("ours", ours), ("reference_ext", ref)):
            if mod is None:
                continue
            best = None
            for rep in range(3):
                bf = _emd_buffers(Bn, n, dev)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                mod.forward(x1, x2, bf["dist"], bf["asg"], bf["price"], bf["asg_inv"], bf["bid"], bf["binc"], bf["minc"], bf["uidx"],
                            bf["ucnt"], bf["ucs"], bf["ctmp"], bf["midx"], eps, iters)
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1)
                best = ms if best is None or (rep > 0 and ms < best) else best
            row[name + "_ms"] = best
            row[name + "_emd_cost"] = float(torch.sqrt(bf["dist"]).mean())
            row[name + "_unassigned"] = int((bf["asg"] < 0).sum())
        if "reference_ext_ms" in row:
            row["speedup_vs_reference_ext"] = row["reference_ext_ms"] / row["ours_ms"]
            row["cost_rel_diff"] = abs(row["ours_emd_cost"] - row["reference_ext_emd_cost"]) / row["reference_ext_emd_cost"]
        row["matchings_per_s"] = Bn / (row["ours_ms"] * 1e-3)
        out[f"B{Bn}"] = row
    return out


def c4_metric(dev):
    """BASELINE config C4: 8-view 512 x 512 point -> depth z-buffer render + depth -> point unprojection of the reference scan
    (data/01184.ply, 71 372 points; committed fixture) and FPS 16384 -> 2048; z-buffer owners and FPS indices checked bit
    for bit against the oracle."""
    import oracle
    from genpc_b200 import depth as D
    from genpc_b200.fps import furthest_point_sample
    from genpc_b200.synthetic import superquadric

    fix = os.path.join(ROOT, "tests", "golden", "scan_01184_xyz.npz")
    scan = np.load(fix)["xyz"] if os.path.exists(fix) else superquadric(0, 71372)
    pts = torch.from_numpy(scan).to(dev)
    V, res = 8, 512
    cams, _ = D.create_cameras(V, 1.6, 49.1, res, dev)
    out = {"workload": f"C4: {V} views x {res}^2, project + z-buffer render + unproject of {scan.shape[0]} points "
                       "(data/01184.ply); FPS 16384 -> 2048", "views": V, "res": res, "points": int(scan.shape[0])}
    for ps in (1, 2):
        def run(ps=ps):
            ndc, uv, b = D.project_uv(cams, pts, True, 0.15)
            r = D.zbuffer_render(uv, ndc, res, ps)
            return ndc, uv, b, r, D.unproject(cams, b, r["zbuf"], ndc, True)
        ms = ev_best(run, reps=5, warm=2)
        ndc, uv, b, r, un = run()
        e_ndc, e_uv, e_b = oracle.project_uv(cams.cpu().numpy(), scan, True, 0.15)
        e_zb = oracle.zbuffer(e_uv, e_ndc, res, ps)
        zb_ok = bool(np.array_equal(r["zbuf"].cpu().numpy().view(np.uint64), np.asarray(e_zb).view(np.uint64)))
        out[f"point_size_{ps}"] = {"ms": ms, "points_per_s": V * scan.shape[0] / (ms * 1e-3),
                                   "pixels_per_s": V * res * res / (ms * 1e-3), "zbuffer_bit_exact_vs_oracle": zb_ok}
        assert zb_ok, "C4: z-buffer differs from the oracle"
    for name, cloud in (("uniform_cube", np.random.default_rng(0).random((1, 16384, 3), dtype=np.float32)),
                        ("c2_shape", superquadric(0, 16384)[None])):
        x = torch.from_numpy(cloud).to(dev)
        ms = ev_best(lambda: furthest_point_sample(x, 2048, 0), reps=3, warm=1)
        idx = furthest_point_sample(x, 2048, 0).cpu().numpy()
        ok = bool(np.array_equal(idx.astype(np.int64), np.asarray(oracle.fps(cloud, 2048, 0)).astype(np.int64)))
        out[f"fps_16384_to_2048_{name}"] = {"ms": ms, "picks_per_s": 2048 / (ms * 1e-3), "indices_bit_exact_vs_oracle": ok}
        assert ok, "C4: FPS indices differ from the oracle"
    xb = torch.rand(32, 16384, 3, device=dev)
    out["fps_16384_to_2048_B32_ms"] = ev_best(lambda: furthest_point_sample(xb, 2048, 0), reps=2, warm=1)
    return out


def c1_metric(dev):
    """BASELINE config C1 on the real scan: data/01184.ply (71 372 pts, committed fixture) vs the 16 384-pt synthetic shape,
    batch 1, CD-L1 forward + backward through the module; indices checked against the reference extension's golden."""
    import hashlib

    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.synthetic import superquadric

    fix = os.path.join(ROOT, "tests", "golden", "scan_01184_xyz.npz")
    gold = os.path.join(ROOT, "tests", "golden", "c1_ref_chamfer.npz")
    if not os.path.exists(fix):
        return {"error": "tests/golden/scan_01184_xyz.npz missing"}
    scan, shape = np.load(fix)["xyz"], superquadric(0, 16384)
    a = torch.from_numpy(scan[None]).to(dev).requires_grad_(True)
    b = torch.from_numpy(shape[None]).to(dev).requires_grad_(True)
    cd = chamfer_3DDist()

    def step():
        a.grad = None
        b.grad = None
        d1, d2, _, _ = cd(a, b)
        ((torch.sqrt(d1).mean() + torch.sqrt(d2).mean()) / 2).backward()

    ms = ev_best(step, reps=10, warm=3)
    d1, d2, i1, i2 = cd(a.detach(), b.detach())
    res = {"workload": "C1: CD-L1 fwd+bwd, data/01184.ply (71372 pts) vs 16384-pt synthetic shape, B=1", "ms_fwd_bwd": ms,
           "pairs_per_s": 2.0 * scan.shape[0] * 16384 / (ms * 1e-3),
           "cd_l1": float((torch.sqrt(d1).mean() + torch.sqrt(d2).mean()) / 2)}
    if os.path.exists(gold):
        g = np.load(gold)
        h = hashlib.sha256()
        for t in (d1, d2, i1, i2):
            h.update(np.ascontiguousarray(t.cpu().numpy()[0]).tobytes())
        res["bit_exact_vs_reference_ext_golden"] = bool(h.hexdigest() == str(g["sha256"]))
        assert res["bit_exact_vs_reference_ext_golden"], "C1: differs from the reference extension's golden"
    return res


def c2_16k_metric(dev, Bk=16, n=16384):
    """The north_star's roofline target shape: batched Chamfer fwd+bwd at 16384 x 16384 points per pair (Bk pairs: 8.6e9 directed
    pairs per step), through Completionloss('cd_l2'): the library default (pruned exact scan) and, beside it, the exhaustive
    symmetric scan (GENPC_CHAMFER_PRUNE=0), each as a fraction of the nominal FP32 peak on ALGORITHMIC flops (8 per directed
    pair); L2 flushed before every timed step; the two paths' outputs compared bit for bit."""
    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist
    from genpc_b200.synthetic import pcn_batch
    from genpc_b200.utils.loss_util import Completionloss

    part, comp = pcn_batch(100, Bk, n, n)
    a = torch.from_numpy(part).to(dev).requires_grad_(True)
    b = torch.from_numpy(comp).to(dev).requires_grad_(True)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    crit = Completionloss("cd_l2")

    def step():
        a.grad = None
        b.grad = None
        crit.get_loss(a, b).backward()

    def timed(reps=10, warm=3):
        ts = []
        for r in range(reps + warm):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            if r >= warm:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    pairs = 2.0 * Bk * n * n
    out = {"workload": f"batched Chamfer fwd+bwd, B={Bk}, {n} x {n} pts, loss = chamfer_l2 (north_star roofline target shape)",
           "pairs_per_step": pairs}
    res = {}
    for name, knob in (("default_pruned_exact", None), ("exhaustive", "0")):
        with _lib.tunable(GENPC_CHAMFER_PRUNE=knob):
            ms = timed()
            res[name] = tuple(x.cpu().numpy() for x in chamfer_3DDist()(a.detach(), b.detach()))
        tf = pairs * FLOP_PER_PAIR / (ms * 1e-3) / 1e12
        out[name] = {"ms_fwd_bwd": ms, "pairs_per_s": pairs / (ms * 1e-3), "algorithmic_tflops": tf,
                     "frac_of_fp32_peak": tf / FP32_NOMINAL_TFLOPS}
    same = all(np.array_equal(x.view(np.int32), y.view(np.int32)) for x, y in zip(res["default_pruned_exact"], res["exhaustive"]))
    out["identical_outputs"] = bool(same)
    assert same, "16K x 16K: the pruned scan differs from the exhaustive scan"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1 / C4 / C5 legs (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, local, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)

    if not torch.cuda.is_available():
        if args.impl == "reference":
            # no GPU: the reference's CUDA extension cannot run; time the CPU port instead
            part, comp = make_inputs(0)
            cb = cpu_baseline(part, comp, budget_b=4)
            print(json.dumps({"impl": "reference", "metric": "chamfer_fwd_bwd_point_pairs_per_sec",
                              "value": cb["value"], "unit": "pairs/s", "n_gpus": 0, "cpu_baseline": cb,
                              "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0}}))
            return
        raise SystemExit("bench.py: no CUDA device (genpc_b200 has no CPU fallback)")

    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    part, comp = make_inputs(rank)
    hbm_peak, peak_src = peaks()

    line = {"metric": "chamfer_fwd_bwd_point_pairs_per_sec", "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: batched Chamfer3D fwd+bwd, B=32 per GPU, 2048 x 16384 pts (PCN shape), "
                                   "loss = chamfer_l2", "B_per_gpu": B, "N": N, "M": M,
                       "pairs_per_step_per_gpu": 2 * B * N * M, "parallelism": f"dp{world} (independent scans)",
                       "l2": "flushed (256 MiB write) before every timed step"}}

    if args.impl == "reference":
        import oracle

        ext = oracle.load_ref_ext("chamfer_3D")
        if ext is None:
            if rank == 0:
                cb = cpu_baseline(part, comp)
                line.update({"impl": "reference", "value": cb["value"], "cpu_baseline": cb, "ms_per_step": None,
                             "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                                     "d2h_bytes_per_step": 0},
                             "note": "oracle/_ref not built: timed the CPU port"})
                print(json.dumps(line))
            return
        api = import_reference_api()
        if api is not None:
            fn, how = api, ("UNMODIFIED reference API utils.loss_util.Completionloss('cd_l2').get_loss (baseline/_ref) "
                            "on its UNMODIFIED CUDA extension (oracle/_ref, sm_100a)")
        else:
            raw = make_ref_function(ext)
            fn, how = (lambda x, y: chamfer_l2_loss(raw, x, y)), ("UNMODIFIED reference chamfer_3D CUDA extension "
                                                                  "(oracle/_ref) through dist_chamfer_3D.py:26-64 and "
                                                                  "loss_util.py:31-33 restated (baseline/_ref not installed)")
        res, _ = run_gpu_arm(args, fn, rank, world, dev, part, comp, "reference")
        line.update(res)
        line.update({"impl": "reference", "gpu_launches": 4 * args.steps,
                     "note": how + "; stock flow = CPU-allocated outputs + .to(device) every call; the reference has "
                                   "no CPU implementation of this path (CPU numbers in cpu_baseline)"})
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(part, comp)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    from genpc_b200.utils.loss_util import Completionloss

    ours_loss = Completionloss("cd_l2")   # the reference's facade name and call: get_loss == chamfer_l2
    from genpc_b200.utils.loss_util import GraphedLossStep

    res, (a, b, flush) = run_gpu_arm(args, ours_loss.get_loss, rank, world, dev, part, comp, "ours",
                                     host_loss_fn=lambda ha, hb: ours_loss.get_loss_from_host(ha, hb, device=dev),
                                     graphed_factory=lambda ga, gb: GraphedLossStep(ours_loss, ga, gb))
    line.update(res)
    # per step: nn_bin_sort_kernel, nn_prune_coop_kernel (partial -> complete direction, side stream), nn_prune_kernel (the other
    # direction), nn_sym_kernel (device-deselected fall-back: returns at once), nn_sym_epilogue_kernel<fused> (unpack + loss +
    # zero-fill), chamfer_grad_kernel<.., LOSS>
    line["gpu_launches"] = 6 * args.steps
    line["api"] = "genpc_b200.utils.loss_util.Completionloss('cd_l2').get_loss(gen, gt); loss.backward()"
    # ---- the other BASELINE configs ride in the same line (outside the C2 timed region) ----
    cfgs = {}
    try:
        cfgs["C3_registration"] = registration_metric(rank, world, dev)
    except Exception as e:  # pragma: no cover
        cfgs["C3_registration"] = {"error": str(e)}
    line["registration"] = cfgs["C3_registration"]
    if not args.no_extras:
        try:
            cfgs["C5_sharded_chamfer"] = c5_sharded_metric(rank, world, dev)   # collective: every rank takes part
        except AssertionError:
            raise
        except Exception as e:  # pragma: no cover
            cfgs["C5_sharded_chamfer"] = {"error": str(e)}
    line["baseline_configs"] = cfgs
    if rank == 0:
        from genpc_b200 import _lib

        t_fwd, t_bwd = time_kernels_ours(dev, a, b, flush)                  # library default: the pruned exact scan on this shape
        with _lib.tunable(GENPC_CHAMFER_PRUNE="0"):
            t_exh, _ = time_kernels_ours(dev, a, b, flush, iters=10)        # the exhaustive symmetric scan, same inputs
        pst = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.lib().genpc_chamfer_prune_stats(_lib.ptr(pst))
        from genpc_b200.loss_functions import chamfer_3DDist
        chamfer_3DDist()(a.detach(), b.detach())
        torch.cuda.synchronize(dev)
        _lib.lib().genpc_chamfer_prune_stats(None)
        pst = pst.cpu().tolist()
        pruned = pst[2] > 0
        gb_total = B * (((N + 31) // 32) * ((M + 63) // 64) + ((M + 31) // 32) * ((N + 63) // 64))
        visited = pst[0] / gb_total if pruned else 1.0
        prof_exh = read_ncu_profile("nn_sym", "nn_sym")
        prof_scan = read_ncu_profile("nn_prune", "nn_prune_kernel") if pruned else prof_exh
        prof_grad = read_ncu_profile("fix_grad", "chamfer_grad_kernel") or read_ncu_profile("fix_grad", "chamfer_loss_grad_kernel")
        flops = 2.0 * B * N * M * FLOP_PER_PAIR
        ach = flops / (t_fwd * 1e-3) / 1e12
        m = measured_fp32_peak()
        peak_src_fp32 = ("nominal FFMA peak 148x128x2x1.965 GHz (MEASURED_PEAKS.json has no FP32 figure; measured issue rates in "
                         "profiles/fp32_peak_b200.json)")
        line["roofline"] = {"bound": "fp32",
                            "kernel": ("nn_bin_sort_kernel + nn_prune_kernel + nn_sym_epilogue_kernel (Hilbert-sorted, block-pruned EXACT "
                                       "scan; timed as one forward op)") if pruned else
                                      "nn_sym_kernel (+ nn_sym_epilogue_kernel, timed as one forward op)",
                            "algorithmic_flops_per_launch": flops,
                            "note": "achieved = ALGORITHMIC flops (8 per directed pair, 2*B*N*M directed pairs: what the reference's "
                                    "kernel computes) / forward time.  The default path on this shape skips, exactly, every 64-target "
                                    "block that cannot hold a nearest neighbour: it EVALUATES only `visited_fraction` of the pairs, so "
                                    "the algorithmic rate can exceed the FP32 peak; `evaluated` is the arithmetic it really does and "
                                    "`roofline_exhaustive` the kernel that evaluates every pair (the one earlier rounds reported)",
                            "achieved": ach,
                            "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_NOMINAL_TFLOPS,
                            "peak_source": peak_src_fp32,
                            "ms": t_fwd, "pairs_per_s": 2.0 * B * N * M / (t_fwd * 1e-3),
                            "visited_fraction": visited,
                            "evaluated": {"pairs_per_s": visited * B * N * M * 2.0 / (t_fwd * 1e-3),
                                          "tflops": visited * flops / (t_fwd * 1e-3) / 1e12,
                                          "frac_of_peak": visited * ach / FP32_NOMINAL_TFLOPS,
                                          "group_block_pairs_visited": pst[0], "of": gb_total},
                            "algorithmic_bytes_per_launch": 20.0 * B * (N + M),
                            "traffic": ncu_dram_bytes(prof_scan),
                            "traffic_source": ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of the scan "
                                               f"kernel, read at run time from {prof_scan['file']}") if prof_scan else None}
        ach_e = flops / (t_exh * 1e-3) / 1e12
        # the symmetric kernel EXECUTES each distance once (6 FMA-pipe lane-ops, counted as 8 flop), i.e. half of the algorithmic
        # work; ncu's FMA-pipe utilisation of the same launch is recorded beside it
        lane_ops = 6.0 * B * N * M / (t_exh * 1e-3)      # 3 sub + 1 mul + 2 fma per distance, each distance evaluated once
        line["roofline_exhaustive"] = {
            "bound": "fp32", "kernel": "nn_sym_kernel (+ nn_sym_epilogue_kernel), GENPC_CHAMFER_PRUNE=0: every pair evaluated",
            "achieved": ach_e, "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": ach_e / FP32_NOMINAL_TFLOPS,
            "peak_source": peak_src_fp32, "ms": t_exh, "pairs_per_s": 2.0 * B * N * M / (t_exh * 1e-3),
            "traffic": ncu_dram_bytes(prof_exh),
            "executed": {"fma_pipe_lane_ops_per_s": lane_ops, "frac_of_fma_pipe_lane_rate": lane_ops / (148 * 128 * 1.965e9),
                         "fma_pipe_cycles_active_pct_ncu":
                             prof_exh["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]
                             if prof_exh and "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active" in prof_exh else None,
                         "note": "the symmetric kernel evaluates each distance once for both directions (FMA-pipe ceiling of that "
                                 "formulation: 98.9 TFLOP/s algorithmic); ncu figure: the scan kernel alone, " +
                                 (prof_exh["file"] if prof_exh else "no profile")}}
        if m and "ffma2" in m:
            line["roofline"]["measured_ffma2_tflops"] = m["ffma2"].get("tflops")
        bwd_bytes = 44.0 * B * (N + M)
        gbs = bwd_bytes / (t_bwd * 1e-3) / 1e9
        line["roofline_bwd"] = {"bound": "hbm", "kernel": "chamfer_grad_kernel", "achieved": gbs, "peak": hbm_peak,
                                "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": peak_src, "ms": t_bwd,
                                "algorithmic_bytes_per_launch": bwd_bytes, "traffic": ncu_dram_bytes(prof_grad),
                                "traffic_source": (f"ncu --set full of the fused-loss instantiation of chamfer_grad_kernel, "
                                                   f"{prof_grad['file']}") if prof_grad else None}
        if world == 1:
            try:
                import oracle

                ext = oracle.load_ref_ext("chamfer_3D")
                if ext is not None:
                    d1 = torch.zeros(B, N, device=dev); d2 = torch.zeros(B, M, device=dev)
                    i1 = torch.zeros(B, N, dtype=torch.int32, device=dev)
                    i2 = torch.zeros(B, M, dtype=torch.int32, device=dev)
                    ts = []
                    for it in range(6):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); ext.forward(a.detach(), b.detach(), d1, d2, i1, i2); e1.record()
                        torch.cuda.synchronize(dev)
                        if it >= 2:
                            ts.append(e0.elapsed_time(e1))
                    t_ref = sum(ts) / len(ts)
                    line["ref_cuda_ext"] = {"fwd_ms": t_ref, "pairs_per_s": 2.0 * B * N * M / (t_ref * 1e-3),
                                            "speedup_fwd_kernel": t_ref / t_fwd}
            except Exception as e:  # pragma: no cover
                line["ref_cuda_ext"] = {"error": str(e)}
            if not args.no_cpu_baseline:
                cb = cpu_baseline(part, comp, keep_forward=True)
                # full-batch parity of the arm that was just timed: every one of the 32 scans, bit for bit, against the oracle
                exp = cb.pop("_oracle_forward")
                from genpc_b200.loss_functions import chamfer_3DDist

                got = chamfer_3DDist()(a.detach(), b.detach())
                ok = all(np.array_equal(g_.cpu().numpy().view(np.int32), e_.view(np.int32)) for g_, e_ in zip(got, exp))
                line["parity"] = {"c2_full_batch_bit_exact_vs_oracle": bool(ok), "scans": B}
                assert ok, "C2: the timed kernels differ from the oracle"
                line["cpu_baseline"] = cb
            if not args.no_extras:
                for key, fn in (("C1_real_scan", c1_metric), ("C2_16k_batched", c2_16k_metric), ("C4_depth_fps", c4_metric),
                                ("C5_emd", emd_c5_metric)):
                    try:
                        cfgs[key] = fn(dev)
                    except AssertionError:
                        raise
                    except Exception as e:  # pragma: no cover
                        cfgs[key] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
