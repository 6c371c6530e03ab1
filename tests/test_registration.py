"""Registration loop: oracle self-checks on CPU (autograd gradient vs finite differences, Adam bookkeeping);
on the GPU the fused step kernel against the oracle -- loss and pose gradient within 1e-5 relative, the Adam
trajectory over tens of iterations, multi-start selection, and pose recovery at the BASELINE size (16384 pts)."""
import math

import numpy as np
import pytest

import oracle
from oracle import registration as OR


def make_pair(seed, nc, nr):
    from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric

    comp = superquadric(seed, nc)
    part, gt = rigid_perturb(partial_view(comp, seed, nr), seed, max_rot_deg=20.0, max_t=0.05, scale_range=(0.7, 0.9))
    return comp, part, gt


def test_oracle_transform_matches_float64():
    comp, part, _ = make_pair(0, 500, 300)
    par = OR.init_params(1)
    par[6:9] = [0.01, -0.02, 0.03]
    pts = oracle.transform(comp, comp.mean(0), par)
    R = oracle.pose_matrix(par[:6]).astype(np.float64)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
    c = comp.mean(0).astype(np.float64)
    ref = (R @ ((comp - c) * math.exp(par[9])).T).T + c + par[6:9]
    assert np.abs(pts - ref).max() < 1e-6


def test_oracle_gradient_matches_finite_differences():
    comp, part, _ = make_pair(1, 800, 400)
    par = OR.init_params(0)
    par[:6] += np.float32(0.03) * np.arange(6, dtype=np.float32)
    par[6:9] = [0.02, 0.01, -0.01]
    c = comp.mean(0)
    loss, g, (iA, iB) = OR.loss_and_grad(par, comp, c, part)

    def f(p):  # same loss with the NN indices frozen, float64
        p = p.astype(np.float64)
        a1, a2 = p[:3], p[3:6]
        b1 = a1 / np.linalg.norm(a1)
        b2 = a2 - (b1 @ a2) * b1
        b2 /= np.linalg.norm(b2)
        R = np.stack([b1, b2, np.cross(b1, b2)])
        pts = (R @ ((comp.astype(np.float64) - c.astype(np.float64)) * math.exp(p[9])).T).T + c + p[6:9]
        return 3.0 * (np.sqrt(((pts - part[iA]) ** 2).sum(1)).mean() + 0.5 * np.sqrt(((part - pts[iB]) ** 2).sum(1)).mean())

    for i in range(10):
        e = np.zeros(10)
        e[i] = 1e-6
        fd = (f(par + e) - f(par - e)) / 2e-6
        assert abs(fd - g[i]) <= 1e-4 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_oracle_run_decreases_loss():
    comp, part, _ = make_pair(2, 600, 300)
    hist, losses = OR.run(comp, part, 15, lr=0.01)
    assert hist.shape == (16, 10) and losses[-1] < losses[0]


@pytest.mark.gpu
def test_step_loss_and_gradient_vs_oracle(cuda):
    import torch

    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch

    pairs = [make_pair(s, 3000, 1500) for s in range(3)]
    comp = torch.from_numpy(np.stack([p[0] for p in pairs])).to(cuda)
    part = torch.from_numpy(np.stack([p[1] for p in pairs])).to(cuda)
    rb = RegistrationBatch(comp, part, n_starts=4, lr=0.01)
    # perturb the initial parameters so the gradient is generic
    g = torch.Generator().manual_seed(0)
    rb.params += (torch.randn(rb.params.shape, generator=g) * 0.02).to(cuda)
    p0 = rb.params.cpu().numpy().copy()
    rb.run(1)
    torch.cuda.synchronize()
    m = rb.adam_m.cpu().numpy()
    loss = rb.losses().cpu().numpy()[:, 0]
    for s in range(rb.S):
        c = s // 4
        eloss, egrad, _ = OR.loss_and_grad(p0[s], pairs[c][0], rb.center[c].cpu().numpy(), pairs[c][1])
        assert abs(loss[s] - eloss) <= 1e-5 * abs(eloss)
        got = m[s] / np.float32(1.0 - 0.9)           # exp_avg after one step = (1-beta1) * grad
        assert np.abs(got - egrad).max() <= 1e-5 * np.abs(egrad).max() + 1e-7, (s, got, egrad)
    # first Adam step moves every parameter by ~lr_group * sign(grad)
    step = rb.params.cpu().numpy() - p0
    lr = np.array([0.01] * 6 + [0.002] * 3 + [0.001])
    assert (np.abs(step) <= lr * (1 + 1e-4)).all() and (np.abs(step) >= 0.5 * lr).mean() > 0.9


@pytest.mark.gpu
def test_trajectory_vs_oracle(cuda):
    import torch

    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch

    comp, part, _ = make_pair(5, 2048, 1024)
    iters = 25
    rb = RegistrationBatch(torch.from_numpy(comp[None]).to(cuda), torch.from_numpy(part[None]).to(cuda), n_starts=1, lr=0.01)
    hist = [rb.params.cpu().numpy()[0].copy()]
    for _ in range(iters):
        rb.run(1)
        hist.append(rb.params.cpu().numpy()[0].copy())
    ehist, elosses = OR.run(comp, part, iters, lr=0.01, center=rb.center[0].cpu().numpy())
    losses = rb.losses().cpu().numpy()[0]
    # tolerance (north_star): 1e-5 relative on every loss of the trajectory, every parameter within 1e-5 absolute
    # (r01 used 5e-5 / 5e-5; measured drift against the oracle is <= 3e-6 over 500 steps, profiles/r02b_c3_spread_500iters.json)
    assert (np.abs(losses - elosses) <= 1e-5 * np.abs(elosses)).all(), np.abs(losses - elosses).max()
    assert np.abs(np.stack(hist) - ehist).max() <= 1e-5, np.abs(np.stack(hist) - ehist).max()


@pytest.mark.gpu
def test_multistart_and_pose_recovery_16k(cuda):
    """BASELINE C3 shape (16384-pt clouds): the optimiser must pull the generated shape onto the partial scan."""
    import torch

    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch, object_pose_optimization_points

    comp, part, (Rg, tg, sg) = make_pair(7, 16384, 16384)
    rb = RegistrationBatch(torch.from_numpy(comp[None]).to(cuda), torch.from_numpy(part[None]).to(cuda), n_starts=4, lr=0.01,
                           max_iters=201)
    rb.run(201)
    torch.cuda.synchronize()
    L = rb.losses().cpu().numpy()
    assert np.isfinite(L).all() and (L[:, -1] < L[:, 0]).all()
    p, k, lo = rb.best()
    assert int(k[0]) == int(np.argmin(L.min(1)))
    assert abs(math.exp(float(p[0, 9])) - float(sg)) < 0.08       # recovered scale close to ground truth
    T = rb.transforms()[0].cpu().numpy()
    T2 = object_pose_optimization_points(comp, part, lr=0.01, iters=200)
    assert np.array_equal(T, T2)                                   # deterministic, identical through the mirror API


@pytest.mark.gpu
def test_trajectory_vs_reference_machinery_on_gpu(cuda):
    """The reference's own machinery on the same GPU: its UNMODIFIED Chamfer extension (oracle/_ref) inside torch
    autograd, the forward expression of ObjectPoseOptim.forward (diff_obj_pose.py:419-423), the Chamfer term of
    compute_loss_function (:326-327, weight 3.0 :334) and torch.optim.Adam with the three lr groups (:524-528), all
    fp32 -- against the fused kernels.  The reference's backward accumulates with unordered float atomics, so it is
    not bit-reproducible itself (its run-to-run spread after 30 steps is ~5e-10, profiles/r02b_c3_spread_500iters.json);
    tolerance: 1e-5 relative on every loss of the trajectory, 1e-5 absolute on the parameters after 30 steps."""
    import torch

    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch, rotation_6d_to_matrix

    ext = oracle.load_ref_ext("chamfer_3D")
    if ext is None:
        pytest.skip("oracle/_ref/chamfer_3D not built")

    class RefChamfer(torch.autograd.Function):          # dist_chamfer_3D.py:26-64 around the unmodified extension
        @staticmethod
        def forward(ctx, a, b):
            B, n, _ = a.shape
            m = b.shape[1]
            d1 = torch.zeros(B, n, device=a.device); d2 = torch.zeros(B, m, device=a.device)
            i1 = torch.zeros(B, n, dtype=torch.int32, device=a.device); i2 = torch.zeros(B, m, dtype=torch.int32, device=a.device)
            ext.forward(a, b, d1, d2, i1, i2)
            ctx.save_for_backward(a, b, i1, i2)
            return d1, d2, i1, i2

        @staticmethod
        def backward(ctx, g1, g2, _a, _b):
            a, b, i1, i2 = ctx.saved_tensors
            ga, gb = torch.zeros_like(a), torch.zeros_like(b)
            ext.backward(a, b, ga, gb, g1.contiguous(), g2.contiguous(), i1, i2)
            return ga, gb

    def partial_l1(p, q):                                # Completionloss.chamfer_partial_l1 (loss_util.py:35-38)
        d1, _, _, _ = RefChamfer.apply(p.contiguous(), q.contiguous())
        return torch.sqrt(d1).mean()

    comp, part, _ = make_pair(11, 4096, 3000)
    V, Rf = torch.from_numpy(comp).to(cuda), torch.from_numpy(part).to(cuda)
    iters, lr = 30, 0.01
    rb = RegistrationBatch(V[None], Rf[None], n_starts=1, lr=lr)
    center = rb.center[0].clone()
    p0 = rb.params[0].clone()
    rot = p0[:6].clone().requires_grad_(True); trans = p0[6:9].clone().requires_grad_(True); ls = p0[9:].clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [rot], "lr": lr}, {"params": [trans], "lr": lr * 0.2}, {"params": [ls], "lr": lr * 0.1}])
    ref_losses = []
    for _ in range(iters):
        opt.zero_grad()
        R = rotation_6d_to_matrix(rot[None])[0]
        local = (V - center) * torch.exp(ls)[0]
        pts = (R @ local.T).T + center + trans
        loss = 3.0 * (partial_l1(pts[None], Rf[None]) + 0.5 * partial_l1(Rf[None], pts[None]))
        loss.backward()
        opt.step()
        ref_losses.append(float(loss.detach()))
    rb.run(iters)
    ours = rb.losses()[0].cpu().numpy()
    ref_losses = np.array(ref_losses)
    assert (np.abs(ours - ref_losses) <= 1e-5 * np.abs(ref_losses)).all(), np.abs(ours - ref_losses).max()
    ref_p = torch.cat([rot, trans, ls]).detach().cpu().numpy()
    assert np.abs(rb.params[0].cpu().numpy() - ref_p).max() <= 1e-5, np.abs(rb.params[0].cpu().numpy() - ref_p).max()


@pytest.mark.gpu
def test_c3_length_500_iterations_vs_oracle_and_reference_self_spread(cuda):
    """BASELINE config C3 runs 500 Adam iterations; north_star: registered pose / scale within 1e-5.  Three arms on the same
    pair: the fused kernels, the oracle (C-oracle NN indices + float64 autograd + torch Adam, oracle/registration.py) and the
    reference's machinery run TWICE (its backward sums with unordered float atomics).  tools/c3_spread.py recorded the full
    picture (profiles/r02b_c3_spread_500iters.json): on this pair the kernels stay within 3e-6 of the oracle for all 500
    steps while the reference drifts up to 3.6e-4 from ITSELF mid-trajectory; on a second pair (4096 x 3000) every arm --
    reference vs reference included -- bifurcates around iteration 200 (one flipped nearest neighbour, amplified by Adam's
    normalisation) and ends 1e-3..5e-3 apart in the parameters with losses still equal to 2e-6.
    Tolerances: against the oracle -- every loss 1e-5 relative, every parameter at EVERY step within max(1e-5, 2 x the
    reference's own run-to-run spread), the final pose / scale within 1e-5; against the reference's two runs -- the nearer
    one within 1e-3 at every step and 1e-4 at the end (a loose bound: its own twin runs can sit 3.6e-4 apart)."""
    import torch

    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch, rotation_6d_to_matrix

    iters, lr = 500, 0.01
    comp, part, _ = make_pair(11, 2048, 1500)
    V, Rf = torch.from_numpy(comp).to(cuda), torch.from_numpy(part).to(cuda)
    rb = RegistrationBatch(V[None], Rf[None], n_starts=1, lr=lr, max_iters=iters)
    center, p0 = rb.center[0].clone(), rb.params[0].clone()
    ours = []
    for _ in range(iters):
        rb.run(1)
        ours.append(rb.params[0].cpu().numpy().copy())
    ours = np.stack(ours)
    ours_l = rb.losses()[0].cpu().numpy()
    eh, el = OR.run(comp, part, iters, lr=lr, center=center.cpu().numpy())
    eh = eh[1:]
    assert abs(ours_l[-1] - el[-1]) <= 1e-5 * abs(el[-1])
    assert (np.abs(ours_l - el) <= 1e-5 * np.abs(el)).all()
    spread = 0.0
    ext = oracle.load_ref_ext("chamfer_3D")
    if ext is not None:
        class RefChamfer(torch.autograd.Function):
            @staticmethod
            def forward(ctx, a, b):
                B, n, _ = a.shape
                m = b.shape[1]
                d1 = torch.zeros(B, n, device=a.device); d2 = torch.zeros(B, m, device=a.device)
                i1 = torch.zeros(B, n, dtype=torch.int32, device=a.device); i2 = torch.zeros(B, m, dtype=torch.int32, device=a.device)
                ext.forward(a, b, d1, d2, i1, i2)
                ctx.save_for_backward(a, b, i1, i2)
                return d1, d2, i1, i2

            @staticmethod
            def backward(ctx, g1, g2, _a, _b):
                a, b, i1, i2 = ctx.saved_tensors
                ga, gb = torch.zeros_like(a), torch.zeros_like(b)
                ext.backward(a, b, ga, gb, g1.contiguous(), g2.contiguous(), i1, i2)
                return ga, gb

        def run_ref():
            rot = p0[:6].clone().requires_grad_(True); trans = p0[6:9].clone().requires_grad_(True); ls = p0[9:].clone().requires_grad_(True)
            opt = torch.optim.Adam([{"params": [rot], "lr": lr}, {"params": [trans], "lr": lr * 0.2}, {"params": [ls], "lr": lr * 0.1}])
            hist = []
            for _ in range(iters):
                opt.zero_grad()
                R = rotation_6d_to_matrix(rot[None])[0]
                pts = (R @ ((V - center) * torch.exp(ls)[0]).T).T + center + trans
                pl1 = lambda p, q: torch.sqrt(RefChamfer.apply(p.contiguous(), q.contiguous())[0]).mean()   # noqa: E731
                loss = 3.0 * (pl1(pts[None], Rf[None]) + 0.5 * pl1(Rf[None], pts[None]))
                loss.backward()
                opt.step()
                hist.append(torch.cat([rot, trans, ls]).detach().cpu().numpy().copy())
            return np.stack(hist)

        r1, r2 = run_ref(), run_ref()
        spread = float(np.abs(r1 - r2).max())
        # against the reference itself only a loose bound can be asserted without flakiness: one of its runs may take another
        # branch at a near-tie for a few dozen steps (3.6e-4 from its twin in the recorded session) -- two samples do not
        # bound that.  The strict 1e-5 checks are the ones against the deterministic oracle below.
        near = min(np.abs(ours - r1).max(), np.abs(ours - r2).max())
        assert near <= max(1e-3, 2.0 * spread), (near, spread)
        assert min(np.abs(ours[-1] - r1[-1]).max(), np.abs(ours[-1] - r2[-1]).max()) <= max(1e-4, 2.0 * spread)
    tol = max(1e-5, 2.0 * spread)
    assert np.abs(ours - eh).max() <= tol, (np.abs(ours - eh).max(), spread)
    assert np.abs(ours[-1] - eh[-1]).max() <= 1e-5, np.abs(ours[-1] - eh[-1]).max()     # the registered pose / scale itself


@pytest.mark.gpu
@pytest.mark.parametrize("nc,nr,starts", [(2500, 1000, 4), (900, 1400, 2), (3000, 3000, 4)])
def test_persistent_small_registration_equals_launch_per_iteration(cuda, nc, nr, starts):
    """Small clouds (the real pipeline's 1-3 K points, 4 starts): run(n) is ONE cooperative launch that iterates on the device
    (register_persistent_kernel); it must reproduce the launch-per-iteration path -- same kernels' arithmetic, Adam scalars
    computed on the device instead of the host -- to float rounding of those scalars (1e-6), and be deterministic."""
    import torch

    from genpc_b200 import _lib
    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch

    comp, part, _ = make_pair(21, nc, nr)
    V, Rf = torch.from_numpy(comp[None]).to(cuda), torch.from_numpy(part[None]).to(cuda)
    iters = 60
    res = []
    for persist in ("1", "1", "0"):
        with _lib.tunable(GENPC_REGISTER_PERSIST=persist):
            rb = RegistrationBatch(V, Rf, n_starts=starts, lr=0.01, max_iters=iters)
            rb.run(iters)
            torch.cuda.synchronize()
        res.append((rb.params.cpu().numpy().copy(), rb.losses().cpu().numpy().copy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])      # deterministic
    assert np.abs(res[0][0] - res[2][0]).max() <= 1e-6, np.abs(res[0][0] - res[2][0]).max()
    assert np.abs(res[0][1] - res[2][1]).max() <= 1e-6 * np.abs(res[2][1]).max()
    assert (res[0][1][:, -1] < res[0][1][:, 0]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("C,nc,nr,starts,iters", [(2, 6000, 4096, 2, 25), (3, 16384, 16384, 1, 12), (1, 2049, 5000, 4, 30)])
def test_pruned_registration_equals_exhaustive_registration(cuda, C, nc, nr, starts, iters):
    """The pruned scan inside the registration loop (fixed clouds Hilbert-sorted once, moving clouds every iteration in their
    current pose) yields the same distances and the same lowest-index neighbours as the exhaustive symmetric scan, so params and
    loss history are IDENTICAL bit for bit, iteration after iteration, multi-start included; run() split in two calls too."""
    import torch

    from genpc_b200 import _lib
    from genpc_b200.optim_registration.diff_obj_pose import RegistrationBatch

    pairs = [make_pair(40 + c, nc, nr) for c in range(C)]
    V = torch.from_numpy(np.stack([p[0] for p in pairs])).to(cuda)
    Rf = torch.from_numpy(np.stack([p[1] for p in pairs])).to(cuda)
    res = {}
    for knob in ("0", "1"):
        with _lib.tunable(GENPC_REGISTER_PRUNE=knob, GENPC_REGISTER_MODE="sym"):
            rb = RegistrationBatch(V, Rf, n_starts=starts, lr=0.01, max_iters=iters)
            rb.run(iters // 2)
            rb.run(iters - iters // 2)
            torch.cuda.synchronize()
        res[knob] = (rb.params.cpu().numpy().copy(), rb.losses().cpu().numpy().copy())
    assert np.array_equal(res["0"][1], res["1"][1]), np.abs(res["0"][1] - res["1"][1]).max()
    assert np.array_equal(res["0"][0], res["1"][0])
    assert (res["1"][1][:, -1] < res["1"][1][:, 0]).all()
