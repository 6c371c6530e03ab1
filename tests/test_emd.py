"""EMD auction: the CUDA kernel bit-exact (dist + assignment) against the oracle restatement -- which itself is
pinned on the reference's own outputs (test_oracle_golden.py) -- against the committed golden vectors, and
against the unmodified reference extension on the same GPU; plus the reference's own `Verified EMD`
self-consistency check (emd_module.py:112-118) and the shape errors of emd_cuda.cu:236-249."""
import glob
import os

import numpy as np
import pytest

import oracle

G = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_emd_verified_identity_and_bijection_trend():
    rng = np.random.default_rng(0)
    x1, x2 = rng.random((2, 512, 3), dtype=np.float32), rng.random((2, 512, 3), dtype=np.float32)
    d, a = oracle.emd_forward(x1, x2, 0.05, 3000)           # test_emd() settings (emd_module.py:98-103)
    assert all(len(np.unique(a[b])) == 512 for b in range(2))   # converged: a bijection
    x2g = np.take_along_axis(x2, a[..., None].astype(np.int64), axis=1)
    ver = ((x1.astype(np.float64) - x2g) ** 2).sum(-1)
    assert np.allclose(d, ver, rtol=1e-5, atol=1e-9)          # "Verified EMD" (emd_module.py:112-118)
    g = oracle.emd_backward(x1, x2, np.ones((2, 512), np.float32), a)
    assert np.allclose(g, 2 * (x1 - x2g), rtol=1e-6, atol=1e-7)


def test_oracle_emd_shape_errors():
    x = np.zeros((1, 300, 3), np.float32)
    with pytest.raises(ValueError):
        oracle.emd_forward(x, x, 0.005, 5)                      # n % 256 != 0


def run_ours(x1, x2, eps, iters, dev):
    import torch

    from genpc_b200.loss_functions import emdModule

    d, a = emdModule()(torch.from_numpy(x1).to(dev), torch.from_numpy(x2).to(dev), eps, iters)
    return d.cpu().numpy(), a.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,eps,iters", [(1, 256, 0.005, 50), (2, 512, 0.005, 50), (3, 1024, 0.002, 80),
                                           (1, 2304, 0.005, 50), (32, 1024, 0.005, 50), (1, 8192, 0.005, 50), (1, 16384, 0.005, 50),
                                           (2, 2048, 0.05, 400)])
def test_emd_gpu_bit_exact_vs_oracle(cuda, B, n, eps, iters):
    from genpc_b200 import _lib

    rng = np.random.default_rng(n + B)
    x1, x2 = rng.random((B, n, 3), dtype=np.float32), rng.random((B, n, 3), dtype=np.float32)
    ed, ea = oracle.emd_forward(x1, x2, eps, iters)
    # default (block-pruned Bid from n = 1024 up), pruned forced, exhaustive scan forced: one answer
    for knob in (None, "1", "0"):
        with _lib.tunable(GENPC_EMD_PRUNE=knob):
            d, a = run_ours(x1, x2, eps, iters, cuda)
        assert np.array_equal(a, ea), f"prune={knob}: {(a != ea).sum()} assignments differ"
        assert np.array_equal(d.view(np.int32), ed.view(np.int32)), f"prune={knob}"


def _adversarial_emd_inputs():
    rng = np.random.default_rng(11)
    g = np.stack(np.meshgrid(*[np.arange(16, dtype=np.float32) / 16] * 3, indexing="ij"), -1).reshape(-1, 3)     # 4096 lattice points
    yield "lattice_ties", g[rng.permutation(4096)][None], g[rng.permutation(4096)][None], 0.005, 60
    x = rng.random((2, 2048, 3), dtype=np.float32)
    yield "identical_clouds", x, x.copy(), 0.005, 50
    d = rng.random((1, 1024, 3), dtype=np.float32)
    yield "duplicated_targets", rng.random((1, 2048, 3), dtype=np.float32), np.concatenate([d, d], 1), 0.005, 50
    c = (rng.random((1, 4096, 3), dtype=np.float32) * 1e-3 + 0.5).astype(np.float32)
    yield "tiny_extent", c, (rng.random((1, 4096, 3), dtype=np.float32) * 1e-3 + 0.5).astype(np.float32), 0.005, 50
    yield "two_far_clusters", np.concatenate([rng.random((1, 1024, 3)), rng.random((1, 1024, 3)) + 40], 1).astype(np.float32), \
        np.concatenate([rng.random((1, 512, 3)), rng.random((1, 1536, 3)) + 40], 1).astype(np.float32), 0.01, 100
    yield "offset_scene", (rng.random((1, 2048, 3)) * 3 - 100).astype(np.float32), (rng.random((1, 2048, 3)) * 3 - 100).astype(np.float32), 0.005, 50
    yield "planar", np.concatenate([rng.random((1, 2048, 2)), np.zeros((1, 2048, 1))], 2).astype(np.float32), \
        np.concatenate([rng.random((1, 2048, 2)), np.zeros((1, 2048, 1))], 2).astype(np.float32), 0.005, 50
    yield "five_blocks", rng.random((3, 320 * 4, 3), dtype=np.float32)[:, :1280], rng.random((3, 1280, 3), dtype=np.float32), 0.005, 50
    yield "largest_pruned_n", rng.random((1, 32768, 3), dtype=np.float32), rng.random((1, 32768, 3), dtype=np.float32), 0.005, 20
    yield "more_clouds_than_ctas_per_group", rng.random((40, 1024, 3), dtype=np.float32), rng.random((40, 1024, 3), dtype=np.float32), 0.005, 50
    yield "large_eps", rng.random((1, 2048, 3), dtype=np.float32), rng.random((1, 2048, 3), dtype=np.float32), 0.5, 30


@pytest.mark.gpu
def test_emd_pruned_bid_equals_exhaustive_bid_on_adversarial_inputs(cuda):
    """The block-pruned Bid (Morton-sorted targets, box test per 64-target block) only skips targets the per-target filter
    rejects, so it must reproduce the exhaustive scan bit for bit on any input: exact ties (lattice, duplicates, identical
    clouds), degenerate boxes (planar, tiny extent), loose bounds (far clusters, large eps), shapes (5 blocks, n = 32768,
    B = 40).  The small cases are also checked against the oracle."""
    from genpc_b200 import _lib

    for name, x1, x2, eps, iters in _adversarial_emd_inputs():
        x1, x2 = np.ascontiguousarray(x1, np.float32), np.ascontiguousarray(x2, np.float32)
        with _lib.tunable(GENPC_EMD_PRUNE="1", GENPC_EMD_SORT="bitonic" if name in ("lattice_ties", "planar", "largest_pruned_n") else None):
            d1, a1 = run_ours(x1, x2, eps, iters, cuda)   # (both target sorts: counting sort by default, bitonic on request)
        with _lib.tunable(GENPC_EMD_PRUNE="0"):
            d0, a0 = run_ours(x1, x2, eps, iters, cuda)
        assert np.array_equal(a1, a0), f"{name}: {(a1 != a0).sum()} assignments differ"
        assert np.array_equal(d1.view(np.int32), d0.view(np.int32)), name
        if x1.shape[0] * x1.shape[1] <= 4096:
            ed, ea = oracle.emd_forward(x1, x2, eps, iters)
            assert np.array_equal(a1, ea) and np.array_equal(d1.view(np.int32), ed.view(np.int32)), name


@pytest.mark.gpu
def test_emd_gpu_golden_and_centered_data(cuda):
    for f in sorted(glob.glob(os.path.join(G, "emd_ref_*.npz"))):
        z = np.load(f)
        # emd_ref_centered.npz is the witness of the reference's GetMax race (tests/test_oracle_golden.py): the reference
        # landed on the "lowest bidder index" outcome there, which the kernel produces on request
        lowest = os.path.basename(f) == "emd_ref_centered.npz"
        from genpc_b200 import _lib

        with _lib.tunable(GENPC_EMD_GETMAX="lowest" if lowest else None):
            d, a = run_ours(z["xyz1"], z["xyz2"], float(z["eps"]), int(z["iters"]), cuda)
        assert np.array_equal(a, z["assignment"]) and np.array_equal(d.view(np.int32), z["dist"].view(np.int32)), f
        ed, ea = oracle.emd_forward(z["xyz1"], z["xyz2"], float(z["eps"]), int(z["iters"]), getmax_lowest=lowest)
        assert np.array_equal(a, ea) and np.array_equal(d.view(np.int32), ed.view(np.int32)), f
    # the metric path feeds [-0.5, 0.5] data un-normalised (main.py:26-33)
    rng = np.random.default_rng(3)
    x1 = rng.random((1, 2048, 3), dtype=np.float32) - 0.5
    x2 = rng.random((1, 2048, 3), dtype=np.float32) - 0.5
    d, a = run_ours(x1, x2, 0.005, 50, cuda)
    ed, ea = oracle.emd_forward(x1, x2, 0.005, 50)
    assert np.array_equal(a, ea) and np.array_equal(d, ed)


@pytest.mark.gpu
def test_emd_gpu_vs_reference_extension_and_backward(cuda):
    import torch

    from genpc_b200.loss_functions import emdModule

    ref = oracle.load_ref_ext("emd")
    rng = np.random.default_rng(9)
    B, n = 4, 4096
    x1 = torch.from_numpy(rng.random((B, n, 3), dtype=np.float32)).to(cuda).requires_grad_(True)
    x2 = torch.from_numpy(rng.random((B, n, 3), dtype=np.float32)).to(cuda)
    d, a = emdModule()(x1, x2, 0.005, 50)
    torch.sqrt(d).mean(1).mean().backward()                     # Completionloss.emd_loss reduction
    eg = oracle.emd_backward(x1.detach().cpu().numpy(), x2.cpu().numpy(),
                             (0.5 / torch.sqrt(d) / (B * n)).detach().cpu().numpy(), a.cpu().numpy())
    assert np.abs(x1.grad.cpu().numpy() - eg).max() <= 1e-5 * np.abs(eg).max()
    if ref is None:
        pytest.skip("oracle/_ref/emd not built")
    dev = cuda
    dist = torch.zeros(B, n, device=dev); asg = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1
    asg_inv = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1; price = torch.zeros(B, n, device=dev)
    bid = torch.zeros(B, n, device=dev, dtype=torch.int32); binc = torch.zeros(B, n, device=dev)
    minc = torch.zeros(B, n, device=dev); uidx = torch.zeros(B * n, device=dev, dtype=torch.int32)
    midx = torch.zeros(B * n, device=dev, dtype=torch.int32)
    z512 = [torch.zeros(512, dtype=torch.int32, device=dev) for _ in range(3)]
    ref.forward(x1.detach(), x2, dist, asg, price, asg_inv, bid, binc, minc, uidx, z512[0], z512[1], z512[2], midx, 0.005, 50)
    torch.cuda.synchronize()
    if torch.equal(asg, a) and torch.equal(dist, d.detach()):
        return
    # The reference's GetMax is a store race (emd_cuda.cu:188-191): where a decisive collision
    # occurs it lands on the "highest index" or the "lowest index" outcome depending on block timing, or on a mixture.
    # Accept the other pure outcome bit for bit, or a mixture that differs in few assignments and not in cost.
    from genpc_b200 import _lib

    with _lib.tunable(GENPC_EMD_GETMAX="lowest"):
        d2, a2 = emdModule()(x1.detach(), x2, 0.005, 50)
    if torch.equal(asg, a2) and torch.equal(dist, d2):
        return
    frac = float((asg != a).float().mean())
    cost_ref, cost = float(torch.sqrt(dist).mean()), float(torch.sqrt(d.detach()).mean())
    assert frac < 0.05 and abs(cost - cost_ref) <= 1e-3 * cost_ref, (frac, cost, cost_ref)


@pytest.mark.gpu
def test_emd_shape_errors_and_completionloss(cuda):
    import torch

    from genpc_b200 import _lib, emd
    from genpc_b200.utils.loss_util import Completionloss

    x = torch.rand(1, 300, 3, device=cuda)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=cuda, dtype=dt)
    with pytest.raises(_lib.GenpcError):
        emd.forward(x, x, z(1, 300), z(1, 300, dt=torch.int32), z(1, 300), z(1, 300, dt=torch.int32),
                    z(1, 300, dt=torch.int32), z(1, 300), z(1, 300), z(300, dt=torch.int32), z(512, dt=torch.int32),
                    z(512, dt=torch.int32), z(512, dt=torch.int32), z(300, dt=torch.int32), 0.005, 5)
    y = torch.rand(1, 256, 3, device=cuda)
    for eps, iters in ((0.0, 5), (-0.01, 5), (0.005, 0), (float("nan"), 5)):   # rejected, not mis-ordered / OOB
        with pytest.raises(_lib.GenpcError):
            emd.forward(y, y, z(1, 256), z(1, 256, dt=torch.int32) - 1, z(1, 256), z(1, 256, dt=torch.int32) - 1,
                        z(1, 256, dt=torch.int32), z(1, 256), z(1, 256), z(256, dt=torch.int32), z(512, dt=torch.int32),
                        z(512, dt=torch.int32), z(512, dt=torch.int32), z(256, dt=torch.int32), eps, iters)
    g = torch.Generator().manual_seed(0)
    p1, p2 = torch.rand(2, 1024, 3, generator=g).to(cuda), torch.rand(2, 1024, 3, generator=g).to(cuda)
    for name in ("cd_l1", "cd_l2", "emd"):
        v = Completionloss(name).get_loss(p1, p2)
        assert v.ndim == 0 and torch.isfinite(v)
    cl = Completionloss("cd_l1")
    d1, d2, _, _ = cl.chamfer_dist(p1, p2)
    assert torch.allclose(cl.chamfer_partial_l1(p1, p2), torch.sqrt(d1).mean())
    assert torch.allclose(cl.chamfer_l1(p1, p2), (torch.sqrt(d1).mean() + torch.sqrt(d2).mean()) / 2)


@pytest.mark.gpu
def test_emd_unaligned_pointers(cuda):
    """xyz2 / price / assignment / max_idx that are only 4-byte aligned (the direct Bid path reads with LDG.64, the
    compaction with LDG.128 / STG.128: both must step aside)."""
    rng = np.random.default_rng(77)
    B, n = 1, 2048
    x1, x2 = rng.random((B, n, 3), dtype=np.float32), rng.random((B, n, 3), dtype=np.float32)
    import torch

    from genpc_b200 import emd as E

    def off(t):
        flat = torch.empty(t.numel() + 5, dtype=t.dtype, device=t.device)
        v = flat[1:1 + t.numel()].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 8 != 0
        return v

    t1, t2 = off(torch.from_numpy(x1).to(cuda)), off(torch.from_numpy(x2).to(cuda))
    dist = torch.zeros(B, n, device=cuda); asg = off(torch.full((B, n), -1, dtype=torch.int32, device=cuda))
    asg_inv = torch.full((B, n), -1, dtype=torch.int32, device=cuda); price = off(torch.zeros(B, n, device=cuda))
    bid = torch.zeros(B, n, dtype=torch.int32, device=cuda); binc = torch.zeros(B, n, device=cuda); minc = torch.zeros(B, n, device=cuda)
    uidx = torch.zeros(B * n, dtype=torch.int32, device=cuda); midx = off(torch.zeros(B * n, dtype=torch.int32, device=cuda))
    z = [torch.zeros(512, dtype=torch.int32, device=cuda) for _ in range(3)]
    E.forward(t1, t2, dist, asg, price, asg_inv, bid, binc, minc, uidx, z[0], z[1], z[2], midx, 0.005, 50)
    ed, ea = oracle.emd_forward(x1, x2, 0.005, 50)
    assert np.array_equal(asg.cpu().numpy(), ea) and np.array_equal(dist.cpu().numpy().view(np.int32), ed.view(np.int32))


@pytest.mark.gpu
def test_emd_c5_batch32_n8192_against_the_racy_reference(cuda):
    """BASELINE C5's EMD leg at full size (B=32, n=8192, eps 0.005, 50 iterations, torch.rand inputs as `test_emd`,
    emd_module.py:98-118).  At this size the reference's GetMax store race (emd_cuda.cu:188-191) is decided differently in
    different clouds of the batch, so its output equals NEITHER pure resolution of the race ("highest" / "lowest" bidder index)
    -- r01 found 0.0362958 for the reference against 0.0362955 / 0.0362985 for the two resolutions.  What can be asserted, and
    is: ours == the oracle bit for bit in both resolutions (sampled clouds), every cloud fully assigned, the matching cost within
    2e-4 relative of the reference's and the assignments equal for more than 90 % of the points.  north_star's 1e-5 on the EMD
    cost holds wherever the reference is race-free (goldens, B=1 sizes); here the reference's own run-to-run spread is printed."""
    import torch

    from genpc_b200 import _lib
    from genpc_b200.loss_functions import emdModule

    B, n = 32, 8192
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.rand(B, n, 3, generator=g).to(cuda), torch.rand(B, n, 3, generator=g).to(cuda)
    d_hi, a_hi = emdModule()(x1, x2, 0.005, 50)
    with _lib.tunable(GENPC_EMD_GETMAX="lowest"):
        d_lo, a_lo = emdModule()(x1, x2, 0.005, 50)
    for b in (0, 17):                                       # the CPU oracle takes ~2 s per cloud at this size
        ed, ea = oracle.emd_forward(x1[b:b + 1].cpu().numpy(), x2[b:b + 1].cpu().numpy(), 0.005, 50)
        assert np.array_equal(a_hi[b].cpu().numpy(), ea[0]) and np.array_equal(d_hi[b].cpu().numpy().view(np.int32), ed[0].view(np.int32))
    ed, ea = oracle.emd_forward(x1[5:6].cpu().numpy(), x2[5:6].cpu().numpy(), 0.005, 50, getmax_lowest=True)
    assert np.array_equal(a_lo[5].cpu().numpy(), ea[0])
    assert int((a_hi < 0).sum()) == 0 and int((a_lo < 0).sum()) == 0
    ref = oracle.load_ref_ext("emd")
    if ref is None:
        pytest.skip("oracle/_ref/emd not built")
    costs = []
    for rep in range(2):
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=cuda)   # noqa: E731
        dist, asg = z(B, n), z(B, n, dt=torch.int32) - 1
        ref.forward(x1, x2, dist, asg, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32), z(B, n), z(B, n),
                    z(B * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
                    z(B * n, dt=torch.int32), 0.005, 50)
        torch.cuda.synchronize()
        costs.append(float(torch.sqrt(dist).mean()))
    c_hi, c_lo = float(torch.sqrt(d_hi).mean()), float(torch.sqrt(d_lo).mean())
    print(f"EMD B=32 n=8192 cost: reference runs {costs}, ours highest {c_hi}, lowest {c_lo}; "
          f"assignments equal to the reference: {float((asg == a_hi).float().mean()):.4f} / {float((asg == a_lo).float().mean()):.4f}")
    assert min(abs(c_hi - costs[-1]), abs(c_lo - costs[-1])) <= 2e-4 * costs[-1]
    assert max(float((asg == a_hi).float().mean()), float((asg == a_lo).float().mean())) > 0.9
