"""GPU parity: libgenpc_b200's Chamfer forward/backward (through the C ABI via the reference-shaped Python
API) against the CPU oracle, against the UNMODIFIED reference extension (oracle/_ref, when built) and
against the committed golden vectors.  Bit-exact dist/idx; gradients within 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

import oracle
from util import lattice_cloud, rand_cloud, shape_cloud

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 1), (1, 5, 3), (2, 37, 513), (3, 600, 64), (1, 1025, 1023), (2, 2048, 16384),
          (1, 255, 4097), (4, 1000, 1000), (1, 16384, 16384), (1, 3, 70000),
          (1, 71372, 16384)]  # the last one is BASELINE config C1's shape (data/01184.ply has 71 372 points)


def run_ours(a, b, dev):
    from genpc_b200.loss_functions import chamfer_3DDist

    d1, d2, i1, i2 = chamfer_3DDist()(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
    return d1.cpu().numpy(), d2.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy()


@pytest.mark.parametrize("B,N,M", SHAPES)
def test_forward_bit_exact_vs_oracle(cuda, B, N, M):
    a, b = rand_cloud(B * 7 + N, B, N), rand_cloud(M + 1, B, M)
    got = run_ours(a, b, cuda)
    exp = oracle.chamfer_forward(a, b)
    for g, e, name in zip(got, exp, ("dist1", "dist2", "idx1", "idx2")):
        assert g.dtype == e.dtype, name
        assert np.array_equal(g.view(np.int32), e.view(np.int32)), f"{name} differs ({(g != e).sum()} entries)"


@pytest.mark.parametrize("B,N,M", [(2, 300, 2000), (1, 5000, 5000), (3, 1024, 1024)])
def test_ties_lowest_index(cuda, B, N, M):
    a, b = lattice_cloud(1, B, N), lattice_cloud(2, B, M)
    got = run_ours(a, b, cuda)
    exp = oracle.chamfer_forward(a, b)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)


def test_shape_clouds_and_negative_coords(cuda):
    a, b = shape_cloud(3, 2, 2048), shape_cloud(4, 2, 16384)
    got = run_ours(a, b, cuda)
    exp = oracle.chamfer_forward(a, b)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)


def test_self_distance(cuda):
    a = rand_cloud(5, 2, 3000)
    d1, d2, i1, i2 = run_ours(a, a, cuda)
    assert (d1 == 0).all() and (d2 == 0).all()
    assert (i1 == np.arange(3000)).all() and (i2 == np.arange(3000)).all()


def test_empty_clouds(cuda):
    from genpc_b200.loss_functions import chamfer_3DDist

    d1, d2, i1, i2 = chamfer_3DDist()(torch.zeros(2, 0, 3, device=cuda), torch.rand(2, 5, 3, device=cuda))
    assert d1.shape == (2, 0) and d2.shape == (2, 5) and (d2 == 0).all() and (i2 == 0).all()


@pytest.mark.parametrize("B,N,M", [(2, 90, 150), (4, 2048, 16384), (1, 7000, 300)])
def test_backward_vs_oracle(cuda, B, N, M):
    from genpc_b200.loss_functions import chamfer_3DDist

    a, b = shape_cloud(6, B, N), shape_cloud(7, B, M)
    rng = np.random.default_rng(0)
    g1 = rng.standard_normal((B, N)).astype(np.float32)
    g2 = rng.standard_normal((B, M)).astype(np.float32)
    ta = torch.from_numpy(a).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(b).to(cuda).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(ta, tb)
    (d1 * torch.from_numpy(g1).to(cuda)).sum().add((d2 * torch.from_numpy(g2).to(cuda)).sum()).backward()
    e1, e2 = oracle.chamfer_backward(a, b, g1, g2, i1.cpu().numpy(), i2.cpu().numpy())
    # tolerance: 1e-5 relative (north_star) on the gradient scale of each cloud
    for got, exp in ((ta.grad.cpu().numpy(), e1), (tb.grad.cpu().numpy(), e2)):
        scale = np.abs(exp).max()
        assert np.abs(got - exp).max() <= 1e-5 * scale + 1e-12


def test_full_size_properties_c2(cuda):
    """BASELINE config C2 (B=32, 2048 x 16384): size-independent properties instead of the oracle."""
    from genpc_b200.loss_functions import chamfer_3DDist

    g = torch.Generator(device="cpu").manual_seed(0)
    a = torch.rand(32, 2048, 3, generator=g).to(cuda)
    b = torch.rand(32, 16384, 3, generator=g).to(cuda)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    # (1) distance recomputed from the returned index (same rounding order) is bit-identical
    def redo(q, t, idx):
        nn = torch.gather(t, 1, idx.long()[..., None].expand(-1, -1, 3))
        dx, dy, dz = (nn - q).unbind(-1)
        return torch.addcmul(torch.addcmul(dy * dy, dx, dx), dz, dz)  # may or may not fuse: compare loosely
    assert torch.allclose(redo(a, b, i1), d1, rtol=1e-6, atol=0)
    assert torch.allclose(redo(b, a, i2), d2, rtol=1e-6, atol=0)
    # (2) no other target is closer (checked with cdist on a slice in float64)
    D = torch.cdist(a[:2].double(), b[:2].double()) ** 2
    assert torch.allclose(D.min(2).values.float(), d1[:2], rtol=1e-5)
    assert torch.allclose(D.min(1).values.float(), d2[:2], rtol=1e-5)
    # (3) permutation equivariance of the targets: permuting b permutes idx1 consistently
    perm = torch.randperm(16384, generator=g).to(cuda)
    d1p, _, i1p, _ = chamfer_3DDist()(a, b[:, perm])
    assert torch.equal(d1p, d1)
    assert torch.equal(torch.gather(b[:, perm], 1, i1p.long()[..., None].expand(-1, -1, 3)),
                       torch.gather(b, 1, i1.long()[..., None].expand(-1, -1, 3)))


def test_vs_reference_extension(cuda):
    """The unmodified reference chamfer_3D extension on the same GPU and inputs (oracle/_ref)."""
    ref = oracle.load_ref_ext("chamfer_3D")
    if ref is None:
        pytest.skip("oracle/_ref/chamfer_3D not built (needs /root/reference at build time)")
    from genpc_b200 import chamfer_3D as ours

    for (B, N, M, seed) in [(2, 2048, 16384, 0), (1, 5000, 777, 1), (3, 513, 1025, 2)]:
        a = torch.from_numpy(shape_cloud(seed, B, N)).to(cuda)
        b = torch.from_numpy(shape_cloud(seed + 100, B, M)).to(cuda)
        outs = []
        for mod in (ref, ours):
            d1 = torch.zeros(B, N, device=cuda); d2 = torch.zeros(B, M, device=cuda)
            i1 = torch.zeros(B, N, dtype=torch.int32, device=cuda); i2 = torch.zeros(B, M, dtype=torch.int32, device=cuda)
            mod.forward(a, b, d1, d2, i1, i2)
            torch.cuda.synchronize()
            g1 = torch.randn(B, N, generator=torch.Generator().manual_seed(seed)).to(cuda)
            g2 = torch.randn(B, M, generator=torch.Generator().manual_seed(seed + 1)).to(cuda)
            gx1 = torch.zeros_like(a); gx2 = torch.zeros_like(b)
            mod.backward(a, b, gx1, gx2, g1, g2, i1, i2)
            torch.cuda.synchronize()
            outs.append((d1, d2, i1, i2, gx1, gx2))
        r, o = outs
        for k in range(4):
            assert torch.equal(r[k], o[k]), f"output {k} differs from the reference extension"
        for k in (4, 5):
            scale = r[k].abs().max()
            assert (r[k] - o[k]).abs().max() <= 1e-5 * scale


def test_golden_vectors(cuda):
    """tests/golden/chamfer_ref_*.npz: outputs of the reference extension (made by tests/golden/make_golden.py)."""
    import glob

    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "chamfer_ref_*.npz")))
    if not files:
        pytest.skip("no golden vectors committed yet")
    for f in files:
        z = np.load(f)
        got = run_ours(z["xyz1"], z["xyz2"], cuda)
        for g, name in zip(got, ("dist1", "dist2", "idx1", "idx2")):
            assert np.array_equal(g, z[name]), f"{os.path.basename(f)}:{name}"


def test_backward_warp_aggregated_scatter(cuda):
    """Many points share one nearest neighbour (8 targets for 5000 queries): the scatter term takes the
    warp-aggregated path (lanes with equal neighbour summed by shuffles, one set of atomics per group)."""
    from genpc_b200.loss_functions import chamfer_3DDist

    a, b = rand_cloud(31, 2, 5000), rand_cloud(32, 2, 8)
    rng = np.random.default_rng(1)
    g1 = rng.standard_normal((2, 5000)).astype(np.float32)
    g2 = rng.standard_normal((2, 8)).astype(np.float32)
    ta = torch.from_numpy(a).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(b).to(cuda).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(ta, tb)
    ((d1 * torch.from_numpy(g1).to(cuda)).sum() + (d2 * torch.from_numpy(g2).to(cuda)).sum()).backward()
    e1, e2 = oracle.chamfer_backward(a, b, g1, g2, i1.cpu().numpy(), i2.cpu().numpy())
    for got, exp in ((ta.grad.cpu().numpy(), e1), (tb.grad.cpu().numpy(), e2)):
        assert np.abs(got - exp).max() <= 1e-5 * np.abs(exp).max() + 1e-12


def test_forward_fuzz_shapes_and_distributions(cuda):
    """Seeded sweep over awkward shapes (unaligned sizes, tiny / huge aspect ratios, both kernel paths) and point
    distributions (uniform, clustered, duplicated points, lattice ties): bit-exact against the oracle every time."""
    rng = np.random.default_rng(2026)
    for trial in range(40):
        B = int(rng.integers(1, 5))
        N = int(rng.choice([1, 2, 31, 33, 127, 129, 511, 513, 1000, 1023, 1025, 2049, 4099, 7001]))
        M = int(rng.choice([1, 3, 32, 100, 255, 257, 512, 777, 1024, 1300, 2047, 3001, 9973]))
        kind = trial % 4
        if kind == 0:
            a, b = rand_cloud(trial, B, N), rand_cloud(trial + 1000, B, M)
        elif kind == 1:   # clustered, far from the origin (large magnitudes, small differences)
            a, b = rand_cloud(trial, B, N, 0.01, 37.5), rand_cloud(trial + 1000, B, M, 0.01, 37.5)
        elif kind == 2:   # duplicated points: exact ties on both sides
            a, b = rand_cloud(trial, B, N), rand_cloud(trial + 1000, B, M)
            a[:, N // 2:] = a[:, :N - N // 2]
            b[:, M // 2:] = b[:, :M - M // 2]
        else:
            a, b = lattice_cloud(trial, B, N, side=5), lattice_cloud(trial + 1000, B, M, side=5)
        got = run_ours(a, b, cuda)
        exp = oracle.chamfer_forward(a, b)
        for g, e, name in zip(got, exp, ("dist1", "dist2", "idx1", "idx2")):
            assert np.array_equal(g.view(np.int32), e.view(np.int32)), (trial, B, N, M, kind, name)


@pytest.mark.parametrize("name", ["chamfer_l1", "chamfer_l2", "chamfer_partial_l1", "chamfer_partial_l2"])
def test_completionloss_fused_equals_torch_expression(cuda, name):
    """Completionloss' fused reduction / backward kernels against the reference's literal torch expressions
    (utils/loss_util.py:25-43): value and gradients within 1e-5 relative."""
    from genpc_b200.utils.loss_util import Completionloss

    a, b = shape_cloud(41, 3, 1500), shape_cloud(42, 3, 4000)
    res = {}
    for fused in (True, False):
        cl = Completionloss("cd_l1")
        cl.fused = fused
        ta = torch.from_numpy(a).to(cuda).requires_grad_(True)
        tb = torch.from_numpy(b).to(cuda).requires_grad_(True)
        loss = getattr(cl, name)(ta, tb)
        (loss * 3.0).backward()
        res[fused] = (float(loss), ta.grad.cpu().numpy(), tb.grad.cpu().numpy())
    assert abs(res[True][0] - res[False][0]) <= 1e-5 * abs(res[False][0])
    for k in (1, 2):
        scale = np.abs(res[False][k]).max()
        assert np.abs(res[True][k] - res[False][k]).max() <= 1e-5 * scale + 1e-12


@pytest.mark.parametrize("B,N,M,use_sqrt,w1,w2", [(3, 1500, 4000, 1, 0.5, 0.5), (2, 5000, 700, 0, 1.0, 1.0),
                                                  (4, 2048, 16384, 1, 1.0, 0.0), (2, 100, 300, 0, 1.0, 1.0)])
def test_fused_forward_c_abi(cuda, B, N, M, use_sqrt, w1, w2):
    """genpc_chamfer_forward_fused straight through the C ABI, three calls on ONE workspace (the 2nd and 3rd with
    workspace_armed = 1, i.e. without a memset): dist/idx bit-identical to the plain entry, the loss equal to the
    reference expression (loss_util.py:25-43) within 1e-6 relative, the two accumulators zero-filled.  The last
    shape does not take the symmetric path (same duties as separate launches)."""
    from genpc_b200 import _lib
    from genpc_b200.loss_functions import chamfer_3DDist

    L = _lib.lib()
    packed_bytes = L.genpc_chamfer_workspace_bytes(B, N, M)
    packed = torch.empty(packed_bytes, dtype=torch.uint8, device=cuda)
    loss_ws = torch.zeros(L.genpc_chamfer_fuse_workspace_bytes(B, N, M), dtype=torch.uint8, device=cuda)
    armed = 0
    for call in range(3):
        a, b = shape_cloud(100 + call, B, N), rand_cloud(200 + call, B, M, 0.8, -0.4)
        ta, tb = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
        d1 = torch.empty(B, N, device=cuda); d2 = torch.empty(B, M, device=cuda)
        i1 = torch.empty(B, N, dtype=torch.int32, device=cuda); i2 = torch.empty(B, M, dtype=torch.int32, device=cuda)
        z1 = torch.full((B, N, 3), 7.0, device=cuda); z2 = torch.full((B, M, 3), -3.0, device=cuda)
        out = torch.full((), float("nan"), device=cuda)
        fuse = _lib.ChamferFuse(armed, use_sqrt, w1, w2, out.data_ptr(), loss_ws.data_ptr(), loss_ws.numel(), z1.data_ptr(),
                                z2.data_ptr())
        rc = L.genpc_chamfer_forward_fused(_lib.ptr(ta), _lib.ptr(tb), _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(i1), _lib.ptr(i2),
                                           B, N, M, _lib.ptr(packed), packed_bytes, fuse, _lib.current_stream(cuda))
        assert rc == 0
        armed = 1
        e1, e2, j1, j2 = chamfer_3DDist()(ta, tb)
        assert torch.equal(d1, e1) and torch.equal(d2, e2) and torch.equal(i1, j1) and torch.equal(i2, j2), call
        f = torch.sqrt if use_sqrt else (lambda t: t)
        exp = w1 * f(e1.double()).mean() + (w2 * f(e2.double()).mean() if w2 else 0.0)
        assert abs(float(out) - float(exp)) <= 1e-6 * abs(float(exp)), call
        assert not z1.any() and not z2.any()


def test_fused_loss_steps_reuse_workspace(cuda):
    """Completionloss on the fused step, several steps at one shape (re-armed workspace) and interleaved shapes, each
    against the literal torch expression; a second backward through a retained graph gets fresh accumulators."""
    from genpc_b200.utils.loss_util import Completionloss

    fused, plain = Completionloss("cd_l1"), Completionloss("cd_l1")
    plain.fused = False
    for step, (B, N, M) in enumerate([(2, 3000, 1200), (2, 3000, 1200), (1, 600, 9000), (2, 3000, 1200), (1, 40, 50)]):
        a, b = shape_cloud(300 + step, B, N), shape_cloud(400 + step, B, M)
        res = []
        for cl in (fused, plain):
            ta = torch.from_numpy(a).to(cuda).requires_grad_(True)
            tb = torch.from_numpy(b).to(cuda).requires_grad_(True)
            loss = cl.get_loss(ta, tb)
            loss.backward(retain_graph=True)
            g_first = ta.grad.clone()
            ta.grad = None
            loss.backward()
            assert torch.allclose(ta.grad, g_first, rtol=1e-5, atol=1e-9)
            res.append((float(loss), ta.grad.cpu().numpy(), tb.grad.cpu().numpy()))
        assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[1][0])
        for k in (1, 2):
            scale = np.abs(res[1][k]).max()
            assert np.abs(res[0][k] - res[1][k]).max() <= 1e-5 * scale + 1e-12


def _off(t, k=1):
    """A contiguous copy of t whose data pointer is only 4-byte aligned (k floats past an aligned allocation)."""
    flat = torch.empty(t.numel() + k + 4, dtype=t.dtype, device=t.device)
    v = flat[k:k + t.numel()].view(t.shape)
    v.copy_(t)
    assert v.data_ptr() % 8 != 0 and v.is_contiguous()
    return v


def test_unaligned_pointers_take_the_scalar_paths(cuda):
    """The C ABI takes plain pointers: clouds, outputs and gradient arrays that are only 4-byte aligned must give the
    same results through the scalar fall-backs (LDG.32 fix-up, scalar zero-fill, three-scalar reductions) as aligned
    ones through the vector paths.  Forward bit-identical; backward within the atomics' 1e-5."""
    from genpc_b200 import chamfer_3D

    B, N, M = 2, 1500, 4100
    a, b = torch.from_numpy(shape_cloud(70, B, N)).to(cuda), torch.from_numpy(shape_cloud(71, B, M)).to(cuda)
    g1, g2 = torch.rand(B, N, device=cuda), torch.rand(B, M, device=cuda)
    res = []
    for mis in (False, True):
        f = _off if mis else (lambda t: t.clone())
        xa, xb = f(a), f(b)
        d1, d2 = f(torch.empty(B, N, device=cuda)), f(torch.empty(B, M, device=cuda))
        i1 = f(torch.empty(B, N, dtype=torch.int32, device=cuda)); i2 = f(torch.empty(B, M, dtype=torch.int32, device=cuda))
        z1, z2 = f(torch.full((B, N, 3), 5.0, device=cuda)), f(torch.full((B, M, 3), 5.0, device=cuda))
        for _ in range(2):                     # second call: re-armed workspace
            chamfer_3D.forward_fused(xa, xb, d1, d2, i1, i2, z1, z2)
        assert not z1.any() and not z2.any()
        chamfer_3D.backward(xa, xb, z1, z2, f(g1), f(g2), i1, i2)
        res.append([t.clone() for t in (d1, d2, i1, i2, z1, z2)])
    for k in range(4):
        assert torch.equal(res[0][k], res[1][k]), k
    for k in (4, 5):
        assert torch.allclose(res[0][k], res[1][k], rtol=1e-5, atol=1e-7), k
    e = oracle.chamfer_forward(a.cpu().numpy(), b.cpu().numpy())
    for k in range(4):
        assert np.array_equal(res[1][k].cpu().numpy(), e[k]), k


def test_full_c2_bit_exact_vs_oracle_and_reference(cuda):
    """BASELINE config C2 AT FULL BATCH (B=32, 2048 x 16384, the bench's own synthetic PCN batch): bit-exact against the
    oracle (all 32 scans: ~0.3 s of OpenMP CPU work) AND against the live reference extension; gradients of
    chamfer_l2 within 1e-5 of the double-accumulated oracle."""
    from genpc_b200.synthetic import pcn_batch
    from genpc_b200.utils.loss_util import Completionloss

    part, comp = pcn_batch(0, 32, 2048, 16384)
    got = run_ours(part, comp, cuda)
    exp = oracle.chamfer_forward(part, comp)
    for g, e, name in zip(got, exp, ("dist1", "dist2", "idx1", "idx2")):
        assert np.array_equal(g.view(np.int32), e.view(np.int32)), f"{name} differs ({(g != e).sum()} entries)"
    assert (exp[0] > 0).mean() > 0.99, "the partial cloud must not be a subset of the complete one (r01 finding)"
    ref = oracle.load_ref_ext("chamfer_3D")
    if ref is not None:
        a, b = torch.from_numpy(part).to(cuda), torch.from_numpy(comp).to(cuda)
        d1 = torch.zeros(32, 2048, device=cuda); d2 = torch.zeros(32, 16384, device=cuda)
        i1 = torch.zeros(32, 2048, dtype=torch.int32, device=cuda); i2 = torch.zeros(32, 16384, dtype=torch.int32, device=cuda)
        ref.forward(a, b, d1, d2, i1, i2)
        for g, r, name in zip(got, (d1, d2, i1, i2), ("dist1", "dist2", "idx1", "idx2")):
            assert np.array_equal(g, r.cpu().numpy()), f"{name} differs from the reference extension"
    ta = torch.from_numpy(part).to(cuda).requires_grad_(True)
    tb = torch.from_numpy(comp).to(cuda).requires_grad_(True)
    loss = Completionloss("cd_l2").get_loss(ta, tb)
    loss.backward()
    want = exp[0].astype(np.float64).mean() + exp[1].astype(np.float64).mean()
    assert abs(float(loss) - want) <= 1e-6 * want
    g1 = np.full((32, 2048), 1.0 / (32 * 2048), np.float32)
    g2 = np.full((32, 16384), 1.0 / (32 * 16384), np.float32)
    e1, e2 = oracle.chamfer_backward(part, comp, g1, g2, exp[2], exp[3])
    assert np.abs(ta.grad.cpu().numpy() - e1).max() <= 1e-5 * np.abs(e1).max()
    assert np.abs(tb.grad.cpu().numpy() - e2).max() <= 1e-5 * np.abs(e2).max()


@pytest.mark.parametrize("N,M", [(300, 200), (700, 600), (5000, 1300), (2048, 16384)])
def test_nan_and_inf_points_stay_in_range(cuda, N, M):
    """a NaN / Inf / overflowing point must not leave a packed word unarmed -- every returned index is
    inside [0, other cloud), the finite points' results are unchanged, and the backward writes nothing outside its two
    gradient arrays (guard bands around them stay intact)."""
    from genpc_b200 import chamfer_3D

    B = 2
    a, b = rand_cloud(81, B, N), rand_cloud(82, B, M)
    clean = oracle.chamfer_forward(a, b)
    a[0, 3] = np.nan                       # a NaN query / row
    b[0, 5, 1] = np.nan                    # a NaN target / column
    a[1, 7] = 3e38                         # squared distance overflows to +inf
    b[1, M - 1] = np.inf
    ta, tb = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    d1 = torch.empty(B, N, device=cuda); d2 = torch.empty(B, M, device=cuda)
    i1 = torch.full((B, N), -7, dtype=torch.int32, device=cuda); i2 = torch.full((B, M), -7, dtype=torch.int32, device=cuda)
    chamfer_3D.forward(ta, tb, d1, d2, i1, i2)
    torch.cuda.synchronize()
    assert int(i1.min()) >= 0 and int(i1.max()) < M and int(i2.min()) >= 0 and int(i2.max()) < N
    # rows / columns that never met a poisoned point keep their exact answers
    keep1 = np.ones((B, N), bool); keep1[0, 3] = keep1[1, 7] = False
    keep1 &= ~((clean[2] == 5) & (np.arange(B)[:, None] == 0)) & ~((clean[2] == M - 1) & (np.arange(B)[:, None] == 1))
    keep2 = np.ones((B, M), bool); keep2[0, 5] = keep2[1, M - 1] = False
    keep2 &= ~((clean[3] == 3) & (np.arange(B)[:, None] == 0)) & ~((clean[3] == 7) & (np.arange(B)[:, None] == 1))
    assert np.array_equal(d1.cpu().numpy()[keep1], clean[0][keep1]) and np.array_equal(i1.cpu().numpy()[keep1], clean[2][keep1])
    assert np.array_equal(d2.cpu().numpy()[keep2], clean[1][keep2]) and np.array_equal(i2.cpu().numpy()[keep2], clean[3][keep2])
    # backward between guard bands
    guard = 64
    buf1 = torch.full((B * N * 3 + 2 * guard,), 123.0, device=cuda); buf2 = torch.full((B * M * 3 + 2 * guard,), 321.0, device=cuda)
    gx1 = buf1[guard:guard + B * N * 3].view(B, N, 3); gx2 = buf2[guard:guard + B * M * 3].view(B, M, 3)
    gx1.zero_(); gx2.zero_()
    chamfer_3D.backward(ta, tb, gx1, gx2, torch.ones(B, N, device=cuda), torch.ones(B, M, device=cuda), i1, i2)
    torch.cuda.synchronize()
    for buf, v in ((buf1, 123.0), (buf2, 321.0)):
        assert bool((buf[:guard] == v).all()) and bool((buf[-guard:] == v).all())
    # caller-supplied garbage indices contribute nothing instead of touching foreign memory
    bad1 = torch.full_like(i1, -1); bad2 = torch.full_like(i2, 2 ** 30)
    gx1.zero_(); gx2.zero_()
    chamfer_3D.backward(ta, tb, gx1, gx2, torch.ones(B, N, device=cuda), torch.ones(B, M, device=cuda), bad1, bad2)
    torch.cuda.synchronize()
    assert not gx1.any() and not gx2.any()
    for buf, v in ((buf1, 123.0), (buf2, 321.0)):
        assert bool((buf[:guard] == v).all()) and bool((buf[-guard:] == v).all())


@pytest.mark.parametrize("name", ["cd_l1", "cd_l2"])
def test_graphed_loss_step_equals_eager(cuda, name):
    """GraphedLossStep (one CUDA-graph launch per step) against the eager fused step on changing inputs: loss and both
    gradients identical bit for bit (same kernels, same deterministic reductions) -- gradients up to the atomics' order."""
    from genpc_b200.utils.loss_util import Completionloss, GraphedLossStep

    B, N, M = 4, 2048, 4096
    cl = Completionloss(name)
    a0, b0 = torch.from_numpy(shape_cloud(1, B, N)).to(cuda), torch.from_numpy(shape_cloud(2, B, M)).to(cuda)
    step = GraphedLossStep(cl, a0, b0)
    for k in range(4):
        a = torch.from_numpy(shape_cloud(10 + k, B, N)).to(cuda)
        b = torch.from_numpy(shape_cloud(20 + k, B, M)).to(cuda)
        loss, ga, gb = step(a, b)
        ea = a.clone().requires_grad_(True)
        eb = b.clone().requires_grad_(True)
        el = cl.get_loss(ea, eb)
        el.backward()
        torch.cuda.synchronize()
        assert float(loss) == float(el)
        assert torch.allclose(ga, ea.grad, rtol=1e-5, atol=1e-9) and torch.allclose(gb, eb.grad, rtol=1e-5, atol=1e-9)
    assert step() [0] is step.loss
