"""Generated-mesh ingestion (reference utils/dataUtils.py:217-250 `glb2point`, reg_xyz.py:99-125):
the dependency-free GLB reader, the seeded GPU surface sampler against its oracle (bit-exact), and reg() /
object_pose_optimization on a workspace laid out exactly like the reference's (color_point.ply + <flag>_<model>.glb)."""
import json
import os
import struct
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import mesh as OM


def cube():
    v = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32)
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                  [1, 5, 7], [1, 7, 3]], np.int32)
    return v, f


def test_glb_write_read_round_trip_and_node_matrix(tmp_path):
    from genpc_b200.utils.glb import read_glb, write_glb

    v, f = cube()
    col = np.random.default_rng(0).random((8, 3)).astype(np.float32)
    p = str(tmp_path / "a.glb")
    write_glb(p, v, f, col)
    rv, rf, rc = read_glb(p)
    assert np.array_equal(rv, v) and np.array_equal(rf, f) and np.array_equal(rc, col)
    M = np.eye(4)
    M[:3, :3] = np.diag([2.0, 1.0, 0.5]) @ np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    M[:3, 3] = [0.5, -1.0, 3.0]
    write_glb(p, v, f, None, node_matrix=M, index_dtype=np.uint16)
    rv, rf, rc = read_glb(p)
    assert rc is None and np.array_equal(rf, f)
    assert np.allclose(rv, v @ M[:3, :3].T + M[:3, 3], atol=1e-6)


def test_glb_reader_scene_graph_trs_strided_views_and_ubyte_colours(tmp_path):
    """A hand-built GLB: two nodes (parent translation, child rotation + scale) sharing one mesh with two primitives,
    interleaved (strided) vertex buffer, normalised-ubyte RGBA colours, u8 indices, plus a line primitive that must be skipped."""
    from genpc_b200.utils.glb import read_glb

    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    rgba = np.array([[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255], [255, 255, 255, 255]], np.uint8)
    inter = b"".join(tri[i].tobytes() + rgba[i].tobytes() for i in range(4))          # 16-byte stride
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint8).tobytes() + b"\0\0"
    blob = inter + idx
    js = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
          "nodes": [{"translation": [10, 0, 0], "children": [1]},
                    {"mesh": 0, "rotation": [0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)], "scale": [2, 2, 2]}],
          "meshes": [{"primitives": [
              {"attributes": {"POSITION": 0, "COLOR_0": 1}, "indices": 2},
              {"attributes": {"POSITION": 0}, "indices": 3, "material": 0},
              {"attributes": {"POSITION": 0}, "mode": 1}]}],
          "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.25, 0.5, 0.75, 1.0]}}],
          "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                        {"bufferView": 0, "byteOffset": 12, "componentType": 5121, "normalized": True, "count": 4, "type": "VEC4"},
                        {"bufferView": 1, "componentType": 5121, "count": 3, "type": "SCALAR"},
                        {"bufferView": 1, "byteOffset": 3, "componentType": 5121, "count": 3, "type": "SCALAR"}],
          "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 64, "byteStride": 16},
                          {"buffer": 0, "byteOffset": 64, "byteLength": 6}],
          "buffers": [{"byteLength": len(blob)}]}
    jb = json.dumps(js).encode()
    jb += b" " * (-len(jb) % 4)
    p = str(tmp_path / "b.glb")
    with open(p, "wb") as fh:
        fh.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(jb) + 8 + len(blob)))
        fh.write(struct.pack("<II", len(jb), 0x4E4F534A) + jb + struct.pack("<II", len(blob), 0x004E4942) + blob)
    v, f, c = read_glb(p)
    Rz = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    want = (tri.astype(np.float64) * 2) @ Rz.T + [10, 0, 0]
    assert v.shape == (8, 3) and np.allclose(v[:4], want, atol=1e-6) and np.allclose(v[4:], want, atol=1e-6)
    assert np.array_equal(f, [[0, 1, 2], [4, 6, 7]])
    assert np.allclose(c[:4], rgba[:, :3] / 255.0) and np.allclose(c[4:], [0.25, 0.5, 0.75])


def test_glb_reader_rejects_garbage(tmp_path):
    from genpc_b200.utils.glb import GlbError, read_glb

    p = str(tmp_path / "c.glb")
    open(p, "wb").write(b"not a glb file at all")
    with pytest.raises(GlbError):
        read_glb(p)
    open(p, "wb").write(struct.pack("<III", 0x46546C67, 1, 12))
    with pytest.raises(GlbError):
        read_glb(p)


def test_oracle_sampler_properties():
    """Samples lie on their triangle, colours are the barycentric blend, faces are drawn in proportion to their area
    (chi-square on a mesh with very uneven faces), degenerate faces are never drawn, same seed -> same samples."""
    v, f = cube()
    v = v * np.array([4.0, 1.0, 0.25], np.float32)                  # face areas 1 : 0.25 : 4 (pairs)
    f = np.concatenate([f, [[0, 0, 1], [2, 2, 2]]]).astype(np.int32)  # two degenerate faces
    col = (v - v.min(0)) / (v.max(0) - v.min(0))
    n = 6000
    xyz, rgb, face = OM.sample(v, f, n, 7, col)
    assert face.max() < 12
    a, b, c = v[f[face, 0]], v[f[face, 1]], v[f[face, 2]]
    T = np.stack([b - a, c - a], -1).astype(np.float64)               # solve the barycentric coordinates back
    uv = np.stack([np.linalg.lstsq(T[i], (xyz[i] - a[i]).astype(np.float64), rcond=None)[0] for i in range(0, n, 37)])
    assert (uv >= -1e-5).all() and (uv.sum(1) <= 1 + 1e-5).all()
    w = np.concatenate([1 - uv.sum(1, keepdims=True), uv], 1)
    blend = (w[:, :, None] * np.stack([col[f[face[::37], k]] for k in range(3)], 1)).sum(1)
    assert np.allclose(rgb[::37], blend, atol=1e-5)
    areas = OM.face_areas(v, f).astype(np.float64)
    assert areas[12] == 0 and areas[13] == 0
    exp = areas[:12] / areas.sum() * n
    chi2 = ((np.bincount(face, minlength=12)[:12] - exp) ** 2 / exp).sum()
    assert chi2 < 40.0                                                # 11 dof: p ~ 4e-5
    xyz2, _, face2 = OM.sample(v, f, 500, 7, col)
    assert np.array_equal(xyz2, xyz[:500]) and np.array_equal(face2, face[:500])
    assert not np.array_equal(OM.sample(v, f, 500, 8, col)[2], face[:500])


@pytest.mark.gpu
def test_sampler_kernel_bit_exact_vs_oracle(cuda):
    import torch

    from genpc_b200.synthetic import superquadric_mesh
    from genpc_b200.utils.glb import sample_mesh

    for seed, (ne, no) in enumerate([(6, 8), (24, 40), (48, 96)]):
        v, f, col = superquadric_mesh(seed, ne, no)
        f = np.concatenate([f, [[0, 0, 0]]]).astype(np.int32)        # a degenerate face rides along
        tv, tf, tc = (torch.from_numpy(x).to(cuda) for x in (v, f, col))
        n = 3000
        xyz, rgb, face = sample_mesh(tv, tf, n, 1234 + seed, tc, return_face=True)
        exyz, ergb, eface = OM.sample(v, f, n, 1234 + seed, col)
        assert np.array_equal(face.cpu().numpy(), eface)
        assert np.array_equal(xyz.cpu().numpy().view(np.int32), exyz.view(np.int32))
        assert np.array_equal(rgb.cpu().numpy().view(np.int32), ergb.view(np.int32))
        xyz2, rgb2 = sample_mesh(tv, tf, n, 1234 + seed, None)        # no vertex colours: 0.5 grey like the reference
        assert torch.equal(xyz2, xyz) and bool((rgb2 == 0.5).all())


@pytest.mark.gpu
def test_glb2point_full_size_and_density(cuda, tmp_path):
    """163 840 samples (reg_xyz.py:125): all on the surface (distance to the analytic point set tiny), area-uniform
    (voxel occupancy roughly flat), voxel down-sampling averages colours with the points."""
    import torch

    from genpc_b200.synthetic import superquadric_mesh
    from genpc_b200.utils.glb import glb2point, write_glb

    v, f, col = superquadric_mesh(3, 64, 128)
    p = str(tmp_path / "shape.glb")
    write_glb(p, v, f, col)
    pts, rgb = glb2point(p, num_points=163840, seed=0, device=cuda)
    assert pts.shape == (163840, 3) and rgb.shape == (163840, 3) and pts.is_cuda
    lo, hi = torch.from_numpy(v.min(0)).to(cuda), torch.from_numpy(v.max(0)).to(cuda)
    assert bool((pts >= lo - 1e-6).all()) and bool((pts <= hi + 1e-6).all())
    assert torch.allclose(rgb, ((pts - lo) / (hi - lo + 1e-8)).clamp(0, 1), atol=2e-3)   # colour = linear in position
    again, _ = glb2point(p, num_points=163840, seed=0, device=cuda)
    assert torch.equal(again, pts)
    dpts, drgb = glb2point(p, down_sample=0.05, num_points=20000, seed=0, device=cuda)
    assert dpts.shape == drgb.shape and dpts.shape[0] < 3000
    assert torch.allclose(drgb, ((dpts - lo) / (hi - lo + 1e-8)).clamp(0, 1), atol=5e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["trellis", "instantmesh"])
def test_scale_adapter_on_reference_workspace_layout(cuda, tmp_path, model):
    """ScaleAdapter(cfg).scaleReg(flag) on <out>/<flag>/color_point.ply + <flag>_<model>.glb (reg_xyz.py:99-108, 220)."""
    import torch

    from genpc_b200.ScaleAdapter import ScaleAdapter
    from genpc_b200.optim_registration.diff_obj_pose import object_pose_optimization
    from genpc_b200.synthetic import partial_view, rigid_perturb, superquadric, superquadric_mesh
    from genpc_b200.utils.dataUtils import read_ply_xyz, write_ply_xyz
    from genpc_b200.utils.glb import write_glb

    flag = "00042"
    os.makedirs(tmp_path / flag)
    v, f, col = superquadric_mesh(9, 48, 96)
    if model == "instantmesh":          # the generator's frame differs by the two 90-degree turns reg() undoes (:133-138)
        from genpc_b200.reg_xyz import get_rotate_matrix
        Rfix = get_rotate_matrix("y", 90) @ get_rotate_matrix("x", 90)
        v = (v.astype(np.float64) @ Rfix).astype(np.float32)          # so that (v @ Rx.T) @ Ry.T is the canonical frame
    write_glb(str(tmp_path / flag / f"{flag}_{model}.glb"), v * 1.3, f, col)
    scan, gt = rigid_perturb(partial_view(superquadric(9, 30000), 9, 9000), 9, max_rot_deg=10, max_t=0.03, scale_range=(0.8, 0.95))
    write_ply_xyz(str(tmp_path / flag / "color_point.ply"), scan, np.full_like(scan, 0.25))
    cfg = SimpleNamespace(output_path=str(tmp_path), generative_model=model, device="cuda:0", dataset="redwood")
    out = ScaleAdapter(cfg).scaleReg(flag)
    fused, frgb = read_ply_xyz(str(tmp_path / flag / f"{flag}_fused.ply"))
    assert fused.shape[0] == out["fused"].shape[0] and 5000 < fused.shape[0] <= 20000 and frgb is not None
    assert np.isfinite(fused).all()
    # the scan's points (grey 0.25) survive into the fused cloud next to generated points (coloured by position)
    grey = np.isclose(frgb, 0.25, atol=1 / 255).all(1)
    assert 0.05 < grey.mean() < 0.95
    # the fused cloud stays in the scan's frame: its scan part is a subset of the scan
    from genpc_b200.loss_functions import chamfer_3DDist
    d, _, _, _ = chamfer_3DDist()(torch.from_numpy(fused[grey]).to(cuda)[None], torch.from_numpy(scan).to(cuda)[None])
    assert float(d.max().sqrt()) < 1e-5       # round trip through diff / coarse transforms and their inverses (fp32)
    if model == "trellis":
        cwd = os.getcwd()
        os.chdir(tmp_path)
        try:
            T = object_pose_optimization(str(tmp_path / flag / f"{flag}_{model}.glb"), str(tmp_path / flag / "color_point.ply"),
                                         radius=0.02, lr=0.01, iters=50, device=cuda)
        finally:
            os.chdir(cwd)
        assert T.shape == (4, 4) and np.isfinite(T).all() and os.path.exists(tmp_path / "final_transform.npy")


def test_glb_reader_base_colour_texture(tmp_path):
    """Colours from the base-colour texture at TEXCOORD_0 (the reference bakes TextureVisuals to vertex colours,
    utils/dataUtils.py:223-224): a 2 x 2 PNG embedded in the GLB, nearest texel, glTF's top-left uv origin."""
    cv2 = pytest.importorskip("cv2")
    from genpc_b200.utils.glb import read_glb

    img = np.array([[[255, 0, 0], [0, 255, 0]], [[0, 0, 255], [255, 255, 0]]], np.uint8)       # RGB, row 0 = top
    ok, png = cv2.imencode(".png", img[..., ::-1])
    assert ok
    png = png.tobytes()
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    uv = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], np.float32)                  # -> texels (0,0) (0,1) (1,0) (1,1)
    idx = np.array([0, 1, 2, 1, 3, 2], np.uint16)
    parts = [pos.tobytes(), uv.tobytes(), idx.tobytes(), png]
    offs, blob = [], b""
    for part in parts:
        offs.append(len(blob))
        blob += part + b"\\0" * (-len(part) % 4)
    js = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
          "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0}]}],
          "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
          "textures": [{"source": 0}], "images": [{"bufferView": 3, "mimeType": "image/png"}],
          "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
                        {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC2"},
                        {"bufferView": 2, "componentType": 5123, "count": 6, "type": "SCALAR"}],
          "bufferViews": [{"buffer": 0, "byteOffset": offs[k], "byteLength": len(parts[k])} for k in range(4)],
          "buffers": [{"byteLength": len(blob)}]}
    jb = json.dumps(js).encode()
    jb += b" " * (-len(jb) % 4)
    p = str(tmp_path / "t.glb")
    with open(p, "wb") as fh:
        fh.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(jb) + 8 + len(blob)))
        fh.write(struct.pack("<II", len(jb), 0x4E4F534A) + jb + struct.pack("<II", len(blob), 0x004E4942) + blob)
    v, f, c = read_glb(p)
    assert np.array_equal(v, pos) and np.array_equal(f, [[0, 1, 2], [1, 3, 2]])
    assert np.allclose(c, np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0]], np.float32))
