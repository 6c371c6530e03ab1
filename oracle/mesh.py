"""CPU oracle of the mesh surface sampler -- TEST INFRASTRUCTURE ONLY (imported by tests/ and smoke(), never by genpc_b200/).

Restates csrc/mesh.cu operation by operation in numpy (float32 ops are individually rounded in numpy exactly like the
__f*_rn intrinsics of the kernel; no fma anywhere).  What it replaces in the reference: the sampling step of glb2point
(utils/dataUtils.py:217-250, trimesh `mesh.sample` + barycentric colours).  trimesh is not vendored and its sampler is
unseeded, so the semantics are defined here: **parity unpinned** against trimesh, bit-exact between oracle and kernel.
"""
import numpy as np

_M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def face_areas(verts, faces):
    v = np.asarray(verts, np.float32)
    f = np.asarray(faces, np.int64)
    ok = ((f >= 0) & (f < len(v))).all(1)
    fs = np.where(ok[:, None], f, 0)
    a, b, c = v[fs[:, 0]], v[fs[:, 1]], v[fs[:, 2]]
    e1, e2 = (b - a).astype(np.float32), (c - a).astype(np.float32)
    with np.errstate(all="ignore"):
        cx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
        cy = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
        cz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
        n2 = (cx * cx + cy * cy) + cz * cz
        ar = np.float32(0.5) * np.sqrt(n2, dtype=np.float32)
    ar = np.where(np.isfinite(ar) & ok, ar, np.float32(0)).astype(np.float32)
    return ar


def cumulative_weights(areas):
    """floor(area / max_area * 2^32) as uint64, inclusive prefix sums (exact integers)."""
    a = np.asarray(areas, np.float64)
    amax = a.max() if len(a) else 0.0
    if not amax > 0:
        raise ValueError("mesh has no face with a positive area")
    w = np.floor(a / amax * 4294967296.0).astype(np.uint64)
    return np.cumsum(w, dtype=np.uint64)


def sample(verts, faces, n, seed, vertex_rgb=None):
    """-> (xyz [n,3] f32, rgb [n,3] f32, face [n] i32), bit-identical to genpc_mesh_sample."""
    v = np.asarray(verts, np.float32)
    f = np.asarray(faces, np.int64)
    cum = cumulative_weights(face_areas(v, f))
    total = int(cum[-1])
    xyz = np.empty((n, 3), np.float32)
    rgb = np.empty((n, 3), np.float32)
    face = np.empty(n, np.int32)
    k24 = np.float32(5.9604644775390625e-08)
    one = np.float32(1.0)
    for i in range(n):
        r0 = splitmix64((seed ^ ((0xD1B54A32D192ED03 * (2 * i + 1)) & _M64)) & _M64)
        r1 = splitmix64(r0)
        target = (r0 * total) >> 64
        fi = int(np.searchsorted(cum, np.uint64(target), side="right"))
        u = np.float32(r1 >> 40) * k24
        w = np.float32((r1 >> 16) & 0xFFFFFF) * k24
        if np.float32(u + w) > one:
            u, w = np.float32(one - u), np.float32(one - w)
        i0, i1, i2 = f[fi]
        a, b, c = v[i0], v[i1], v[i2]
        e1, e2 = (b - a).astype(np.float32), (c - a).astype(np.float32)
        xyz[i] = a + ((e1 * u).astype(np.float32) + (e2 * w).astype(np.float32)).astype(np.float32)
        if vertex_rgb is None:
            rgb[i] = 0.5
        else:
            col = np.asarray(vertex_rgb, np.float32)
            w0 = np.float32(np.float32(one - u) - w)
            rgb[i] = np.clip(((w0 * col[i0]).astype(np.float32) + (u * col[i1]).astype(np.float32)).astype(np.float32)
                             + (w * col[i2]).astype(np.float32), 0, 1)
        face[i] = fi
    return xyz, rgb, face
