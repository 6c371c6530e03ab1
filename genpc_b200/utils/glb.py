"""Binary glTF 2.0 (.glb) ingestion for the fusion stage: a dependency-free reader (+ a minimal writer for synthetic
workspaces and tests) and `glb2point`, the drop-in for the reference's utils/dataUtils.py:217-250.

The reference loads the generated mesh with trimesh (`trimesh.load(..., file_type='glb')`, `Scene.dump(concatenate=True)`,
`TextureVisuals.to_color()`), samples `num_points` surface points with `mesh.sample` and interpolates vertex colours
barycentrically.  trimesh is not vendored: the container format is parsed here (scene graph flattened with node
matrices / TRS in float64, triangle primitives only), the sampling runs on the GPU (csrc/mesh.cu, seeded and
bit-reproducible against oracle/mesh.py).  Colours: COLOR_0 if present, else the base-colour texture sampled at
TEXCOORD_0 (nearest texel, needs cv2 to decode the embedded PNG/JPEG), else baseColorFactor, else 0.5 grey (:234-236).
"""
import json
import struct

import numpy as np
import torch

from .. import _lib

_CT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NC = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class GlbError(ValueError):
    pass


def _chunks(raw):
    if len(raw) < 12:
        raise GlbError("not a GLB file (too short)")
    magic, version, length = struct.unpack_from("<III", raw, 0)
    if magic != 0x46546C67:
        raise GlbError("not a GLB file (bad magic)")
    if version != 2:
        raise GlbError(f"glTF version {version} is not supported (2 expected)")
    off, js, binc = 12, None, b""
    while off + 8 <= min(length, len(raw)):
        clen, ctype = struct.unpack_from("<II", raw, off)
        data = raw[off + 8:off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(data.decode("utf-8"))
        elif ctype == 0x004E4942 and not binc:
            binc = data
        off += 8 + clen + (-clen % 4)
    if js is None:
        raise GlbError("GLB without a JSON chunk")
    return js, binc


def _accessor(js, binc, idx):
    acc = js["accessors"][idx]
    dt, nc, cnt = np.dtype(_CT[acc["componentType"]]), _NC[acc["type"]], acc["count"]
    if "bufferView" not in acc:
        return np.zeros((cnt, nc), dt), acc
    bv = js["bufferViews"][acc["bufferView"]]
    if bv.get("buffer", 0) != 0:
        raise GlbError("external buffers are not supported (GLB-embedded only)")
    start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = bv.get("byteStride", 0) or dt.itemsize * nc
    if stride == dt.itemsize * nc:
        arr = np.frombuffer(binc, dt, cnt * nc, start).reshape(cnt, nc)
    else:
        rows = np.lib.stride_tricks.as_strided(np.frombuffer(binc, np.uint8, offset=start), (cnt, dt.itemsize * nc), (stride, 1))
        arr = np.ascontiguousarray(rows).view(dt).reshape(cnt, nc)
    return arr, acc


def _node_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], np.float64).reshape(4, 4).T      # column-major on disk
    M = np.eye(4)
    if "scale" in node:
        M = np.diag(list(node["scale"]) + [1.0]) @ M
    if "rotation" in node:
        x, y, z, w = node["rotation"]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        R4 = np.eye(4)
        R4[:3, :3] = R
        M = R4 @ M
    if "translation" in node:
        T = np.eye(4)
        T[:3, 3] = node["translation"]
        M = T @ M
    return M


def _decode_image(js, binc, tex_index):
    try:
        import cv2
    except Exception:                                                    # pragma: no cover
        return None
    img = js["images"][js["textures"][tex_index]["source"]]
    if "bufferView" not in img:
        return None
    bv = js["bufferViews"][img["bufferView"]]
    data = np.frombuffer(binc, np.uint8, bv["byteLength"], bv.get("byteOffset", 0))
    im = cv2.imdecode(data, cv2.IMREAD_UNCHANGED)
    if im is None:
        return None
    if im.ndim == 2:
        im = np.repeat(im[..., None], 3, 2)
    return im[..., 2::-1].astype(np.float32) / 255.0                     # BGR(A) -> RGB


def _primitive_colors(js, binc, prim, n):
    attr = prim["attributes"]
    if "COLOR_0" in attr:
        col, acc = _accessor(js, binc, attr["COLOR_0"])
        col = col[:, :3].astype(np.float32)
        if acc["componentType"] != 5126:
            col = col / float(np.iinfo(_CT[acc["componentType"]]).max)
        return col, True
    mat = js["materials"][prim["material"]] if "material" in prim and "materials" in js else {}
    pbr = mat.get("pbrMetallicRoughness", {})
    if "baseColorTexture" in pbr and "TEXCOORD_0" in attr:
        im = _decode_image(js, binc, pbr["baseColorTexture"]["index"])
        if im is not None:
            uv, acc = _accessor(js, binc, attr["TEXCOORD_0"])
            uv = uv.astype(np.float64)
            if acc["componentType"] != 5126:
                uv = uv / float(np.iinfo(_CT[acc["componentType"]]).max)
            h, w = im.shape[:2]
            x = np.round(uv[:, 0] * (w - 1)).astype(np.int64) % w        # glTF uv origin is top-left: nearest texel, wrapped
            y = np.round(uv[:, 1] * (h - 1)).astype(np.int64) % h
            return im[y, x].astype(np.float32), True
    if "baseColorFactor" in pbr:
        return np.tile(np.array(pbr["baseColorFactor"][:3], np.float32), (n, 1)), True
    return np.full((n, 3), 0.5, np.float32), False


def read_glb(path):
    """-> (verts float32 [V,3], faces int32 [F,3], vertex_rgb float32 [V,3] in [0,1] or None).
    All triangle primitives of the default scene, node transforms applied, concatenated (trimesh: Scene.dump(concatenate=True))."""
    js, binc = _chunks(open(path, "rb").read())
    nodes = js.get("nodes", [])
    scenes = js.get("scenes", [])
    roots = scenes[js.get("scene", 0)].get("nodes", []) if scenes else list(range(len(nodes)))
    V, Fs, C, any_col, base = [], [], [], False, 0
    stack = [(r, np.eye(4)) for r in reversed(roots)]
    seen = 0
    while stack:
        ni, parent = stack.pop()
        seen += 1
        if seen > 100000:
            raise GlbError("scene graph too deep / cyclic")
        node = nodes[ni]
        M = parent @ _node_matrix(node)
        for ch in reversed(node.get("children", [])):
            stack.append((ch, M))
        if "mesh" not in node:
            continue
        for prim in js["meshes"][node["mesh"]]["primitives"]:
            if prim.get("mode", 4) != 4 or "POSITION" not in prim["attributes"]:
                continue                                                  # points / lines / strips carry no surface to sample
            pos, _ = _accessor(js, binc, prim["attributes"]["POSITION"])
            pos = pos[:, :3].astype(np.float64)
            if "indices" in prim:
                idx, _ = _accessor(js, binc, prim["indices"])
                idx = idx.reshape(-1).astype(np.int64)
            else:
                idx = np.arange(len(pos), dtype=np.int64)
            idx = idx[:len(idx) // 3 * 3].reshape(-1, 3)
            col, has = _primitive_colors(js, binc, prim, len(pos))
            any_col |= has
            V.append(pos @ M[:3, :3].T + M[:3, 3])
            Fs.append(idx + base)
            C.append(col)
            base += len(pos)
    if not V:
        raise GlbError(f"{path}: no triangle primitive found")
    verts = np.concatenate(V).astype(np.float32)
    faces = np.concatenate(Fs).astype(np.int32)
    if faces.size and (faces.min() < 0 or faces.max() >= len(verts)):
        raise GlbError(f"{path}: face index out of range")
    return verts, faces, (np.concatenate(C).astype(np.float32) if any_col else None)


def write_glb(path, verts, faces, vertex_rgb=None, node_matrix=None, index_dtype=np.uint32):
    """Minimal GLB writer (one mesh, one triangle primitive, optional float COLOR_0 and node matrix): synthetic stand-ins for
    the generators' outputs (InstantMesh / TRELLIS .glb files are out of scope) and reader tests."""
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(np.asarray(faces).reshape(-1), index_dtype)
    blobs, views, accs = [], [], []

    def add(arr, target, ctype, typ, minmax=False):
        data = arr.tobytes()
        off = sum(len(b) for b in blobs)
        blobs.append(data + b"\0" * (-len(data) % 4))
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(data), "target": target})
        a = {"bufferView": len(views) - 1, "componentType": ctype, "count": len(arr) if arr.ndim > 1 else arr.size, "type": typ}
        if minmax:
            a["min"], a["max"] = arr.min(0).tolist(), arr.max(0).tolist()
        accs.append(a)
        return len(accs) - 1

    prim = {"attributes": {"POSITION": add(verts, 34962, 5126, "VEC3", True)}, "mode": 4}
    prim["indices"] = add(faces, 34963, {np.uint32: 5125, np.uint16: 5123, np.uint8: 5121}[index_dtype], "SCALAR")
    if vertex_rgb is not None:
        prim["attributes"]["COLOR_0"] = add(np.ascontiguousarray(vertex_rgb, np.float32), 34962, 5126, "VEC3")
    node = {"mesh": 0}
    if node_matrix is not None:
        node["matrix"] = np.asarray(node_matrix, np.float64).reshape(4, 4).T.reshape(-1).tolist()
    js = {"asset": {"version": "2.0", "generator": "genpc_b200"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [node],
          "meshes": [{"primitives": [prim]}], "accessors": accs, "bufferViews": views,
          "buffers": [{"byteLength": sum(len(b) for b in blobs)}]}
    jb = json.dumps(js, separators=(",", ":")).encode("utf-8")
    jb += b" " * (-len(jb) % 4)
    bb = b"".join(blobs)
    with open(path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(jb) + 8 + len(bb)))
        f.write(struct.pack("<II", len(jb), 0x4E4F534A) + jb)
        f.write(struct.pack("<II", len(bb), 0x004E4942) + bb)


def sample_mesh(verts, faces, num_points, seed=0, vertex_rgb=None, return_face=False):
    """Area-weighted surface samples on the GPU (genpc_mesh_face_areas + genpc_mesh_sample).  verts [V,3] f32, faces [F,3] i32
    CUDA tensors.  -> (xyz [n,3], rgb [n,3]) (+ face [n] int32)."""
    _lib.require_cuda(verts, faces)
    v, f = verts.contiguous().float(), faces.contiguous().int()
    dev, F = v.device, f.shape[0]
    if F == 0:
        raise _lib.GenpcError("sample_mesh: the mesh has no faces")
    L = _lib.lib()
    areas = torch.empty(F, dtype=torch.float32, device=dev)
    col = None if vertex_rgb is None else vertex_rgb.contiguous().float()
    xyz = torch.empty(num_points, 3, dtype=torch.float32, device=dev)
    rgb = torch.empty(num_points, 3, dtype=torch.float32, device=dev)
    face = torch.empty(num_points, dtype=torch.int32, device=dev) if return_face else None
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        _lib.check(L.genpc_mesh_face_areas(_lib.ptr(v), _lib.ptr(f), v.shape[0], F, _lib.ptr(areas), st), "genpc_mesh_face_areas")
        a64 = areas.double()
        amax = a64.max()
        if not bool(amax > 0):
            raise _lib.GenpcError("sample_mesh: the mesh has no face with a positive area")
        cum = torch.cumsum(torch.floor(a64 / amax * 4294967296.0).long(), 0).contiguous()   # exact integers (< 2^57)
        _lib.check(L.genpc_mesh_sample(_lib.ptr(v), _lib.ptr(f), _lib.ptr(col), _lib.ptr(cum), F, int(num_points),
                                       int(seed) & ((1 << 64) - 1), _lib.ptr(xyz), _lib.ptr(rgb), _lib.ptr(face), st),
                   "genpc_mesh_sample")
    return (xyz, rgb, face) if return_face else (xyz, rgb)


def glb2point(glb_path, down_sample=None, num_points=16384, seed=0, device=None):
    """utils/dataUtils.py:217-250 on the GPU: `num_points` seeded surface samples of the .glb mesh with barycentrically
    interpolated colours, optionally voxel down-sampled (points and colours averaged per voxel, Open3D semantics).
    Returns (points [n,3], colors [n,3]) float32 CUDA tensors (the reference wraps them in an o3d.geometry.PointCloud)."""
    from .dataUtils import voxel_down_sample

    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    verts, faces, col = read_glb(glb_path)
    pts, rgb = sample_mesh(torch.from_numpy(verts).to(device), torch.from_numpy(faces).to(device), num_points, seed,
                           None if col is None else torch.from_numpy(col).to(device))
    if down_sample:
        pts, rgb = voxel_down_sample(pts, float(down_sample), rgb)
    return pts, rgb
