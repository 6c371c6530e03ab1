"""Dependency-free IO / format glue the hot path needs (the reference's utils/dataUtils.py leans on open3d and
trimesh, which are not vendored): binary/ascii PLY xyz reader + writer, normalize_numpy, voxel down-sampling.
Not a hot path; kept device-friendly so scans can stay resident between stages."""
import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
              "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2",
              "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def read_ply_xyz(path):
    """-> (points float32 [N,3], colors float32 [N,3] in [0,1] or None).  Vertex element only."""
    with open(path, "rb") as f:
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list property on vertex element not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            arr = np.loadtxt(f, max_rows=n, ndmin=2)
            cols = {name: arr[:, i] for i, (name, _) in enumerate(props)}
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(name, end + t) for name, t in props])
            rec = np.frombuffer(f.read(dt.itemsize * n), dtype=dt, count=n)
            cols = {name: rec[name] for name, _ in props}
    pts = np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float32)
    col = None
    if all(k in cols for k in ("red", "green", "blue")):
        col = np.stack([cols["red"], cols["green"], cols["blue"]], 1).astype(np.float32)
        if col.max() > 1.0:
            col = col / 255.0
    return pts, col


def write_ply_xyz(path, points, colors=None):
    points = np.asarray(points, dtype=np.float64)
    n = points.shape[0]
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {n}", "property double x", "property double y",
           "property double z"]
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if colors is not None:
        hdr += ["property uchar red", "property uchar green", "property uchar blue"]
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    hdr.append("end_header")
    rec = np.zeros(n, dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = points[:, 0], points[:, 1], points[:, 2]
    if colors is not None:
        c = np.clip(np.asarray(colors) * 255.0, 0, 255).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = c[:, 0], c[:, 1], c[:, 2]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        f.write(rec.tobytes())


def voxel_down_sample(points, voxel_size, colors=None):
    """Open3D semantics: one output point per occupied voxel = mean of its points (voxel grid anchored at
    min_bound - voxel/2); colours, when given, are averaged the same way.  torch, runs on whatever device `points`
    lives on; output order = sorted voxel key.  -> points, or (points, colors) when colors is given."""
    p = torch.as_tensor(points)
    lo = p.min(0).values - voxel_size * 0.5
    key = torch.floor((p - lo) / voxel_size).long()
    uniq, inv = torch.unique(key, dim=0, return_inverse=True)
    cnt = torch.zeros(uniq.shape[0], dtype=p.dtype, device=p.device).index_add_(0, inv, torch.ones_like(p[:, 0]))
    out = torch.zeros(uniq.shape[0], 3, dtype=p.dtype, device=p.device).index_add_(0, inv, p) / cnt[:, None]
    if colors is None:
        return out
    c = torch.as_tensor(colors).to(p.dtype)
    return out, torch.zeros(uniq.shape[0], c.shape[1], dtype=p.dtype, device=p.device).index_add_(0, inv, c) / cnt[:, None]


def load_xyz(path, down_sample=None):
    """reference utils/dataUtils.py:174-189: (points f32 [N,3], colors); colourless files get per-axis min-max
    normalised coordinates as colours."""
    pts, col = read_ply_xyz(path)
    if down_sample:
        pts = voxel_down_sample(torch.from_numpy(pts), float(down_sample)).numpy().astype(np.float32)
        col = None
    if col is None or np.allclose(col, 0):
        col = (pts - pts.min(axis=0)) / (pts.max(axis=0) - pts.min(axis=0) + 1e-8)
        col = np.clip(col, 0, 1)
    return pts, col.astype(np.float32)


def normalize_numpy(xyz, range=1.0):
    """reference utils/dataUtils.py:561-581."""
    vmin, vmax = xyz.min(axis=0), xyz.max(axis=0)
    center = (vmax + vmin) / 2.0
    scale_factor = (vmax - vmin).max()
    out = (xyz - center) / scale_factor
    out = out * (range / 0.5)
    return out, center, scale_factor
