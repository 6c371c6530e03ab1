"""BASELINE config C1 fixture: the float32 xyz of the reference's own scan `data/01184.ply` (71 372 points, stored there
as binary little-endian doubles) + the 16 384-point synthetic shape it is compared with.

    python tests/golden/make_c1_fixture.py            # here (needs /root/reference): writes tests/golden/scan_01184_xyz.npz
    gpurun -- python tests/golden/make_c1_fixture.py --golden   # on a B200: adds the reference extension's answer
                                                                # (gpurun_out/golden/c1_ref_chamfer.npz -> tests/golden/)

The PLY is parsed here by an independent 10-line reader (header scan + np.frombuffer), NOT by the product's
read_ply_xyz, so that tests/test_c1_scan.py can check the product reader against it.
The golden keeps idx1/idx2 in full (int16 / int32) and the distances as float32, compressed.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SCAN = "/root/reference/data/01184.ply"
FIX = os.path.join(ROOT, "tests", "golden", "scan_01184_xyz.npz")


def parse_ply_doubles(path):
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    hdr = raw[:end].decode("ascii").split("\n")
    assert "format binary_little_endian 1.0" in hdr
    n = int([h for h in hdr if h.startswith("element vertex")][0].split()[2])
    props = [h.split()[1:] for h in hdr if h.startswith("property")]
    assert [p[1] for p in props[:3]] == ["x", "y", "z"] and all(p[0] == "double" for p in props[:3])
    stride = sum({"double": 8, "float": 4, "uchar": 1}[p[0]] for p in props)
    rec = np.frombuffer(raw[end:end + n * stride], dtype=np.uint8).reshape(n, stride)
    return np.ascontiguousarray(rec[:, :24]).view("<f8").reshape(n, 3).astype(np.float32)


def c1_shape():
    from genpc_b200.synthetic import superquadric

    return superquadric(0, 16384)     # seed 0, normalize_numpy(range=0.5) frame


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    if "--golden" in sys.argv:
        import torch

        import oracle

        scan = np.load(FIX)["xyz"]
        shape = c1_shape()
        ext = oracle.load_ref_ext("chamfer_3D")
        dev = torch.device("cuda:0")
        a, b = torch.from_numpy(scan[None]).to(dev), torch.from_numpy(shape[None]).to(dev)
        d1 = torch.zeros(1, a.shape[1], device=dev); d2 = torch.zeros(1, b.shape[1], device=dev)
        i1 = torch.zeros(1, a.shape[1], dtype=torch.int32, device=dev); i2 = torch.zeros(1, b.shape[1], dtype=torch.int32, device=dev)
        ext.forward(a, b, d1, d2, i1, i2)
        torch.cuda.synchronize()
        out = os.path.join(ROOT, "gpurun_out", "golden")
        os.makedirs(out, exist_ok=True)
        d1, d2, i1, i2 = (t.cpu().numpy()[0] for t in (d1, d2, i1, i2))
        assert i1.max() < 32768
        np.savez_compressed(os.path.join(out, "c1_ref_chamfer.npz"), dist1=d1, dist2=d2, idx1=i1.astype(np.int16), idx2=i2,
                            sha256=np.array(digest(d1, d2, i1, i2)), shape_sha256=np.array(digest(shape)))
        print("c1_ref_chamfer.npz", digest(d1, d2, i1, i2))
        return
    xyz = parse_ply_doubles(SCAN)
    np.savez_compressed(FIX, xyz=xyz, sha256=np.array(digest(xyz)))
    print(FIX, xyz.shape, digest(xyz), os.path.getsize(FIX))


if __name__ == "__main__":
    main()
