"""CPU: pin the oracle on outputs of the UNMODIFIED reference extensions (tests/golden/*.npz, produced on a
B200 by tests/golden/make_golden.py from oracle/_ref).  Chamfer: dist/idx bit-exact.  EMD: assignment and
dist bit-exact (the reference was run 3x per case and was run-to-run identical on all of them)."""
import glob
import os

import numpy as np
import pytest

import oracle

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("f", sorted(glob.glob(os.path.join(G, "chamfer_ref_*.npz"))))
def test_chamfer_oracle_equals_reference_output(f):
    z = np.load(f)
    got = oracle.chamfer_forward(z["xyz1"], z["xyz2"])
    for g, name in zip(got, ("dist1", "dist2", "idx1", "idx2")):
        assert np.array_equal(g.view(np.int32), z[name].view(np.int32)), name


@pytest.mark.parametrize("f", sorted(glob.glob(os.path.join(G, "emd_ref_*.npz"))))
def test_emd_oracle_equals_reference_output(f):
    z = np.load(f)
    assert bool(z["reproducible"])
    d, a = oracle.emd_forward(z["xyz1"], z["xyz2"], float(z["eps"]), int(z["iters"]))
    assert np.array_equal(a, z["assignment"])
    assert np.array_equal(d.view(np.int32), z["dist"].view(np.int32))


def test_golden_files_present():
    assert len(glob.glob(os.path.join(G, "chamfer_ref_*.npz"))) >= 3
    assert len(glob.glob(os.path.join(G, "emd_ref_*.npz"))) >= 3
