"""CPU: pin the oracle on outputs of the UNMODIFIED reference extensions (tests/golden/*.npz, produced on a
B200 by tests/golden/make_golden.py from oracle/_ref).  Chamfer: dist/idx bit-exact.  EMD: assignment and
dist bit-exact (the reference was run 3x per case and was run-to-run identical on all of them), under the resolution
of its GetMax race that the file pins."""
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


# The reference's GetMax is a plain-store race (emd_cuda.cu:188-191): when several bidders fall inside the +-1e-6 window of one
# target, the winner is whichever block stores last.  Goldens without a decisive collision pin the oracle bit for bit under
# BOTH resolutions; emd_ref_centered.npz is a witness of the race -- the reference's (run-to-run identical) output equals
# the "lowest index" resolution there, while live comparisons at n = 16384 land on "highest index" (DESIGN.md section 2).
RACE_WITNESS = {"emd_ref_centered.npz": True}   # file -> getmax_lowest that reproduces it


@pytest.mark.parametrize("f", sorted(glob.glob(os.path.join(G, "emd_ref_*.npz"))))
def test_emd_oracle_equals_reference_output(f):
    z = np.load(f)
    assert bool(z["reproducible"])
    name = os.path.basename(f)
    rules = [RACE_WITNESS[name]] if name in RACE_WITNESS else [False, True]
    for lowest in rules:
        d, a = oracle.emd_forward(z["xyz1"], z["xyz2"], float(z["eps"]), int(z["iters"]), getmax_lowest=lowest)
        assert np.array_equal(a, z["assignment"]), (name, lowest)
        assert np.array_equal(d.view(np.int32), z["dist"].view(np.int32)), (name, lowest)


def test_emd_race_witness_differs_under_the_other_resolution():
    z = np.load(os.path.join(G, "emd_ref_centered.npz"))
    d, a = oracle.emd_forward(z["xyz1"], z["xyz2"], float(z["eps"]), int(z["iters"]), getmax_lowest=False)
    assert 0 < int((a != z["assignment"]).sum()) < a.size // 10      # 44 of 1024 when this was written


def test_golden_files_present():
    assert len(glob.glob(os.path.join(G, "chamfer_ref_*.npz"))) >= 6
    assert len(glob.glob(os.path.join(G, "emd_ref_*.npz"))) >= 5
