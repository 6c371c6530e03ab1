"""CPU: the reference's import spellings resolve to this package after dropin.install(); signatures match the
reference's."""
import inspect
import subprocess
import sys

from conftest import ROOT


def test_dropin_aliases_in_a_fresh_interpreter():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import genpc_b200.dropin as d; d.install()\n"
        "from loss_functions import chamfer_3DDist, emdModule\n"
        "from loss_functions.Chamfer3D.dist_chamfer_3D import chamfer_3DFunction\n"
        "from utils.loss_util import Completionloss\n"
        "import chamfer_3D, emd\n"
        "assert chamfer_3DDist.__module__.startswith('genpc_b200')\n"
        "print(sorted(n for n in dir(chamfer_3D) if n in ('forward','backward')), sorted(n for n in dir(emd) if n in ('forward','backward')))\n"
    ) % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "['backward', 'forward'] ['backward', 'forward']" in out.stdout


def test_signatures_match_the_reference():
    from genpc_b200 import chamfer_3D, emd
    from genpc_b200.optim_registration import diff_obj_pose as dop
    from genpc_b200.utils.loss_util import Completionloss

    assert list(inspect.signature(chamfer_3D.forward).parameters) == ["xyz1", "xyz2", "dist1", "dist2", "idx1", "idx2"]
    assert list(inspect.signature(chamfer_3D.backward).parameters) == [
        "xyz1", "xyz2", "gradxyz1", "gradxyz2", "graddist1", "graddist2", "idx1", "idx2"]
    assert list(inspect.signature(emd.forward).parameters) == [
        "xyz1", "xyz2", "dist", "assignment", "price", "assignment_inv", "bid", "bid_increments", "max_increments",
        "unass_idx", "unass_cnt", "unass_cnt_sum", "cnt_tmp", "max_idx", "eps", "iters"]
    assert list(inspect.signature(emd.backward).parameters) == ["xyz1", "xyz2", "gradxyz", "graddist", "idx"]
    sig = inspect.signature(dop.object_pose_optimization)
    assert list(sig.parameters) == ["glb_path", "point_path", "radius", "lr", "iters", "render_size", "vis", "save_path",
                                    "device", "cam_bias_num"]
    assert sig.parameters["lr"].default == 0.005 and sig.parameters["iters"].default == 300
    for m in ("chamfer_l1", "chamfer_l2", "chamfer_partial_l1", "chamfer_partial_l2", "emd_loss", "get_loss"):
        assert hasattr(Completionloss, m)
    import pytest

    with pytest.raises(Exception):
        Completionloss("nope")
