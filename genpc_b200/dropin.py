"""Zero-edit drop-in: make the reference's import names resolve to this package.

    import genpc_b200.dropin; genpc_b200.dropin.install()
    from loss_functions import chamfer_3DDist, emdModule          # reference spelling, B200 kernels
    from utils.loss_util import Completionloss
    import chamfer_3D, emd                                        # the pybind module names (chamfer_cuda.cpp:30, emd.cpp:25)

Only the hot-path modules are aliased; every other reference module (`utils.dataUtils`, `tools.*` ...) keeps
resolving to the reference's own files.
"""
import sys


def install():
    from . import DepthPrompting, chamfer_3D, emd, loss_functions
    from .loss_functions.Chamfer3D import dist_chamfer_3D
    from .loss_functions.emd import emd_module
    from .optim_registration import diff_obj_pose
    from .utils import loss_util

    alias = {
        "chamfer_3D": chamfer_3D,
        "emd": emd,
        "loss_functions": loss_functions,
        "loss_functions.Chamfer3D": sys.modules[dist_chamfer_3D.__package__],
        "loss_functions.Chamfer3D.dist_chamfer_3D": dist_chamfer_3D,
        "loss_functions.emd": sys.modules[emd_module.__package__],
        "loss_functions.emd.emd_module": emd_module,
        "utils.loss_util": loss_util,
        "optim_registration.diff_obj_pose": diff_obj_pose,
        "DepthPrompting": DepthPrompting,
    }
    sys.modules.update(alias)
    return sorted(alias)
