"""Drop-in for the reference's `loss_functions` package (loss_functions/__init__.py:2-3)."""
from .Chamfer3D.dist_chamfer_3D import chamfer_3DDist, chamfer_3DFunction  # noqa: F401
from .emd.emd_module import emdFunction, emdModule  # noqa: F401
