"""`dmip` — importable name of the package that lives in
`diffusion-modelling-for-inverse-problems_b200/` (a directory name Python cannot import directly).

    from dmip.models.diffusion import CDE, CDiffE, PosteriorDiffusionEstimator
    from dmip import sdes, nets, losses

mirror the reference's `models/diffusion.py`, `sdes.py`, `nets.py`, `losses.py` (same class names,
constructor arguments, attributes and return types); their hot methods call the sm_100a kernels of
`libdmip_sm100.so` through the C ABI in `include/dmip.h`.
"""
import os as _os

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "diffusion-modelling-for-inverse-problems_b200")
__path__.append(_pkg)

from . import _lib  # noqa: E402,F401
from ._lib import build, library_path, is_available  # noqa: E402,F401
