"""pynmap_b200 -- B200-native (sm_100a) implementation of pynmap's particle-projection hot
path behind pynmap's own call surface:

    rotation.rotate_mat / rotateXYZ_mat        pynmap/rotation.py:14-119
    grid.points_2vmaps / points_2map / points_2losvds   pynmap/grid.py:70-262
    losvd.LOSVD(...).convolve_spatial / vmoments / fitlosvd   pynmap/losvd.py:16-141
    fit_losvd.fitgh                                      pynmap/fit_losvd.py:57-235
    pynmap.snapshot.rotate / velocity_moments / create_map / create_losvds   pynmap/pynmap.py:246-455

Import surface mirrors ``pynmap/__init__.py:61`` (``from . import pynmap, grid, misc_io, misc_functions, losvd``);
``dropin/pynmap`` re-exports it under the name ``pynmap`` itself.
All arithmetic runs in hand-written CUDA (libpnm_b200.so, C ABI in include/pnm_b200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib, engine, rotation, losvd, grid, pynmap, hostpath, sharded, ssp, fit_losvd, amr  # noqa: F401
from . import misc_io, misc_functions  # noqa: F401
from .pynmap import snapshot, SnapMap  # noqa: F401
from .losvd import LOSVD  # noqa: F401
