"""Drop-in import name: put this directory's parent (``dropin/``) on PYTHONPATH *instead of* the reference and

    import pynmap
    from pynmap.misc_functions import gausshermite
    snap = pynmap.pynmap.snapshot(...)

resolve to pynmap_b200 (the B200 implementation) with the module layout of the reference's ``pynmap/__init__.py:61``
(``from . import pynmap, grid, misc_io, misc_functions, losvd``) plus ``rotation`` and ``fit_losvd``.  It lives in its own
directory so that it never shadows a real reference installation by accident (the parity tests import both)."""
import sys

import pynmap_b200 as _impl
from pynmap_b200 import fit_losvd, grid, losvd, misc_functions, misc_io, pynmap, rotation  # noqa: F401
from pynmap_b200 import LOSVD, SnapMap, snapshot  # noqa: F401

__version__ = _impl.__version__
for _name in ("pynmap", "grid", "misc_io", "misc_functions", "losvd", "rotation", "fit_losvd"):
    sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
