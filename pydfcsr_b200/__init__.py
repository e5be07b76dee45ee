"""pydfcsr_b200 — B200-native CSR-wake hot path behind pyDFCSR's own call boundary.

Importing the package loads ``libdfcsr_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/dfcsr_b200.h``); a missing library is an ImportError — there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)
from .csr import CSR2D  # noqa: E402
from .deposit import DF_tracker  # noqa: E402
from .beams import Beam  # noqa: E402
from .lattice import Lattice  # noqa: E402

__all__ = ["CSR2D", "DF_tracker", "Beam", "Lattice"]
