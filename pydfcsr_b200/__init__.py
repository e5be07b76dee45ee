"""pydfcsr_b200 — B200-native CSR-wake hot path behind pyDFCSR's own call boundary.

``CSR2D``, ``DF_tracker``, ``Beam`` (and everything in ``ops``) run on ``libdfcsr_b200.so`` (hand-written sm_100a
kernels behind the C ABI of ``include/dfcsr_b200.h``).  The library is loaded the first time one of them is
touched; a missing library is an ImportError — there is no CPU fallback.  The pure-host helpers (``synth``,
``hostmaps``, ``lattice``, ``params``, ``yaml_parser``) import without it, so a process that only needs those
(the CPU reference arm of bench.py) never maps the CUDA library.
"""
__version__ = "0.2.0"

_LAZY = {"CSR2D": ".csr", "DF_tracker": ".deposit", "Beam": ".beams", "Lattice": ".lattice"}

__all__ = ["CSR2D", "DF_tracker", "Beam", "Lattice"]


def __getattr__(name):
    if name in _LAZY:
        import importlib
        value = getattr(importlib.import_module(_LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
