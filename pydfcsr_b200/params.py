"""Run-time parameter blocks of the YAML input (``CSR_integration`` and ``CSR_computation``).

Same class names, keys and default values as the reference (params.py:15-43), table-driven here: each
block is a dict of defaults; unknown keys raise ``TypeError`` exactly like an unexpected keyword argument
does in the reference's ``configure_params(**input_dic)``.
"""
from __future__ import annotations

from .yaml_parser import full_path


class _ParameterBlock:
    DEFAULTS: dict = {}

    def __init__(self, input_dic=None):
        self.configure_params(**(input_dic or {}))

    def configure_params(self, **given):
        unknown = sorted(set(given) - set(self.DEFAULTS))
        if unknown:
            raise TypeError(f"{type(self).__name__}.configure_params() got an unexpected keyword argument {unknown[0]!r}")
        for key, default in self.DEFAULTS.items():
            setattr(self, key, given.get(key, default))
        self._finalize()

    def _finalize(self):
        pass

    def as_dict(self):
        return {key: getattr(self, key) for key in self.DEFAULTS}


class Integration_params(_ParameterBlock):
    """Quadrature of the retarded-field integral: window depth in formation lengths and node counts of the
    (x', s') rectangles (CSR.py:539-553)."""
    DEFAULTS = {"n_formation_length": 4, "zbins": 200, "xbins": 200}


class CSR_params(_ParameterBlock):
    """What to compute and where: switches, the observation mesh (CSR.py:361-394) and output options."""
    DEFAULTS = {"workdir": ".", "apply_CSR": 1, "compute_CSR": 1, "transverse_on": 1,
                "xbins": 20, "zbins": 30, "xlim": 5, "zlim": 5,
                "write_beam": None, "write_wakes": True, "write_name": ""}

    def _finalize(self):
        self.workdir = full_path(self.workdir)
