"""Parameter objects with the reference's names and defaults (params.py:15-43)."""
from __future__ import annotations

from .yaml_parser import full_path


class Integration_params:
    def __init__(self, input_dic=None):
        self.configure_params(**(input_dic or {}))

    def configure_params(self, n_formation_length=4, zbins=200, xbins=200):
        self.n_formation_length = n_formation_length
        self.zbins = zbins
        self.xbins = xbins


class CSR_params:
    def __init__(self, input_dic=None):
        self.configure_params(**(input_dic or {}))

    def configure_params(self, workdir=".", apply_CSR=1, compute_CSR=1, transverse_on=1, xbins=20, zbins=30,
                         xlim=5, zlim=5, write_beam=None, write_wakes=True, write_name=""):
        self.compute_CSR = compute_CSR
        self.apply_CSR = apply_CSR
        self.transverse_on = transverse_on
        self.xbins = xbins
        self.zbins = zbins
        self.xlim = xlim
        self.zlim = zlim
        self.write_beam = write_beam
        self.write_wakes = write_wakes
        self.workdir = full_path(workdir)
        self.write_name = write_name
