"""Lattice: reference-orbit tables and the step schedule (host side, numpy).

Mirrors the public surface of the reference's ``Lattice`` (lattice.py:112-204): attributes
``coords, n_vec, tau_vec, rho, distance, nsep, min_x, delta_x, step_size, steps_per_element,
lattice_config, total_steps, steps_record, Nelement, lattice_length`` and ``update(name)``.
The tables are static per run (~100 KB); ``device_tables()`` uploads them once for the wake kernel.
"""
from __future__ import annotations

import numpy as np

from .yaml_parser import parse_yaml

N_SAMPLE = 2000   # lattice.py:136


def reference_trajectory(lattice_config, n_sample=N_SAMPLE):
    """Planar reference orbit sampled at `n_sample` equidistant s (lattice.py:4-110): in a dipole each
    sample advances on the arc about the instantaneous centre of curvature, elsewhere along the
    current tangent; n = (sin theta, -cos theta), tau = (cos theta, sin theta)."""
    names = list(lattice_config.keys())[1:]
    n_el = len(names)
    distance = np.zeros(n_el)
    rho = np.zeros(n_el)
    nsep = np.zeros(n_el)
    run = 0.0
    for k, name in enumerate(names):
        el = lattice_config[name]
        run = el["L"] if k == 0 else el["L"] + distance[k - 1]
        distance[k] = run
        nsep[k] = el["nsep"]
        if el["type"] == "dipole":
            rho[k] = el["angle"] / el["L"]
    s = np.linspace(0, distance[-1], n_sample)
    coords = np.zeros((n_sample, 2))
    tau = np.zeros((n_sample, 2))
    nrm = np.zeros((n_sample, 2))
    theta = 0
    tau[0] = (np.cos(theta), np.sin(theta))
    nrm[0] = (np.sin(theta), -1 * np.cos(theta))
    cur = 0
    for k in range(1, n_sample):
        if s[k] > distance[cur]:
            cur += 1
        ds = s[k] - s[k - 1]
        el = lattice_config[names[cur]]
        if el["type"] == "dipole":
            phi = ds / el["L"] * el["angle"]
            rad = el["L"] / el["angle"]
            cx = coords[k - 1, 0] - rad * np.sin(theta)
            cy = coords[k - 1, 1] + rad * np.cos(theta)
            coords[k] = (cx + rad * np.sin(phi + theta), cy - rad * np.cos(phi + theta))
        else:
            phi = 0
            coords[k] = (coords[k - 1, 0] + ds * np.cos(theta), coords[k - 1, 1] + ds * np.sin(theta))
        theta += phi
        tau[k] = (np.cos(theta), np.sin(theta))
        nrm[k] = (np.sin(theta), -1 * np.cos(theta))
    return s, rho, distance, nsep, coords, nrm, tau


class Lattice:
    def __init__(self, input_lattice):
        assert "lattice_input_file" in input_lattice or "lattice_config" in input_lattice, \
            "Error in parsing lattice: must include the keyword <lattice_input_file>"
        if "lattice_config" in input_lattice:          # in-memory lattice (tests, benchmarks)
            self.lattice_input_file = None
            cfg = input_lattice["lattice_config"]
        else:
            self.lattice_input_file = input_lattice["lattice_input_file"]
            cfg = parse_yaml(self.lattice_input_file)
        assert "step_size" in cfg, "Required input parameter step_size to Lattice.__init__(**kwargs) was not found."
        self.lattice_config = cfg
        self._Nelement = len(cfg) - 1
        (self.s, self.rho, self.distance, self.nsep, self.coords, self.n_vec,
         self.tau_vec) = reference_trajectory(cfg)
        self._lattice_length = self.distance[-1]
        self._schedule()
        self.min_x, self.max_x = self.s[0], self.s[-1]
        self.delta_x = (self.max_x - self.min_x) / (self.s.shape[0] - 1)
        self.current_element = None
        self._device = {}

    def _schedule(self):
        """Positions of the tracking steps and how many fall in each element (lattice.py:152-174)."""
        self.step_size = self.lattice_config["step_size"]
        self._positions_record = np.arange(0, self.lattice_length + self.step_size / 2, self.step_size)
        self._total_steps = len(self._positions_record)
        self.steps_per_element = np.zeros((self.Nelement,), dtype=int)
        csr_idx = []
        prev = 0
        for k, d in enumerate(self.distance):
            ind = int(np.searchsorted(self._positions_record, d, side="right"))
            csr_idx.append(np.arange(prev, ind, self.nsep[k]))
            self.steps_per_element[k] = ind - prev - (1 if k == 0 else 0)
            prev = ind
        self._CSR_steps_index = np.concatenate(csr_idx) if csr_idx else np.array([])
        self._CSR_steps_count = len(self._CSR_steps_index)

    lattice_length = property(lambda self: self._lattice_length)
    CSR_steps_index = property(lambda self: self._CSR_steps_index)
    total_steps = property(lambda self: self._total_steps)
    steps_record = property(lambda self: self._positions_record)
    Nelement = property(lambda self: self._Nelement)

    def update(self, ele_name):
        self.current_element = ele_name

    def device_tables(self, device):
        from . import ops
        key = str(device)
        if key not in self._device:
            self._device[key] = ops.DeviceLattice.upload(self.coords, self.n_vec, self.tau_vec, self.rho,
                                                         self.distance, self.min_x, self.delta_x, device)
        return self._device[key]
