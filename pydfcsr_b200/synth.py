"""Synthetic Gaussian bunches for tests and benchmarks (host side, numpy).

``distgen`` (the reference's beam generator, beams.py:38-46) is not available offline, so the
bench and the tests use a seeded 6-D Gaussian modelled on the reference's bundled chicane beam
(example/input/chicane_init_beam.yaml): 5 GeV, 1 nC, sigma_x = 63.937 um, sigma_z = 200 um,
Twiss alpha = 2.6 / beta = 40 m / emittance 0.102 nm in x (beta = 13 m in y), energy chirp
-36 / m with 2e-6 uncorrelated spread.  Coordinates follow Bmad-X: (x, px, y, py, z, pz), z > 0 is
the bunch head and pz is the relative momentum deviation.
"""
from __future__ import annotations

import numpy as np

CHICANE_BEAM = dict(energy=5.0e9, charge=1.0e-9, sigma_x=63.937e-6, sigma_z=200.0e-6,
                    alpha_x=2.6, beta_x=40.0, alpha_y=2.6, beta_y=13.0, emittance=0.102e-9,
                    chirp=-36.0, sigma_delta=2.0e-6)


def gaussian_bunch(n_particle: int, seed: int = 0, tilt: float = 0.0, modulation: float = 0.0,
                   modulation_wavelength_sigma: float = 0.05, sigma_z: float | None = None,
                   **overrides) -> np.ndarray:
    """Return a (6, n) float64 array [x, px, y, py, z, pz].

    tilt        adds ``x += tilt * z`` (a tilt > ~1 switches the reference to its chirp-band
                quadrature, CSR.py:480-491, and to the YAML deposit grid, deposit.py:160-163).
    modulation  relative depth of a longitudinal density modulation applied by accept/reject-free
                phase-space displacement z += A sin(kz)/k (micro-bunched beam of BASELINE config 5).
    """
    p = dict(CHICANE_BEAM)
    p.update(overrides)
    if sigma_z is not None:
        p["sigma_z"] = sigma_z
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((6, n_particle))
    z = p["sigma_z"] * g[4]
    if modulation:
        k = 2.0 * np.pi / (modulation_wavelength_sigma * p["sigma_z"])
        z = z + modulation * np.sin(k * z) / k
    ex = p["emittance"]
    x = np.sqrt(ex * p["beta_x"]) * g[0]
    px = -p["alpha_x"] / p["beta_x"] * x + np.sqrt(ex / p["beta_x"]) * g[1]
    y = np.sqrt(ex * p["beta_y"]) * g[2]
    py = -p["alpha_y"] / p["beta_y"] * y + np.sqrt(ex / p["beta_y"]) * g[3]
    pz = p["chirp"] * z + p["sigma_delta"] * g[5]
    if tilt:
        x = x + tilt * z
    return np.ascontiguousarray(np.stack([x, px, y, py, z, pz]))


# The reference's bundled 4-dipole chicane (example/input/chicane_lattice.yaml), restated as data:
# (name, type, L [m], angle [rad], E1, E2, nsep); step_size 0.1 m.
CHICANE_STEP = 0.1
CHICANE_ELEMENTS = [
    ("D0", "drift", 0.1, 0.0, 0.0, 0.0, 1),
    ("B1", "dipole", 0.5002, 0.0483, 0.0, 0.0483, 1),
    ("D1", "drift", 5.0058, 0.0, 0.0, 0.0, 5),
    ("B2", "dipole", 0.5002, -0.0483, -0.0483, 0.0, 1),
    ("D2", "drift", 1.0, 0.0, 0.0, 0.0, 1),
    ("B3", "dipole", 0.5002, -0.0483, 0.0, -0.0483, 1),
    ("D3", "drift", 5.0058, 0.0, 0.0, 0.0, 5),
    ("B4", "dipole", 0.5002, 0.0483, 0.0483, 0.0, 1),
    ("Df", "drift", 0.2, 0.0, 0.0, 0.0, 1),
]


def chicane_lattice_config(step_size: float = CHICANE_STEP, elements=None) -> dict:
    """Lattice dictionary in the reference's YAML schema (lattice.py:118-135)."""
    cfg = {"step_size": step_size}
    for name, kind, length, angle, e1, e2, nsep in (elements or CHICANE_ELEMENTS):
        ent = {"type": kind, "L": length, "nsep": nsep}
        if kind == "dipole":
            ent.update(angle=angle, E1=e1, E2=e2)
        cfg[name] = ent
    return cfg
