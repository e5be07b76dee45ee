"""Shared seeded scenarios for the parity tests: a synthetic bunch pushed through the first cells of
the chicane with the stand-in linear tracker, logged into the ORACLE's history (CPU).  The same
particle arrays are then fed to the CUDA path."""
from __future__ import annotations

import functools

import numpy as np

from oracle import dfcsr_oracle as O
from pydfcsr_b200 import synth, tracking

DEPOSIT_CFG = dict(xbins=64, zbins=96, xlim=5, zlim=5, filter_order=1, filter_window=9,
                   velocity_threhold=1000, upper_limit=2000)
R_BEND = 0.5002 / 0.0483


def lattice_tables() -> O.LatticeTables:
    return O.reference_orbit([(e[1], e[2], e[3]) for e in synth.CHICANE_ELEMENTS])


def formation_length(sigma_z: float) -> float:
    return (24 * (R_BEND ** 2) * 5 * sigma_z) ** (1 / 3)      # CSR.py:140-142,256


@functools.lru_cache(maxsize=8)
def chicane_entry(n_particle=100_000, seed=1, tilt=0.0, n_bend_steps=4, deposit_items=tuple(DEPOSIT_CFG.items()),
                  upper_limit_none=False):
    """Drift 0.1 m then `n_bend_steps` x 0.1 m inside B1.  Returns a dict with per-step particle
    coordinates, the oracle's density functions, history stack and wake scalars at the last step."""
    dep = dict(deposit_items)
    if upper_limit_none:
        dep["upper_limit"] = None
    cfg = O.DepositConfig(**dep)
    hist = O.HistoryOracle(cfg)
    coords = tuple(synth.gaussian_bunch(n_particle, seed=seed, tilt=tilt))
    pos = 0.0
    steps = []

    def log(coords, pos, fl):
        x, px, y, py, z, pz = coords
        df = O.make_density_functions(x, z, px, pos, cfg)
        hist.append(df)
        rebuilt = hist.push(fl, 1)
        steps.append(dict(coords=coords, pos=pos, formation_length=fl, df=df, rebuilt=rebuilt))

    log(coords, 0, float("inf"))                      # CSR2D.initialization (CSR.py:75-78)
    coords = tracking.track_linear(coords, tracking.Drift(0.1))
    pos += 0.1
    fl = 0.1                                          # first drift: formation_length += L (CSR.py:269)
    log(coords, pos, fl)
    for _ in range(n_bend_steps):
        coords = tracking.track_linear(coords, tracking.SBend(L=0.1, G=0.0483 / 0.5002, FRINGE_AT="no_end"))
        pos += 0.1
        fl = formation_length(float(np.std(coords[4])))
        log(coords, pos, fl)
    x, px, y, py, z, pz = coords
    sc = O.beam_scalars(x, z)
    wake_sc = dict(t=pos, sigma_x=float(sc["sigma_x"]), sigma_z=float(sc["sigma_z"]), slope0=float(sc["slope"][0]),
                   mean_x=float(sc["mean_x"]), formation_window=1 * fl, csr_scaling=8.98755e3 * 1.0e-9)
    return dict(steps=steps, history=hist, stack=hist.stack(), coords=coords, pos=pos, scalars=sc,
                wake_scalars=wake_sc, lattice=lattice_tables(), deposit_cfg=cfg)
