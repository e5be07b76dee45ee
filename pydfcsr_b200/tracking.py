"""First-order stand-in for the Bmad-X ``track_element`` call of the reference's tracking loop
(beams.py:101-106, CSR.py:146-199).

``bmadx`` is not installable offline.  When it is importable the driver uses it unchanged; when it
is not, these linear maps (drift, sector bend with pole-face rotation, thin-lens-free thick
quadrupole, drift-like sextupole) keep the lattice loop runnable.  They act on any array type with
numpy semantics (numpy arrays on the host, torch tensors on the device), so a device-resident beam
never leaves HBM between the kick and the next deposit.

Coordinates: Bmad-X canonical (x, px, y, py, z, pz); z > 0 is the head; pz = delta.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

try:  # the reference's tracker (beams.py:8, CSR.py:3); works on torch tensors, so the beam stays on the device
    if os.environ.get("DFCSR_USE_BMADX", "1") != "1":
        raise ImportError("disabled")
    import bmadx as _bmadx
    HAVE_BMADX = True
except Exception:  # not installable offline: fall back to the first-order maps below
    _bmadx = None
    HAVE_BMADX = False


@dataclass
class Drift:
    L: float


@dataclass
class SBend:
    L: float
    G: float = 0.0            # curvature 1/rho (angle / L)
    E1: float = 0.0
    E2: float = 0.0
    FRINGE_AT: str = "both_ends"
    P0C: float = 0.0


@dataclass
class Quadrupole:
    L: float
    K1: float = 0.0


@dataclass
class Sextupole:
    L: float
    K2: float = 0.0


def _edge(x, px, y, py, g, e):
    if e == 0.0 or g == 0.0:
        return px, py
    k = g * math.tan(e)
    return px + k * x, py - k * y


def linear_matrix(element):
    """The 6 x 6 transfer matrix of `element`: the maps below applied to the unit vectors."""
    import numpy as np
    cols = _apply_map(tuple(np.eye(6)), element)          # row k of eye = coordinate k of the six unit particles
    return np.stack([np.asarray(c, dtype=np.float64) for c in cols])


def track_linear(coords, element):
    """coords: sequence (x, px, y, py, z, pz) of equally shaped arrays; returns the same.
    CUDA tensors are transformed in place by one kernel launch (ops.track_linear); host arrays by the
    expressions in `_apply_map`."""
    if getattr(coords[0], "is_cuda", False):
        from . import ops
        coords = tuple(c if c.is_contiguous() else c.contiguous() for c in coords)
        ops.track_linear(coords, linear_matrix(element))
        return coords
    return _apply_map(coords, element)


def _apply_map(coords, element):
    x, px, y, py, z, pz = coords
    L = element.L
    if isinstance(element, SBend) and element.G != 0.0:
        g = element.G
        th = g * L
        c, s = math.cos(th), math.sin(th)
        if element.FRINGE_AT in ("both_ends", "entrance_end"):
            px, py = _edge(x, px, y, py, g, element.E1)
        x1 = c * x + (s / g) * px + ((1.0 - c) / g) * pz
        px1 = -(g * s) * x + c * px + s * pz
        z1 = z - s * x - ((1.0 - c) / g) * px - ((th - s) / g) * pz
        y1 = y + L * py
        x, px, y, z = x1, px1, y1, z1
        if element.FRINGE_AT in ("both_ends", "exit_end"):
            px, py = _edge(x, px, y, py, g, element.E2)
        return x, px, y, py, z, pz
    if isinstance(element, Quadrupole) and element.K1 != 0.0:
        k = element.K1
        w = math.sqrt(abs(k))
        cf, sf = math.cos(w * L), math.sin(w * L)
        cd, sd = math.cosh(w * L), math.sinh(w * L)
        if k > 0:
            x, px = cf * x + (sf / w) * px, -(w * sf) * x + cf * px
            y, py = cd * y + (sd / w) * py, (w * sd) * y + cd * py
        else:
            x, px = cd * x + (sd / w) * px, (w * sd) * x + cd * px
            y, py = cf * y + (sf / w) * py, -(w * sf) * y + cf * py
        return x, px, y, py, z, pz
    return x + L * px, px, y + L * py, py, z, pz


def make_element(kind: str, L: float, p0c: float = 0.0, **kw):
    """Element factory used by CSR2D.get_bmadx_element (CSR.py:146-199): Bmad-X elements when the package is
    importable, the stand-ins above otherwise.  kind in {'drift', 'dipole', 'quad', 'sextupole'}."""
    if HAVE_BMADX:
        if kind == "dipole":
            return _bmadx.SBend(L=L, P0C=p0c, G=kw.get("G", 0.0), E1=kw.get("E1", 0.0), E2=kw.get("E2", 0.0),
                                FRINGE_AT=kw.get("FRINGE_AT", "both_ends"))
        if kind == "quad":
            return _bmadx.Quadrupole(L=L, K1=kw["K1"])
        if kind == "sextupole":
            return _bmadx.Sextupole(L=L, K2=kw["K2"])
        return _bmadx.Drift(L=L)
    if kind == "dipole":
        return SBend(L=L, G=kw.get("G", 0.0), E1=kw.get("E1", 0.0), E2=kw.get("E2", 0.0),
                     FRINGE_AT=kw.get("FRINGE_AT", "both_ends"), P0C=p0c)
    if kind == "quad":
        return Quadrupole(L=L, K1=kw["K1"])
    if kind == "sextupole":
        return Sextupole(L=L, K2=kw["K2"])
    return Drift(L=L)


def track(coords, element, s=0.0, p0c=0.0, mc2=0.51099895e6):
    """beams.py:101-102: track_element(particle, element).  coords = (x, px, y, py, z, pz)."""
    if HAVE_BMADX and not isinstance(element, (Drift, SBend, Quadrupole, Sextupole)):
        part = _bmadx.track_element(_bmadx.Particle(*coords, s, p0c, mc2), element)
        return part.x, part.px, part.y, part.py, part.z, part.pz
    return track_linear(coords, element)
