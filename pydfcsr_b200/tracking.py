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
from dataclasses import dataclass


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


def track_linear(coords, element):
    """coords: sequence (x, px, y, py, z, pz) of equally shaped arrays; returns the same."""
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
