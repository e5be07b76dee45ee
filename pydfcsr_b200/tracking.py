"""Particle transport for the tracking loop of the reference (beams.py:101-106, CSR.py:146-199):
``self.particle = track_element(self.particle, element)`` with Bmad-X elements.

``bmadx`` is a third-party package that cannot be installed offline.  When it is importable the driver uses it
unchanged.  Otherwise the same element types (Drift, SBend with the FRINGE_AT variants of the step splitting,
Quadrupole, Sextupole) are transported by a restatement of Bmad's maps — exact drift, sector bend = "linear_edge"
hard-edge kicks + the exact body solution, thick quadrupole with quad_mat2_calc and low_energy_z_correction — which
reproduces the three Bmad-X known answers the reference holds (test/test_BmadX_tracking.ipynb cells 25, 28, 31) to
3e-13.  CUDA tensors are transported in place by one kernel (``dfcsr_track_element``, csrc/track.cu); host arrays by
the numpy expressions below (same formulas), so a device-resident beam never leaves HBM between kick and deposit.
``order="first"`` selects the first-order transfer matrices instead (``dfcsr_track_linear``).

This module imports without the CUDA library (the CPU reference arm of bench.py uses the host maps).

Coordinates: Bmad-X canonical (x, px, y, py, z, pz); z > 0 is the head; pz = delta.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

try:  # the reference's tracker (beams.py:8, CSR.py:3); works on torch tensors, so the beam stays on the device
    if os.environ.get("DFCSR_USE_BMADX", "1") != "1":
        raise ImportError("disabled")
    import bmadx as _bmadx
    HAVE_BMADX = True
except Exception:  # not installable offline: the restated maps below
    _bmadx = None
    HAVE_BMADX = False

M_ELECTRON = 0.51099895e6      # eV (bmadx.M_ELECTRON; physical_constants.py)
KIND = {"drift": 0, "sbend": 1, "quadrupole": 2, "sextupole": 3}


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
    NUM_STEPS: int = 1


@dataclass
class Sextupole:
    L: float
    K2: float = 0.0


# ------------------------------------------------------------------------------- exact maps (host, numpy)
def _sqrt_one(x):
    return x / (np.sqrt(1.0 + x) + 1.0)          # sqrt(1 + x) - 1 without cancellation


def _asin_over(u):
    u = np.asarray(u, dtype=np.float64)
    small = np.abs(u) < 1e-4
    u2 = u * u
    series = 1.0 + u2 * (1.0 / 6.0 + u2 * (3.0 / 40.0 + u2 * (15.0 / 336.0)))
    with np.errstate(divide="ignore", invalid="ignore"):
        full = np.arcsin(u) / np.where(small, 1.0, u)
    return np.where(small, series, full)


def _beta_ratio_minus_one(pz, p0c, mc2):
    P = 1.0 + pz
    return _sqrt_one((mc2 ** 2 * (2.0 * pz + pz ** 2)) / ((p0c * P) ** 2 + mc2 ** 2))


def _drift(x, px, y, py, z, pz, L, p0c, mc2):
    P = 1.0 + pz
    Px, Py = px / P, py / P
    Pxy2 = Px * Px + Py * Py
    Pl = np.sqrt(1.0 - Pxy2)
    return (x + L * Px / Pl, px, y + L * Py / Pl, py,
            z + L * (_beta_ratio_minus_one(pz, p0c, mc2) + _sqrt_one(-Pxy2) / Pl), pz)


def _bend_body(x, px, y, py, z, pz, L, g, p0c, mc2):
    th = g * L
    ct, st = math.cos(th), math.sin(th)
    if abs(th) < 1e-7:
        S, C = L * (1.0 - th * th / 6.0), 0.5 * g * L * L
    else:
        S, C = st / g, 2.0 * math.sin(0.5 * th) ** 2 / g
    P = 1.0 + pz
    pt2 = P * P - py * py
    ps = np.sqrt(pt2 - px * px)
    a = ps - 1.0 - g * x
    dpx_g = -px * C + a * S
    pxf = px + g * dpx_g
    psf = np.sqrt(pt2 - pxf * pxf)
    dps_g = -(pxf + px) / (psf + ps) * dpx_g
    xf = x * ct + (ps - 1.0) * C + px * S + dps_g
    D = (px * dps_g - ps * dpx_g) / pt2
    w = D * _asin_over(D * g)
    return xf, pxf, y + py * (L + w), py, z + L * _beta_ratio_minus_one(pz, p0c, mc2) - pz * L - P * w, pz


def _quad_mat2(k1, length, rel_p):
    sqrt_k = np.sqrt(np.abs(k1) + 2.220446049250313e-16)
    sk_l = sqrt_k * length
    pos = k1 > 0
    cx = np.where(pos, np.cosh(sk_l), np.cos(sk_l))
    sx = np.where(pos, np.sinh(sk_l), np.sin(sk_l)) / sqrt_k
    return (cx, sx / rel_p, k1 * sx * rel_p,
            k1 * (-cx * sx + length) / 4.0, -k1 * sx * sx / (2.0 * rel_p), -(cx * sx + length) / (4.0 * rel_p ** 2))


def _low_energy_z_correction(pz, p0c, mc2, ds):
    e_tot = math.sqrt(p0c ** 2 + mc2 ** 2)
    beta0 = p0c / e_tot
    ev = mc2 * (beta0 * pz) ** 2
    m2, b2 = (mc2 / e_tot) ** 2, beta0 ** 2
    series = ds * pz * (1.0 - 3.0 * (pz * b2) / 2.0 + pz ** 2 * b2 * (2.0 * b2 - m2 / 2.0)) * m2
    pc = (1.0 + pz) * p0c
    exact = ds * (pc / np.sqrt(pc ** 2 + mc2 ** 2) - beta0) / beta0
    return np.where(ev < 3e-7 * e_tot, series, exact)


def _fringes(element):
    return (element.FRINGE_AT in ("both_ends", "entrance_end"), element.FRINGE_AT in ("both_ends", "exit_end"))


def _apply_exact(coords, element, p0c, mc2):
    x, px, y, py, z, pz = (np.asarray(c, dtype=np.float64) for c in coords)
    L = element.L
    if isinstance(element, SBend) and element.G != 0.0:
        g = element.G
        f_in, f_out = _fringes(element)
        if f_in:
            t1 = g * math.tan(element.E1)
            px, py = px + t1 * x, py - t1 * y
        x, px, y, py, z, pz = _bend_body(x, px, y, py, z, pz, L, g, p0c, mc2)
        if f_out:
            t2 = g * math.tan(element.E2)
            px, py = px + t2 * x, py - t2 * y
        return x, px, y, py, z, pz
    if isinstance(element, Quadrupole) and element.K1 != 0.0:
        n_step = max(int(element.NUM_STEPS), 1)
        step = L / n_step
        for _ in range(n_step):
            rel_p = 1.0 + pz
            k1 = element.K1 / rel_p
            a11, a12, a21, c1, c2, c3 = _quad_mat2(-k1, step, rel_p)
            b11, b12, b21, d1, d2, d3 = _quad_mat2(k1, step, rel_p)
            z = z + c1 * x * x + c2 * x * px + c3 * px * px + d1 * y * y + d2 * y * py + d3 * py * py
            x, px = a11 * x + a12 * px, a21 * x + a11 * px
            y, py = b11 * y + b12 * py, b21 * y + b11 * py
            z = z + _low_energy_z_correction(pz, p0c, mc2, step)
        return x, px, y, py, z, pz
    if isinstance(element, Sextupole) and element.K2 != 0.0:
        x, px, y, py, z, pz = _drift(x, px, y, py, z, pz, 0.5 * L, p0c, mc2)
        kl = element.K2 * L
        px, py = px - 0.5 * kl * (x * x - y * y), py + kl * x * y
        return _drift(x, px, y, py, z, pz, 0.5 * L, p0c, mc2)
    return _drift(x, px, y, py, z, pz, L, p0c, mc2)


def _device_element(element):
    """dfcsr_element for the kernel (pydfcsr_b200._lib.Element)."""
    from . import _lib
    if isinstance(element, SBend) and element.G != 0.0:
        f_in, f_out = _fringes(element)
        return _lib.Element(KIND["sbend"], int(f_in), int(f_out), 1, element.L, element.G, element.E1, element.E2, 0.0, 0.0)
    if isinstance(element, Quadrupole) and element.K1 != 0.0:
        return _lib.Element(KIND["quadrupole"], 0, 0, max(int(element.NUM_STEPS), 1), element.L, 0.0, 0.0, 0.0, element.K1, 0.0)
    if isinstance(element, Sextupole) and element.K2 != 0.0:
        return _lib.Element(KIND["sextupole"], 0, 0, 1, element.L, 0.0, 0.0, 0.0, 0.0, element.K2)
    return _lib.Element(KIND["drift"], 0, 0, 1, element.L, 0.0, 0.0, 0.0, 0.0, 0.0)


def track_exact(coords, element, p0c, mc2=M_ELECTRON):
    """coords: sequence (x, px, y, py, z, pz) of equally shaped arrays; returns the same.  CUDA tensors are transported
    in place by one kernel launch; host arrays by the numpy expressions above."""
    if getattr(coords[0], "is_cuda", False):
        from . import ops
        coords = tuple(c if c.is_contiguous() else c.contiguous() for c in coords)
        ops.track_element(coords, _device_element(element), p0c, mc2)
        return coords
    return _apply_exact(coords, element, float(p0c), float(mc2))


# ------------------------------------------------------------------------------- first-order option
def _edge(x, px, y, py, g, e):
    if e == 0.0 or g == 0.0:
        return px, py
    k = g * math.tan(e)
    return px + k * x, py - k * y


def linear_matrix(element):
    """The 6 x 6 first-order transfer matrix of `element`: the maps below applied to the unit vectors."""
    cols = _apply_map(tuple(np.eye(6)), element)          # row k of eye = coordinate k of the six unit particles
    return np.stack([np.asarray(c, dtype=np.float64) for c in cols])


def track_linear(coords, element):
    """First-order transport.  CUDA tensors are transformed in place by one kernel launch (ops.track_linear); host
    arrays by the expressions in `_apply_map`."""
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


# ------------------------------------------------------------------------------- driver interface
def make_element(kind: str, L: float, p0c: float = 0.0, **kw):
    """Element factory used by CSR2D.get_bmadx_element (CSR.py:146-199): Bmad-X elements when the package is
    importable, the dataclasses above otherwise.  kind in {'drift', 'dipole', 'quad', 'sextupole'}."""
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


def track(coords, element, s=0.0, p0c=0.0, mc2=M_ELECTRON, order="exact"):
    """beams.py:101-102: track_element(particle, element).  coords = (x, px, y, py, z, pz)."""
    if HAVE_BMADX and not isinstance(element, (Drift, SBend, Quadrupole, Sextupole)):
        part = _bmadx.track_element(_bmadx.Particle(*coords, s, p0c, mc2), element)
        return part.x, part.px, part.y, part.py, part.z, part.pz
    if order == "first":
        return track_linear(coords, element)
    return track_exact(coords, element, p0c, mc2)
