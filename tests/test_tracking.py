"""Particle transport (beams.py:101-106 -> Bmad-X track_element) restated in pydfcsr_b200/tracking.py, pinned on the
single-particle known answers the reference holds: test/test_BmadX_tracking.ipynb cells 25 (Drift L = 1), 28 (SBend
L = 0.1, G = 0.5, E1 = 0, E2 = 0.1) and 31 (Quadrupole L = 0.1, K1 = 10) on the particle 1e-3 * ones(6) at p0c = 4e7 eV,
mc2 = 510998.95 eV.  The numbers below are the notebook's stored outputs, copied as data."""
import numpy as np
import pytest

from pydfcsr_b200 import tracking

P0C = 4.0e7
MC2 = 510998.94999999995
START = tuple(np.full(1, 1e-3) for _ in range(6))
KNOWN = {
    "drift": (tracking.Drift(L=1.0),
              [0.0019990019960084817, 0.001, 0.0019990019960084817, 0.001, 0.000999164924440183, 0.001]),
    "sbend": (tracking.SBend(L=0.1, G=0.5, E1=0.0, E2=0.1, P0C=P0C),
              [0.001101157687964056, 0.0010789320635333618, 0.0010999526674531663, 0.000944818304900785,
               0.0009473961722735597, 0.0009999999999998899]),
    "quadrupole": (tracking.Quadrupole(L=0.1, K1=10.0),
                   [0.0010487094525905673, -3.296855035329286e-05, 0.00115193903813484, 0.0020671006533949897,
                    0.000999879829429573, 0.001]),
}
# first-order results of the reference's own r_gen6 for the same elements (cells 26, 29, 32): what the exact maps must NOT be
R_GEN6 = {"drift": [0.002, 0.001, 0.002, 0.001, 0.001, 0.001],
          "sbend": [0.00110121, 0.00107898, 0.0011, 0.00094482, 0.00094748, 0.001],
          "quadrupole": [1.04875693e-03, -3.30011883e-05, 1.15209308e-03, 2.06716826e-03, 1.0e-03, 1.0e-03]}


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_bmadx_known_answers(name):
    element, want = KNOWN[name]
    got = [float(np.asarray(c).ravel()[0]) for c in tracking.track_exact(START, element, P0C, MC2)]
    for g, w in zip(got, want):
        assert abs(g - w) <= 1e-12 * abs(w), (name, got, want)
    # ... and they do differ from first order where Bmad-X does (drift x: 0.0019990 vs 0.002)
    first = [float(np.asarray(c).ravel()[0]) for c in tracking.track_linear(START, element)]
    assert np.allclose(first, R_GEN6[name], rtol=2e-6)
    assert max(abs(g - f) / abs(f) for g, f in zip(got, first)) > 1e-5


def test_exact_maps_are_symplectic_and_compose():
    """Jacobian of every map by central differences is symplectic (the maps are exact Hamiltonian flows or products of
    them); two half-length bends without inner edges compose to the full bend; g -> 0 is the drift."""
    J = np.zeros((6, 6))
    for k in range(3):
        J[2 * k, 2 * k + 1], J[2 * k + 1, 2 * k] = 1.0, -1.0
    v0 = np.array([2e-4, -3e-5, 1e-4, 2e-5, 3e-4, 1.5e-3])
    elements = [tracking.Drift(0.7), tracking.SBend(L=0.5002, G=0.0483 / 0.5002, E1=0.02, E2=0.0483),
                tracking.SBend(L=0.3, G=-0.2, FRINGE_AT="no_end"), tracking.Quadrupole(L=0.2, K1=1.7),
                tracking.Quadrupole(L=0.2, K1=-0.9, NUM_STEPS=3), tracking.Sextupole(L=0.1, K2=3.0)]

    def f(v, el):
        # canonical pairs of Bmad: (x, px), (y, py), (z, pz) with z the coordinate and pz its momentum
        out = tracking.track_exact(tuple(np.array([c]) for c in v), el, P0C, MC2)
        return np.array([float(c[0]) for c in out])

    for el in elements:
        M = np.zeros((6, 6))
        for k in range(6):
            h = 1e-6 * (1.0 if k % 2 == 0 else 1.0)
            e = np.zeros(6)
            e[k] = h
            M[:, k] = (f(v0 + e, el) - f(v0 - e, el)) / (2 * h)
        assert np.max(np.abs(M.T @ J @ M - J)) < 2e-8, (el, np.max(np.abs(M.T @ J @ M - J)))
    a = f(f(v0, tracking.SBend(L=0.25, G=0.1, E1=0.03, FRINGE_AT="entrance_end")), tracking.SBend(L=0.25, G=0.1, E2=0.05, FRINGE_AT="exit_end"))
    b = f(v0, tracking.SBend(L=0.5, G=0.1, E1=0.03, E2=0.05))
    assert np.max(np.abs(a - b)) < 1e-15
    assert np.max(np.abs(f(v0, tracking.SBend(L=0.4, G=1e-14, FRINGE_AT="no_end")) - f(v0, tracking.Drift(0.4)))) < 1e-15
