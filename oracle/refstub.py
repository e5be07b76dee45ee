"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference (slaclab/pyDFCSR) in the
build container so that (a) the restatement in ``oracle/dfcsr_oracle.py`` can be pinned
against it and (b) golden vectors under ``tests/golden/`` can be generated.

``/root/reference`` does not exist on the GPU box: nothing that runs there may call
``load_reference()``; tests that do are skipped when the directory is missing.

The reference imports six third-party packages that are absent in this image and are
never touched by the hot path (tracking, beam generation, I/O, plotting).  They are
replaced by inert stub modules *before* ``pyDFCSR_2D`` is imported.
"""
from __future__ import annotations

import collections
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DFCSR_REFERENCE_ROOT", "/root/reference")

_Particle = collections.namedtuple("Particle", "x px y py z pz s p0c mc2")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyDFCSR_2D"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stubs() -> None:
    class _Elem:  # inert element classes: only constructed by the reference's run() loop
        def __init__(self, **kw):
            self.__dict__.update(kw)

    def _no_tracking(*_a, **_k):
        raise RuntimeError("bmadx is stubbed: tracking is not part of the oracle path")

    if "bmadx" not in sys.modules:
        b = _stub("bmadx", Particle=_Particle, M_ELECTRON=0.51099895e6, track_element=_no_tracking,
                  Drift=type("Drift", (_Elem,), {}), SBend=type("SBend", (_Elem,), {}),
                  Quadrupole=type("Quadrupole", (_Elem,), {}), Sextupole=type("Sextupole", (_Elem,), {}))
        b.constants = _stub("bmadx.constants", C_LIGHT=299792458.0, M_ELECTRON=0.51099895e6,
                            E_CHARGE=1.602176634e-19)
        b.structures = _stub("bmadx.structures", Particle=_Particle)
    if "mpi4py" not in sys.modules:
        m = _stub("mpi4py")
        m.MPI = _stub("mpi4py.MPI", COMM_WORLD=None, DOUBLE=None)
    if "h5py" not in sys.modules:
        _stub("h5py")
    if "distgen" not in sys.modules:
        _stub("distgen", Generator=object)
    if "pmd_beamphysics" not in sys.modules:
        _stub("pmd_beamphysics", ParticleGroup=object)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.cm = _stub("matplotlib.cm")


def load_reference():
    """Return the imported reference package ``pyDFCSR_2D`` (always under that module name,
    so numba's on-disk cache keys stay consistent)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/dfcsr_numba_cache")  # reference tree is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pyDFCSR_2D  # noqa: F401
    return sys.modules["pyDFCSR_2D"]


class FakeBeam:
    """Duck-typed stand-in for ``pyDFCSR_2D.beams.Beam`` exposing exactly what the hot path reads
    (beams.py:88-98, 137-215).  Statistics are computed with the same numpy calls as the reference."""

    def __init__(self, x, px, z, pz, position, charge=1.0e-9, init_energy=5.0e9):
        import numpy as np
        self._np = np
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.px = np.ascontiguousarray(px, dtype=np.float64)
        self.z = np.ascontiguousarray(z, dtype=np.float64)
        self.pz = np.ascontiguousarray(pz, dtype=np.float64)
        self.position = position
        self.charge = charge
        self.init_energy = init_energy
        self.step = 0
        self.update_status()

    def update_status(self):
        np = self._np
        self._sigma_x = np.std(self.x)
        self._sigma_z = np.std(self.z)
        self._slope = np.polyfit(self.z, self.x, deg=1)
        self._mean_x = np.mean(self.x)
        self._mean_z = np.mean(self.z)

    sigma_x = property(lambda self: self._np.std(self.x))
    sigma_z = property(lambda self: self._np.std(self.z))
    mean_x = property(lambda self: self._np.mean(self.x))
    mean_z = property(lambda self: self._np.mean(self.z))
    slope = property(lambda self: self._np.polyfit(self.z, self.x, deg=1))

    @property
    def x_transform(self):
        return self.x - self._np.polyval(self.slope, self.z)


def make_reference_csr(lattice_yaml, deposition_cfg, integration_cfg, csr_cfg):
    """Build a reference ``CSR2D`` without running its constructor (which needs distgen/bmadx),
    wired with the real ``Lattice``, ``DF_tracker`` and parameter objects."""
    ref = load_reference()
    from pyDFCSR_2D.CSR import CSR2D
    from pyDFCSR_2D.deposit import DF_tracker
    from pyDFCSR_2D.lattice import Lattice
    from pyDFCSR_2D.params import CSR_params, Integration_params
    csr = object.__new__(CSR2D)
    csr.lattice = Lattice({"lattice_input_file": lattice_yaml})
    csr.DF_tracker = DF_tracker(dict(deposition_cfg))
    csr.integration_params = Integration_params(dict(integration_cfg))
    csr.CSR_params = CSR_params(dict(csr_cfg))
    csr.parallel = False
    csr.formation_length = 0.0
    return csr
