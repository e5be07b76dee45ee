"""Device-resident ``Beam`` with the reference's public surface (beams.py:12-229): construction from
the ``input_beam`` YAML block, ``track``, ``apply_wakes``, ``update_status`` and the statistics
properties read by the hot path (``_sigma_x, _sigma_z, _slope, _mean_x, _mean_z, x_transform``).

Particles are six CUDA float64 tensors (Bmad-X order x, px, y, py, z, pz); they are uploaded once
and stay in HBM across tracking, deposition and kick application.  All O(Np) statistics come from
the device reduction kernels (``ops.beam_stats`` per state change, ``ops.beam_cov`` for Twiss), not from numpy.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops, synth, tracking
from ._lib import Axis

MC2 = 0.51099895e6      # electron rest energy [eV] (physical_constants.py)


class Beam:
    def __init__(self, input_beam, device=None, shards=None):
        """shards: None (all particles on this GPU, like every MPI rank of the reference), or a callable
        n_total -> distributed.ParticleShards: this rank then keeps only its chunks of the bunch, and the statistics,
        the covariance and (through DF_tracker) the deposit combine the shards exactly (bit-identical to one GPU)."""
        self.check_inputs(input_beam)
        self.input_beam_config = input_beam
        self.style = input_beam["style"]
        self.device = torch.device(device if device is not None else "cuda")
        if self.style == "from_file":                                  # beams.py:22-35
            coords = np.loadtxt(input_beam["beamfile"])
            assert coords.shape[1] == 6, f"Error: input beam must have 6 dimension, but get {coords.shape[1]} instead"
            coords = coords.T
            self._charge = input_beam["charge"]
            self._init_energy = input_beam["energy"]
        elif self.style == "array":                                    # in-memory (6, Np) array or tensor
            coords = input_beam["coords"]
            self._charge = input_beam["charge"]
            self._init_energy = input_beam["energy"]
        elif self.style == "synthetic":                                # seeded Gaussian (distgen stand-in)
            kw = {k: v for k, v in input_beam.items() if k not in ("style", "verbose", "tracking_order")}
            n = kw.pop("n_particle")
            coords = synth.gaussian_bunch(n, **kw)
            self._charge = input_beam.get("charge", synth.CHICANE_BEAM["charge"])
            self._init_energy = input_beam.get("energy", synth.CHICANE_BEAM["energy"])
        elif self.style == "distgen":                                  # beams.py:38-46
            try:
                from distgen import Generator
            except ImportError as e:
                raise ImportError("input_beam style 'distgen' needs the distgen package; use style "
                                  "'from_file', 'array' or 'synthetic' instead") from e
            gen = Generator(input_beam["distgen_input_file"])
            gen.run()
            pg = gen.particles
            self._charge = pg["charge"]
            self._init_energy = float(np.mean(pg["energy"]))
            p0c = self._init_energy
            coords = np.stack([pg.x, pg.px / p0c, pg.y, pg.py / p0c,
                               -299792458.0 * pg.beta * (pg.t - np.mean(pg.t)), (pg.p - p0c) / p0c])
        else:
            raise ImportError("input_beam style 'ParticleGroup' needs pmd_beamphysics/h5py (not available offline)")
        self.n_total = int(coords.shape[1])
        # centre of the first statistics pass: the first particle of the WHOLE bunch (known on every rank before sharding)
        self._centre = [float(coords[0][0]), float(coords[4][0]), float(coords[5][0])]
        self._centre6 = [float(coords[k][0]) for k in range(6)]
        self.shards = shards(self.n_total) if callable(shards) else shards
        if self.shards is not None:
            coords = self.shards.local_slice(coords)
        if isinstance(coords, torch.Tensor):
            c = coords.to(self.device, torch.float64)
        else:
            c = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).to(self.device)
        self.coords = [c[k].contiguous() for k in range(6)]
        self._init_gamma = self._init_energy / MC2
        self.tracking_order = input_beam.get("tracking_order", "exact") if self.style == "synthetic" else "exact"
        self.position = 0
        self.step = 0
        self.update_status()

    def check_inputs(self, input_beam):
        assert "style" in input_beam, "ERROR: input_beam must have keyword <style>"
        required = {"from_file": ["style", "beamfile", "charge", "energy"],
                    "distgen": ["style", "distgen_input_file"],
                    "ParticleGroup": ["style", "ParticleGroup_h5"],
                    "array": ["style", "coords", "charge", "energy"],
                    "synthetic": ["style", "n_particle"]}
        if input_beam["style"] not in required:
            raise Exception("input beam parsing Error: invalid input style")
        self.required_inputs = required[input_beam["style"]]
        for req in self.required_inputs:
            assert req in input_beam, f"Required input parameter {req} to Beam.__init__(**kwargs) was not found."
        if input_beam["style"] != "synthetic":
            allowed = self.required_inputs + ["verbose"]
            for key in input_beam:
                assert key in allowed, f"Incorrect param given to Beam.__init__(**kwargs): {key}\nAllowed params: {allowed}"

    # ------------------------------------------------------------------------------- state
    def update_status(self):
        """One pass of the device reductions replaces np.std/np.mean/np.polyfit (beams.py:88-98).  The pass is only
        ENQUEUED here; the host waits for it when a statistic is first read (`stats`, `_sigma_x`, `_slope`, ...), so the
        caller can keep launching work -- the reduction that follows a kick overlaps with whatever is enqueued next."""
        self._pending_stats = ops.beam_stats_async(self.x, self.z, self.pz, self.px, centre=self._centre, shards=self.shards)

    @property
    def stats(self):
        """The 16 doubles of `dfcsr_beam_stats` for the current particle state (blocks until they have arrived).
        Reading them also moves the centre of the next first pass to the current means -- a deterministic rule (it follows
        the program order, not the timing), identical on every rank and for every way of sharding the particles."""
        p = self._pending_stats
        if p._value is None:              # first read of this pass
            st = p.get()
            self._centre = [float(st[_lib.S_MEAN_X]), float(st[_lib.S_MEAN_Z]), float(st[_lib.S_MEAN_PZ])]
            return st
        return p._value

    _sigma_x = property(lambda self: float(self.stats[_lib.S_SIGMA_X]))
    _sigma_z = property(lambda self: float(self.stats[_lib.S_SIGMA_Z]))
    _slope = property(lambda self: np.array([self.stats[_lib.S_SLOPE], self.stats[_lib.S_INTERCEPT]]))
    _mean_x = property(lambda self: float(self.stats[_lib.S_MEAN_X]))
    _mean_z = property(lambda self: float(self.stats[_lib.S_MEAN_Z]))

    def track(self, element, step_size, update_step=True):
        """beams.py:101-106.  `element` comes from tracking.make_element: a Bmad-X element when that package is
        importable (tracked with bmadx.track_element on the device tensors), else an element of pydfcsr_b200.tracking
        transported by the restated Bmad maps on the device (`tracking_order = "first"`: first-order matrices)."""
        self.coords = [c.contiguous() for c in tracking.track(tuple(self.coords), element, self.position,
                                                              self._init_energy, MC2, order=self.tracking_order)]
        self.position += step_size
        if update_step:
            self.step += 1
        self.update_status()

    def apply_wakes(self, dE_dct, x_kick, xrange, zrange, step_size, transverse_on):
        """beams.py:108-131 on the device, in place."""
        dE = dE_dct if isinstance(dE_dct, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(dE_dct)).to(self.device)
        kick = x_kick if isinstance(x_kick, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_kick)).to(self.device)
        xa = xrange if isinstance(xrange, Axis) else Axis.make(xrange[0], xrange[-1], len(xrange))
        za = zrange if isinstance(zrange, Axis) else Axis.make(zrange[0], zrange[-1], len(zrange))
        ops.apply_kick(self.x, self.z, self.coords[1], self.coords[5], self._slope[0], self._slope[1],
                       dE.contiguous(), kick.contiguous(), xa, za, step_size, self.init_energy, transverse_on)
        self.update_status()

    # ------------------------------------------------------------------------------- properties
    x = property(lambda self: self.coords[0])
    px = property(lambda self: self.coords[1])
    y = property(lambda self: self.coords[2])
    py = property(lambda self: self.coords[3])
    z = property(lambda self: self.coords[4])
    pz = property(lambda self: self.coords[5])
    mean_x = property(lambda self: self._mean_x)
    mean_z = property(lambda self: self._mean_z)
    sigma_x = property(lambda self: self._sigma_x)
    sigma_z = property(lambda self: self._sigma_z)
    slope = property(lambda self: self._slope)
    init_energy = property(lambda self: self._init_energy)
    init_gamma = property(lambda self: self._init_gamma)
    charge = property(lambda self: self._charge)
    mean_y = property(lambda self: float(ops.beam_cov(self.coords, centre=self._centre6, shards=self.shards)[0][2]))
    mean_energy = property(lambda self: (float(self.stats[_lib.S_MEAN_PZ]) + 1) * self._init_energy)
    sigma_energy = property(lambda self: float(self.stats[_lib.S_SIGMA_PZ]) * self._init_energy)
    sigma_x_transform = property(lambda self: float(self.stats[_lib.S_SIGMA_XT]))
    mean_x_transform = property(lambda self: float(self.stats[_lib.S_MEAN_XT]))

    @property
    def energy(self):
        return (self.pz + 1) * self._init_energy

    @property
    def gamma(self):
        return self.energy / MC2

    @property
    def x_transform(self):
        """x with the x-z chirp removed (beams.py:206-211); a device tensor."""
        return self.x - (self._slope[0] * self.z + self._slope[1])

    def twiss_async(self):
        """Enqueue the covariance pass behind the Twiss statistics; returns a callable that waits for it (once) and
        returns the dictionary of `twiss`.  Lets the driver collect per-step statistics without a host sync per step."""
        pending = ops.beam_cov_async(self.coords, centre=self._centre6, shards=self.shards)
        energy = self._init_energy

        def resolve():
            _, cov6 = pending.get()
            out = {}
            for plane, idx in (("x", (0, 1, 5)), ("y", (2, 3, 5))):
                cov = cov6[np.ix_(idx, idx)]
                d2, xd, pd = cov[2, 2], cov[0, 2], cov[1, 2]
                eb, eg, ea = cov[0, 0] - xd ** 2 / d2, cov[1, 1] - pd ** 2 / d2, -cov[0, 1] + xd * pd / d2
                emit = np.sqrt(eb * eg - ea ** 2)
                vals = dict(alpha=ea / emit, beta=eb / emit, gamma=eg / emit, emit=emit, eta=xd / d2, etap=pd / d2,
                            norm_emit=emit * energy / MC2)
                out.update({f"{k}_{plane}": v for k, v in vals.items()})
            return out
        return resolve

    @property
    def twiss(self):
        """Twiss/dispersion from the 3x3 covariances of (x, px, pz) and (y, py, pz) (twiss.py:2-71); the 6x6
        covariance comes from one device reduction (ops.beam_cov)."""
        return self.twiss_async()()

    def to_host(self) -> np.ndarray:
        """(6, n_total) host array of ALL particles (collective when the particles are sharded)."""
        if self.shards is not None:
            return torch.stack([self.shards.gather(c) for c in self.coords]).cpu().numpy()
        return torch.stack(self.coords).cpu().numpy()
