"""``CSR2D`` driver: the reference's orchestration (CSR.py:27-879) kept in Python, with the hot path
behind it running on the B200.

Same construction (`CSR2D(input_file, parallel)`), same YAML schema and defaults, same public
methods and result attributes:

    run(stop_time, debug)          CSR.py:202   element loop / step splitting / formation length
    get_CSR_mesh()                 CSR.py:361   -> CSR_xmesh, CSR_zmesh, CSR_zrange, CSR_xrange_transformed
    calculate_2D_CSR()             CSR.py:397   -> dE_dct, x_kick  (one K4 launch)
    calculate_2D_CSR_parallel()    CSR.py:420   mesh sharded with the reference's MPI block split,
                                                NCCL all-gather in place of the two Allgatherv calls
    get_CSR_wake(s, x, debug)      CSR.py:454   single point / integrand arrays

`parallel=True` expects one process per GPU under torchrun (RANK/LOCAL_RANK/WORLD_SIZE), replacing
`mpirun -n P python -m pyDFCSR_mpi_run` (pyDFCSR_mpi_run.py).  Every rank replicates tracking,
deposition and history exactly as every MPI rank does in the reference (CSR.py:202-307).
"""
from __future__ import annotations

import datetime
import os
import time

import numpy as np
import torch

from . import distributed as dist_utils
from . import _lib, ops, outputs, tracking
from ._lib import Axis
from .beams import Beam
from .deposit import DF_tracker
from .lattice import Lattice
from .params import CSR_params, Integration_params
from .yaml_parser import full_path, parse_yaml


def isotime():
    return datetime.datetime.now(datetime.timezone.utc).astimezone().replace(microsecond=0).isoformat().replace(":", "_")


class CSR2D:
    def __init__(self, input_file=None, parallel=False, device=None, verbose=True, precision="fp64", shard_particles=None):
        """precision='fp32' selects the optional mixed-precision history (wakes within 1e-4; default is
        the fp64 parity mode).  shard_particles (parallel runs): True = every rank tracks, deposits and kicks 1/N of
        the particles and the shards are combined exactly (statistics tables, integer deposit grids) -- same bits as
        replicating them; False = every rank keeps all particles like every MPI rank of the reference; None = True
        unless DFCSR_SHARD_PARTICLES=0.  Everything else is the reference's signature."""
        self.precision = precision
        if shard_particles is None:
            shard_particles = os.environ.get("DFCSR_SHARD_PARTICLES", "1") != "0"
        self.shard_particles = bool(shard_particles) and bool(parallel)
        # how the observation mesh is dealt out to the ranks when the exchange is fused into the wake kernel: "interleaved"
        # (point k to rank k mod P: equal work everywhere) or "block" (the reference's contiguous count/displ blocks,
        # CSR.py:121-125, whose cost differs from rank to rank with the number of in-grid samples).  Same grid bits either
        # way; the NCCL all-gather path always uses the reference's blocks.
        self.mesh_split = os.environ.get("DFCSR_MESH_SPLIT", "interleaved")
        # which work split the wake kernel uses: "auto" = one lane per observation point (x-groups, ops.wake_grid_xgroups)
        # whenever its plan applies to the step -- no chirp band, a bunch that fills its history grid, a mesh row of
        # >= 23 points -- else one CTA per point (ops.wake_grid); "point" forces the latter.  The plan depends on the step's
        # scalars and the whole mesh only, so every rank of a parallel run takes the same decision.
        self.wake_mapping = os.environ.get("DFCSR_WAKE_MAPPING", "auto")
        self.timestamp = isotime()
        self.verbose = verbose
        self.parallel = bool(parallel)
        if self.parallel:
            self.init_MPI()
        else:
            self.rank, self.world_size = 0, 1
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)) if self.parallel else torch.cuda.current_device())
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)                       # make sure the device has a context
        _lib.check(_lib.lib.dfcsr_wake_preload(), "dfcsr_wake_preload")   # kernel code is loaded now, not in the first step
        if input_file is not None:
            self.parse_input(input_file)
            self.input_file = input_file
        self.formation_length = None
        self.initialization()
        self.prefix = f"{self.CSR_params.write_name}-{self.timestamp}"
        if self.parallel:
            self._split_mesh()

    # ------------------------------------------------------------------------------- input
    def parse_input(self, input_file):
        inp = parse_yaml(input_file)
        self.check_input_consistency(inp)
        self.input = inp
        shards = None
        if self.shard_particles and self.world_size > 1:
            dep = inp.get("particle_deposition") or {}
            cells = max(100 * 100, int(dep.get("xbins", 100)) * int(dep.get("zbins", 100)))     # deposit.py:160-167
            shards = lambda n: dist_utils.ParticleShards(n, self.device, cells)                  # noqa: E731
        self.beam = Beam(inp["input_beam"], device=self.device, shards=shards)
        self.lattice = Lattice(inp["input_lattice"])
        self.DF_tracker = DF_tracker(inp.get("particle_deposition"), device=self.device, precision=self.precision,
                                     shards=self.beam.shards)
        self.integration_params = Integration_params(inp.get("CSR_integration"))
        self.CSR_params = CSR_params(inp.get("CSR_computation"))

    def check_input_consistency(self, inp):
        self.required_inputs = ["input_beam", "input_lattice"]
        allowed = self.required_inputs + ["particle_deposition", "distribution_interpolation", "CSR_integration",
                                          "CSR_computation"]
        for key in inp:
            assert key in allowed, f"Incorrect param given to CSR2D.__init__(**kwargs): {key}\nAllowed params: {allowed}"
        for req in self.required_inputs:
            assert req in inp, f"Required input parameter {req} to CSR2D.__init__(**kwargs) was not found."

    def initialization(self):
        """Deposit the initial beam (CSR.py:70-81)."""
        b = self.beam
        self.DF_tracker.get_DF(x=b.x, z=b.z, px=b.px, t=b.position, stats=b.stats)
        self.DF_tracker.append_DF()
        self.DF_tracker.append_interpolant(formation_length=float("inf"),
                                           n_formation_length=self.integration_params.n_formation_length)
        self.CSR_scaling = 8.98755e3 * b.charge
        self.init_statistics()

    def init_statistics(self):
        n = self.lattice.total_steps
        keys = ["alpha", "beta", "gamma", "emit", "eta", "etap", "norm_emit"]
        self.statistics = {"twiss": {f"{k}_{p}": np.zeros(n) for p in ("x", "y") for k in keys},
                           "slope": np.zeros((n, 2))}
        for k in ("sigma_x", "sigma_z", "sigma_energy", "mean_x", "mean_z", "mean_energy"):
            self.statistics[k] = np.zeros(n)
        self._pending_twiss = []
        self.update_statistics(step=0)
        self.inbend = False
        self.afterbend = False
        self.R_rec = None
        self.phi_rec = None

    def _flush_statistics(self):
        e0 = self.beam.init_energy
        for step, resolve, pending in self._pending_twiss:
            for k, v in resolve().items():
                self.statistics["twiss"][k][step] = v
            st = pending.get()                       # the statistics pass that was current when the step was recorded
            self.statistics["slope"][step, :] = (st[_lib.S_SLOPE], st[_lib.S_INTERCEPT])
            self.statistics["sigma_x"][step] = st[_lib.S_SIGMA_X]
            self.statistics["sigma_z"][step] = st[_lib.S_SIGMA_Z]
            self.statistics["sigma_energy"][step] = st[_lib.S_SIGMA_PZ] * e0
            self.statistics["mean_x"][step] = st[_lib.S_MEAN_X]
            self.statistics["mean_z"][step] = st[_lib.S_MEAN_Z]
            self.statistics["mean_energy"][step] = (st[_lib.S_MEAN_PZ] + 1) * e0
        self._pending_twiss = []

    def update_statistics(self, step):
        if step >= self.lattice.total_steps:
            return
        # the covariance pass is only enqueued; its 27 numbers are collected by _flush_statistics (end of run(), or
        # when the statistics are written), so that a lattice step does not pay a host synchronisation for them
        self._pending_twiss.append((step, self.beam.twiss_async(), self.beam._pending_stats))

    # ------------------------------------------------------------------------------- multi-GPU
    def init_MPI(self):
        """Name kept from the reference (CSR.py:116-125); the communicator is torch.distributed/NCCL."""
        self.rank, self.world_size = dist_utils.init_process_group()

    def _split_mesh(self):
        n = self.CSR_params.xbins * self.CSR_params.zbins
        self.count, displ = dist_utils.split_counts(n, self.world_size)
        self.displ = np.array(displ)
        # fused K4 + exchange over NVLink peer memory when the ranks can map each other's grids, else NCCL
        self._peer_grid = dist_utils.make_peer_wake_grid(n, self.device)

    # ------------------------------------------------------------------------------- run loop
    def get_formation_length(self, R, sigma_z, phi=0.0, inbend=True):
        if inbend:
            self.formation_length = (24 * (R ** 2) * sigma_z) ** (1 / 3)
        else:
            self.formation_length = (3 * R ** 2 * phi ** 4) / (4 * (-6 * sigma_z + R * phi ** 3))

    def get_bmadx_element(self, ele, DL, entrance=False, exit=False):
        """Element for one (partial) step (CSR.py:146-199); returns a `tracking` stand-in element."""
        cfg = dict(self.lattice.lattice_config[ele])
        kind = cfg["type"]
        p0c = self.beam.init_energy
        if kind == "dipole":
            L = cfg["L"]
            G = cfg["G"] if "G" in cfg else cfg["angle"] / L
            E1, E2 = cfg.get("E1", 0), cfg.get("E2", 0)
            if entrance and exit:
                return tracking.make_element("dipole", DL, p0c, G=G, E1=E1, E2=E2, FRINGE_AT=cfg.get("FRINGE_AT", "both_ends"))
            if entrance:
                return tracking.make_element("dipole", DL, p0c, G=G, E1=E1, E2=0.0, FRINGE_AT="entrance_end")
            if exit:
                return tracking.make_element("dipole", DL, p0c, G=G, E1=0.0, E2=E2, FRINGE_AT="exit_end")
            return tracking.make_element("dipole", DL, p0c, G=G, E1=0.0, E2=0.0, FRINGE_AT="no_end")
        if kind == "quad":
            return tracking.make_element("quad", DL, p0c, K1=cfg["K1"])
        if kind == "sextupole":
            return tracking.make_element("sextupole", DL, p0c, K2=cfg["K2"])
        return tracking.make_element("drift", DL, p0c)

    def _log(self, *a):
        if self.verbose and self.rank == 0:
            print(*a)

    def hot_path_step(self, apply=True, kick_length=None):
        """One pass of the ★ path at the current beam state (CSR.py:299-307, 324-334): deposit,
        history push, mesh, wake, kick.  Returns nothing; results in dE_dct / x_kick."""
        kick_length = self.lattice.step_size if kick_length is None else kick_length
        b = self.beam
        self.DF_tracker.get_DF(x=b.x, z=b.z, px=b.px, t=b.position, stats=b.stats)
        self.DF_tracker.append_DF()
        self.DF_tracker.append_interpolant(formation_length=self.formation_length,
                                           n_formation_length=self.integration_params.n_formation_length)
        self.DF_tracker.build_interpolant()
        self.get_CSR_mesh()
        if self.parallel:
            self.calculate_2D_CSR_parallel()
        else:
            self.calculate_2D_CSR()
        if apply:
            b.apply_wakes(self.dE_dct, self.x_kick, self._mesh_axes[0], self._mesh_axes[1],
                          kick_length, self.CSR_params.transverse_on)

    def run(self, stop_time=None, debug=False):
        self._log("Starting the DFCSR run")
        lat, b = self.lattice, self.beam
        step_count = 1
        DL = lat.step_size
        ele_count = 0
        skip_ele = False
        self.inbend = self.afterbend = False
        self.formation_length = 0.0
        ele_prev = None
        for ele in list(lat.lattice_config.keys())[1:]:
            lat.update(ele)
            cfg = lat.lattice_config[ele]
            L, kind = cfg["L"], cfg["type"]
            steps = lat.steps_per_element[ele_count]
            if (not skip_ele) and ele_count > 0:                           # CSR.py:228-236
                DL_1 = lat.distance[ele_count - 1] - b.position
                if DL_1 > 1.0e-6:
                    b.track(self.get_bmadx_element(ele=ele_prev, DL=DL_1, exit=True), DL_1, update_step=False)
            if steps == 0:                                                 # CSR.py:238-241
                skip_ele = True
                b.track(self.get_bmadx_element(ele=ele, DL=L, exit=True, entrance=True), L, update_step=False)
            if kind == "dipole":                                           # CSR.py:246-256
                R = L / cfg["angle"]
                self.inbend = self.afterbend = True
                self.R_rec, self.phi_rec = R, cfg["angle"]
                self.get_formation_length(R=R, sigma_z=5 * b.sigma_z, inbend=True)
            else:
                self.inbend = False
                if self.afterbend:
                    self.get_formation_length(R=self.R_rec, sigma_z=5 * b.sigma_z, inbend=True)
                else:
                    self.formation_length += L
            for step in range(steps):
                t0 = time.time()
                if step == 0 and ele_count > 0:                            # CSR.py:279-288
                    DL_2 = lat._positions_record[step_count] - lat.distance[ele_count - 1]
                    b.track(self.get_bmadx_element(ele=ele, DL=DL_2, entrance=True), DL_2)
                    skip_ele = False
                else:
                    b.track(self.get_bmadx_element(ele=ele, DL=DL), DL)
                if debug or self.CSR_params.compute_CSR:                   # CSR.py:297-307
                    self.DF_tracker.prefetch_DF(b)         # deposit + density functions enqueued behind the statistics pass
                    self.DF_tracker.get_DF(x=b.x, z=b.z, px=b.px, t=b.position, stats=b.stats)
                    self.DF_tracker.append_DF()
                    self.DF_tracker.append_interpolant(formation_length=self.formation_length,
                                                       n_formation_length=self.integration_params.n_formation_length)
                    self.DF_tracker.build_interpolant()
                if self.CSR_params.compute_CSR and step % lat.nsep[ele_count] == 0:   # CSR.py:321-339
                    self.get_CSR_mesh()
                    if self.parallel:
                        self.calculate_2D_CSR_parallel()
                    else:
                        self.calculate_2D_CSR()
                    if self.CSR_params.apply_CSR:
                        b.apply_wakes(self.dE_dct, self.x_kick, self._mesh_axes[0], self._mesh_axes[1],
                                      DL * lat.nsep[ele_count], self.CSR_params.transverse_on)
                    wb = self.CSR_params.write_beam
                    if wb == "all" or (isinstance(wb, list) and step_count in wb):
                        self.dump_beam(label=step_count)
                    if self.CSR_params.write_wakes:
                        self.write_wakes()
                self.update_statistics(step=step_count)
                self._log("Finish step {}, s = {},  in {} seconds".format(step_count, b.position, time.time() - t0))
                step_count += 1
                if stop_time and b.position > stop_time:
                    self._flush_statistics()
                    return
            ele_prev = ele
            ele_count += 1
        self._flush_statistics()
        self.dump_beam(label="end")
        self.write_statistics()

    # ------------------------------------------------------------------------------- mesh + wake
    def get_CSR_mesh(self):
        """CSR.py:361-394.  Only the two axes (xbins + zbins numbers) are built on the host; the N mesh
        points themselves are generated inside the wake kernel from the axes and the chirp line
        (`dfcsr_wake_grid`), so nothing O(N) is built or uploaded per step.  `CSR_xmesh` / `CSR_zmesh`
        remain available as lazily evaluated host arrays."""
        b, p = self.beam, self.CSR_params
        st = b.stats
        sig_x, mean_x = float(st[_lib.S_SIGMA_XT]), float(st[_lib.S_MEAN_XT])
        mean_z, sig_z = float(st[_lib.S_MEAN_Z]), float(st[_lib.S_SIGMA_Z])
        # np.linspace(a, b, n)[0] is a and [-1] is b exactly: the axes need only the end points (the arrays are lazy)
        self._mesh_ends = (mean_x - p.xlim * sig_x, mean_x + p.xlim * sig_x, mean_z - p.zlim * sig_z, mean_z + p.zlim * sig_z)
        self._mesh_slope = (float(st[_lib.S_SLOPE]), float(st[_lib.S_INTERCEPT]))
        self._mesh_axes = (Axis.make(self._mesh_ends[0], self._mesh_ends[1], p.xbins),
                           Axis.make(self._mesh_ends[2], self._mesh_ends[3], p.zbins))
        self._mesh_host = None
        self._ranges_host = None

    def _ranges(self):
        if self._ranges_host is None:
            e, p = self._mesh_ends, self.CSR_params
            self._ranges_host = (np.linspace(e[0], e[1], p.xbins), np.linspace(e[2], e[3], p.zbins))
        return self._ranges_host

    CSR_xrange_transformed = property(lambda self: self._ranges()[0])
    CSR_zrange = property(lambda self: self._ranges()[1])

    def _mesh_arrays(self):
        if self._mesh_host is None:
            xm, zm = np.meshgrid(self.CSR_xrange_transformed, self.CSR_zrange, indexing="ij")
            zm = zm.flatten()
            self._mesh_host = (xm.flatten() + np.polyval(np.array(self._mesh_slope), zm), zm)
        return self._mesh_host

    CSR_xmesh = property(lambda self: self._mesh_arrays()[0])
    CSR_zmesh = property(lambda self: self._mesh_arrays()[1])

    def _wake_params(self):
        b, ip = self.beam, self.integration_params
        return ops.wake_params(t=b.position, sigma_x=b._sigma_x, sigma_z=b._sigma_z, slope0=b._slope[0],
                               mean_x=b._mean_x, formation_window=ip.n_formation_length * self.formation_length,
                               csr_scaling=self.CSR_scaling, nx=ip.xbins, nz=ip.zbins, skip=getattr(self, "skip_mode", "auto"))

    def _xgroup_plan(self, wp):
        """The x-group plan of this step, or None when the point kernel serves it (dfcsr_wake_xgroup_plan)."""
        if self.wake_mapping == "point":
            self.last_wake_mapping = "point"
            return None
        xa, za = self._mesh_axes
        plan = ops.wake_xgroup_plan(self.DF_tracker.history, wp, xa, za)
        self.last_wake_mapping = "xgroup" if plan.n_groups > 0 else "point"
        return plan if plan.n_groups > 0 else None

    def calculate_2D_CSR(self):
        """CSR.py:397-418: the whole mesh in one launch; results stay on the device
        (`dE_dct`, `x_kick` are (xbins, zbins) CUDA tensors; `.cpu().numpy()` for host copies)."""
        p = self.CSR_params
        lat = self.lattice.device_tables(self.device)
        xa, za = self._mesh_axes
        wp = self._wake_params()
        plan = self._xgroup_plan(wp)
        if plan is not None:
            de, kick = ops.wake_grid_xgroups(self.DF_tracker.history, lat, wp, xa, za, *self._mesh_slope, plan=plan,
                                             counters=getattr(self, "wake_counters", None))
        else:
            de, kick = ops.wake_grid(self.DF_tracker.history, lat, wp, xa, za, *self._mesh_slope,
                                     counters=getattr(self, "wake_counters", None))
        self.dE_dct = de.reshape(p.xbins, p.zbins)
        self.x_kick = kick.reshape(p.xbins, p.zbins)

    def calculate_2D_CSR_parallel(self):
        """CSR.py:420-451: this rank's contiguous block.  With peer memory the wake kernel stores its results into
        the grids of all ranks itself (fused exchange) and one barrier publishes them; otherwise one NCCL all-gather
        of [dE | kick] replaces the two Allgatherv calls."""
        p = self.CSR_params
        n = p.xbins * p.zbins
        lat = self.lattice.device_tables(self.device)
        xa, za = self._mesh_axes
        peer = getattr(self, "_peer_grid", None)
        wp = self._wake_params()
        plan = self._xgroup_plan(wp)
        if plan is not None:
            # groups rank, rank + P, ... of the x-group mapping (every rank derives the same plan; the grid bits do not
            # depend on the split).  With peer memory the kernel stores into all ranks' grids; otherwise every rank fills
            # its groups of a full grid and ONE all-gather + a per-point selection of the owning rank assembles them.
            if peer is not None:
                grid, ptrs, handle = peer.next()
                ops.wake_grid_xgroups(self.DF_tracker.history, lat, wp, xa, za, *self._mesh_slope, plan=plan,
                                      group_first=self.rank, group_stride=self.world_size, peer_ptrs=ptrs,
                                      counters=getattr(self, "wake_counters", None))
                handle.barrier(channel=0)
                full = grid.clone()
            else:
                mine = torch.zeros((2, n), dtype=torch.float64, device=self.device)
                ops.wake_grid_xgroups(self.DF_tracker.history, lat, wp, xa, za, *self._mesh_slope, plan=plan,
                                      group_first=self.rank, group_stride=self.world_size, out=mine,
                                      counters=getattr(self, "wake_counters", None))
                key = (p.xbins, p.zbins, plan.group_points)
                if getattr(self, "_xgroup_owner_key", None) != key:
                    self._xgroup_owner = dist_utils.xgroup_owner(p.xbins, p.zbins, self.world_size, self.device,
                                                                 plan.group_points)
                    self._xgroup_owner_key = key
                full = dist_utils.all_gather_select(mine, self._xgroup_owner)
            self.dE_dct = full[0].reshape(p.xbins, p.zbins)
            self.x_kick = full[1].reshape(p.xbins, p.zbins)
            return
        if peer is not None:
            grid, ptrs, handle = peer.next()
            if self.mesh_split == "interleaved":      # points rank, rank + P, ...: equal work on every rank
                first, stride = self.rank, self.world_size
                count = (n - self.rank + self.world_size - 1) // self.world_size
            else:                                     # the reference's contiguous blocks (CSR.py:121-125)
                first, stride, count = int(self.displ[self.rank]), 1, int(self.count[self.rank])
            try:
                ops.wake_grid_peers(self.DF_tracker.history, lat, wp, xa, za, *self._mesh_slope,
                                    first=first, count=count, stride=stride, peer_ptrs=ptrs,
                                    counters=getattr(self, "wake_counters", None))
            except _lib.DfcsrError as e:
                # The launch was refused before anything ran (e.g. a shared-memory or slice-size limit).  Every rank
                # launches with the same history geometry and scalars, so every rank lands here in the same step:
                # the switch to the NCCL all-gather below is collective without any extra exchange.
                self._log(f"fused K4 exchange refused ({e}); using the NCCL all-gather from now on")
                self._peer_grid = peer = None
        if peer is not None:
            handle.barrier(channel=0)
            full = grid.clone()          # value semantics like the NCCL path: the mapped grid is rewritten two steps later
        else:
            pad = max(self.count)
            send = torch.zeros((2, pad), dtype=torch.float64, device=self.device)
            ops.wake_grid(self.DF_tracker.history, lat, self._wake_params(), xa, za, *self._mesh_slope,
                          first=int(self.displ[self.rank]), count=int(self.count[self.rank]), out=send,
                          counters=getattr(self, "wake_counters", None))
            full = dist_utils.all_gather_blocks(send, self.count, n)
        self.dE_dct = full[0].reshape(p.xbins, p.zbins)
        self.x_kick = full[1].reshape(p.xbins, p.zbins)

    def get_CSR_wake(self, s, x, debug=False):
        """CSR.py:454-602 for one point.  debug=True returns the per-region node and integrand arrays
        (list of dicts with xp, sp, integrand_z, integrand_x) like the reference's debug tuple."""
        lat = self.lattice.device_tables(self.device)
        wp = self._wake_params()
        if debug:
            return ops.wake_point_debug(self.DF_tracker.history, lat, wp, s, x)
        pt = torch.tensor([[x], [s - self.beam.position]], dtype=torch.float64, device=self.device)
        de, kick = ops.wake_mesh(self.DF_tracker.history, lat, wp, pt[0], pt[1])
        return float(de[0]), float(kick[0])

    # ------------------------------------------------------------------------------- output
    # The reference's files (CSR.py:784-879), written through the same create_group / create_dataset / attrs calls with
    # the same names.  The container is HDF5 when h5py is importable, else an .npz with the same tree (outputs.py).
    def dump_beam(self, label):
        """CSR.py:784-797.  The reference writes an openPMD file through pmd_beamphysics (absent offline); here the six
        Bmad-X coordinates and the beam scalars go to one store (collective when the particles are sharded)."""
        if not getattr(self.CSR_params, "write_beam", None):
            return
        coords = self.beam.to_host() if (self.beam.shards is not None or self.rank == 0) else None   # collective if sharded
        if self.rank != 0:
            return
        path = full_path(self.CSR_params.workdir)
        os.makedirs(path, exist_ok=True)
        fn = outputs.store_path(os.path.join(path, f"{self.prefix}-particles-{label}"))
        if os.path.isfile(fn):
            os.remove(fn)
        with outputs.open_store(fn, "w") as hf:
            for k, name in enumerate(("x", "px", "y", "py", "z", "pz")):
                hf.create_dataset(name, data=coords[k])
            hf.attrs["position"] = self.beam.position
            hf.attrs["charge"] = self.beam.charge
            hf.attrs["p0c"] = self.beam.init_energy
            hf.attrs["coordinates"] = "Bmad-X canonical (x, px, y, py, z, pz)"
        self._log("Beam at position {} is written to {}".format(self.beam.position, fn))

    def write_wakes(self):
        """CSR.py:799-835: one group per CSR step, step_k/{longitudinal,transverse}/{x_grids,z_grids,dE_dct|xkicks}."""
        if self.rank != 0:
            return
        path = full_path(self.CSR_params.workdir)
        os.makedirs(path, exist_ok=True)
        fn = outputs.store_path(os.path.join(path, f"{self.prefix}-wakes"))
        if self.beam.step == 1 and os.path.isfile(fn):
            os.remove(fn)
        shape = tuple(self.dE_dct.shape)
        with outputs.open_store(fn, "a") as hf:
            step = self.beam.step
            g = hf.create_group("step_" + str(step))
            g.attrs["step"] = step
            g.attrs["position"] = self.beam.position
            g.attrs["mean_gamma"] = self.beam.init_gamma
            g.attrs["beam_energy"] = self.beam.init_energy
            g.attrs["element"] = str(self.lattice.current_element)
            g.attrs["charge"] = self.beam.charge
            g1 = g.create_group("longitudinal")
            g1.attrs["unit"] = "MeV/m"
            g1.create_dataset("x_grids", data=self.CSR_xmesh.reshape(shape))
            g1.create_dataset("z_grids", data=self.CSR_zmesh.reshape(shape))
            g1.create_dataset("dE_dct", data=self.dE_dct.cpu().numpy())
            g2 = g.create_group("transverse")
            g2.attrs["unit"] = "MeV/m"
            g2.create_dataset("x_grids", data=self.CSR_xmesh.reshape(shape))
            g2.create_dataset("z_grids", data=self.CSR_zmesh.reshape(shape))
            g2.create_dataset("xkicks", data=self.x_kick.cpu().numpy())

    def write_statistics(self):
        """CSR.py:860-879: step_positions, the reference orbit tables and the statistics dictionary (twiss/...)."""
        self._flush_statistics()
        if self.rank != 0 or not self.CSR_params.write_wakes:
            return
        path = full_path(self.CSR_params.workdir)
        os.makedirs(path, exist_ok=True)
        fn = outputs.store_path(os.path.join(path, f"{self.prefix}-statistics"))
        if os.path.isfile(fn):
            os.remove(fn)
        with outputs.open_store(fn, "w") as hf:
            hf.create_dataset(name="step_positions", data=self.lattice.steps_record, shape=self.lattice.steps_record.shape)
            hf.create_dataset(name="coords", data=self.lattice.coords)
            hf.create_dataset(name="n_vec", data=self.lattice.n_vec)
            hf.create_dataset(name="tau_vec", data=self.lattice.tau_vec)
            outputs.dict2hdf5(hf, self.statistics)
