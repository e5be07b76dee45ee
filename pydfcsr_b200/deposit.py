"""Device-resident ``DF_tracker``: same method names, arguments and published attributes as the
reference's class (deposit.py:90-426), with every array living in HBM.

    get_DF(x, z, px, t)                      deposit.py:145   K1 (CIC) + K2 (smooth/gradient) kernels
    append_DF()                              deposit.py:247   logs the raw record (device tensors)
    append_interpolant(fl, n_fl)             deposit.py:312   window pop + K3 re-grid into the ring
    build_interpolant()                      deposit.py:395   publishes metadata (no re-stack copy)

The (t', x, z) history is a ring of 48-byte voxels (`ops.DeviceHistory`); it never round-trips to
the host.  Host-visible numpy copies (`density`, `data_density_interp`, ...) are materialised
lazily for debugging/parity only.  The policy decisions (grid choice deposit.py:157-167, window
pop :265-280, re-grid/rebuild :321-361) are taken on the host from 16 device-computed scalars.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass

import ctypes as C
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import Axis

FIELD_NAMES = ("density", "density_x", "density_z", "vx", "vx_x")


@dataclass
class _Record:
    """One logged density-function record (deposit.py:252): device field stack + grid + scalars."""
    fields: torch.Tensor      # (5, xb, zb)
    scalars: torch.Tensor     # (8,) device; [4] = mean(vx_x)
    x_axis: Axis
    z_axis: Axis
    t: float
    sigma_x: float
    sigma_z: float
    xmean: float
    zmean: float


class DF_tracker:
    def __init__(self, input_dic=None, device=None, deposit_mode=0, precision="fp64", shards=None):
        """precision: storage format of the history ring — 'fp64' (parity mode, default) or 'fp32'
        (optional mixed-precision mode: fields stored/blended in fp32, everything else fp64)."""
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        self.precision = precision
        self.configure_params(**(input_dic or {}))
        self.device = torch.device(device if device is not None else "cuda")
        self.deposit_mode = deposit_mode
        self.sigma_x = self.sigma_z = None
        self.xmean = self.zmean = None
        self.start_time = 0.0
        self.t = 0.0
        self.end_time = 0.0
        self._current: _Record | None = None
        # log of raw records (kept on the device for rebuilds)
        self.DF_log = deque([])
        self.time_log = deque([])
        self.sigma_x_log = deque([])
        self.sigma_z_log = deque([])
        # interpolant state
        self.sigma_x_interp = None
        self.sigma_z_interp = None
        self.time_interp = deque([])
        self.interp_start = 0
        self.x_grid_interp = None      # numpy linspace (host metadata only)
        self.z_grid_interp = None
        self._x_axis_interp = None
        self._z_axis_interp = None
        self._ring = None              # (cap, X, Z, 6) device tensor (a view of _ring_storage)
        self._ring_storage = None
        self._support_storage = None
        self._support = None           # (cap, X, 2) int32 row hulls of the non-zero density voxels, slot-aligned with the ring
        self._head = 0
        self.history: ops.DeviceHistory | None = None
        self.rebuilds = 0
        self._deposit_scratch = None
        self._q_scratch = None
        self._count_max = None
        self._spec_shape = None        # (xb, zb, window) of the last get_DF: the guess of the next prefetch_DF
        self._spec = None              # what prefetch_DF has enqueued, until get_DF adopts or discards it
        self._limits = None
        self.prefetch_hits = 0         # get_DF calls served by a prefetch (diagnostics)
        self.prefetch = os.environ.get("DFCSR_PREFETCH_DF", "1") != "0"
        self.prefetch_shards = os.environ.get("DFCSR_PREFETCH_DF_SHARDS", "1") != "0"
        self.shards = shards           # distributed.ParticleShards when x, z, px are this rank's shard of the bunch

    def configure_params(self, xbins=100, zbins=100, xlim=5, zlim=5, filter_order=0, filter_window=0,
                         velocity_threhold=5, upper_limit=None):
        self.xbins = xbins
        self.zbins = zbins
        self.xlim = xlim
        self.zlim = zlim
        self.velocity_threhold = velocity_threhold
        self.filter_order = filter_order
        self.filter_window = filter_window
        self.upper_limit = upper_limit

    # ------------------------------------------------------------------------------- get_DF
    def _as_device(self, a):
        if isinstance(a, torch.Tensor):
            if a.is_cuda and a.dtype == torch.float64 and a.is_contiguous() and a.device == self.device:
                return a                      # the usual case (Beam's coordinate arrays): nothing to do, nothing to call
            return a.to(self.device, torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device, non_blocking=True)

    def prefetch_DF(self, beam):
        """Optional, between `Beam.update_status()` and `get_DF`: enqueue the deposit and the density functions of the
        beam's current state NOW, behind the statistics pass that is still in flight, with the grid limits and max|px|
        read from the device (dfcsr_get_df_from_stats).  The grid shape (deposit.py:157-167 derives it from the statistics)
        is a guess: the previous step's.  `get_DF` checks the guess when the statistics have arrived and adopts the
        result -- same bits as computing it then -- or discards it.  Saves the GPU the wait for the host round trip
        statistics -> Python -> five launches at the start of every lattice step.  No-op when there is nothing to go on."""
        self._spec = None
        pending = getattr(beam, "_pending_stats", None)
        if (not self.prefetch or self._spec_shape is None or self.deposit_mode != 0 or pending is None
                or pending.device_stats is None or pending._value is not None or beam.px is None):
            return
        if self.shards is not None and not self.prefetch_shards:
            return
        xb, zb, window = self._spec_shape
        x, z, px = beam.x, beam.z, beam.px
        if not (x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and z.is_contiguous() and px.is_contiguous()):
            return
        if self._deposit_scratch is None or self._deposit_scratch.shape[1:] != (xb, zb):
            self._deposit_scratch = torch.empty((2, xb, zb), dtype=torch.float64, device=self.device)
        if self._count_max is None:
            self._count_max = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self._q_scratch is None or self._q_scratch.numel() < 2 * xb * zb:
            self._q_scratch = torch.empty(2 * xb * zb, dtype=torch.int64, device=self.device)
        if self._limits is None:
            self._limits = torch.empty(4, dtype=torch.float64, device=self.device)
        if self.shards is not None:
            # every rank holds the same statistics, hence the same guess, and takes the same decision in get_DF: the
            # barriers of the sharded deposit stay matched across ranks whether the result is adopted or not
            fields, scalars = ops.get_df_from_stats_sharded(x, z, px, pending.device_stats, self.xlim, self.zlim, xb, zb, window,
                                                            self.filter_order, self.velocity_threhold, self.shards,
                                                            self._deposit_scratch, self._count_max, self._limits)
            self._spec = (pending, x.data_ptr(), z.data_ptr(), px.data_ptr(), (xb, zb, window), fields, scalars)
            return
        fields, scalars = ops.get_df_from_stats(x, z, px, pending.device_stats, self.xlim, self.zlim, xb, zb, window,
                                                self.filter_order, self.velocity_threhold, self._q_scratch,
                                                self._deposit_scratch, self._count_max, self._limits)
        self._spec = (pending, x.data_ptr(), z.data_ptr(), px.data_ptr(), (xb, zb, window), fields, scalars)

    def get_DF(self, x, z, px, t, stats=None):
        """x, z, px: CUDA tensors (host arrays are uploaded).  `stats` may carry the 16 beam scalars
        already computed by Beam.update_status so that the reduction kernels run once per step."""
        x, z, px = self._as_device(x), self._as_device(z), self._as_device(px)
        if stats is None:
            stats = ops.beam_stats(x, z, None, px, shards=self.shards)
        sigma_x, sigma_z = float(stats[_lib.S_SIGMA_X]), float(stats[_lib.S_SIGMA_Z])
        xmean, zmean = float(stats[_lib.S_MEAN_X]), float(stats[_lib.S_MEAN_Z])
        self.sigma_x, self.sigma_z, self.xmean, self.zmean = sigma_x, sigma_z, xmean, zmean
        frac = sigma_x / float(stats[_lib.S_SLICE_SIGMA_X])            # deposit.py:157-159
        if frac > 5:
            xb, zb, window = self.xbins, self.zbins, self.filter_window
        else:
            xb, zb, window = 100, 100, 5                               # deposit.py:164-167
        x_lo, x_hi = xmean - self.xlim * sigma_x, xmean + self.xlim * sigma_x
        z_lo, z_hi = zmean - self.zlim * sigma_z, zmean + self.zlim * sigma_z
        self._spec_shape = (xb, zb, window)
        spec, self._spec = self._spec, None
        if (spec is not None and spec[0]._value is stats and spec[1:4] == (x.data_ptr(), z.data_ptr(), px.data_ptr())
                and spec[4] == (xb, zb, window)):
            # prefetch_DF has already enqueued exactly this computation (limits taken from the device statistics with the
            # roundings of the two lines above): adopt it
            x_axis, z_axis = Axis.make(x_lo, x_hi, xb), Axis.make(z_lo, z_hi, zb)
            self._current = _Record(spec[5], spec[6], x_axis, z_axis, t, sigma_x, sigma_z, xmean, zmean)
            self.t = t
            self.prefetch_hits += 1
            return
        if self._deposit_scratch is None or self._deposit_scratch.shape[1:] != (xb, zb):
            self._deposit_scratch = torch.empty((2, xb, zb), dtype=torch.float64, device=self.device)
        absmax = float(stats[_lib.S_ABSMAX_PX]) if len(stats) > _lib.S_ABSMAX_PX else -1.0
        cmax = None                     # device slot with max(count) when the deposit delivers it
        if self._count_max is None:
            self._count_max = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self.shards is not None and self.deposit_mode == 0:
            # particles sharded over ranks: fixed-point deposit of this rank's shard, then the exact integer sum over the
            # ranks fused with the conversion to fp64 (reads all ranks' buffers over NVLink; NCCL all-reduce otherwise)
            if not absmax >= 0.0:
                raise _lib.DfcsrError("the sharded deposit needs max|px| from the statistics pass (pass px to beam_stats)")
            q, ptrs = self.shards.q_buffer(xb * zb)
            ops.deposit_cic_q(x, z, px, self.shards.n_total, xb, x_lo, x_hi, zb, z_lo, z_hi, absmax, q)
            self.shards.reduce_q(q)
            count, vxsum = ops.deposit_cic_finish(ptrs, self.shards.n_total, xb, zb, absmax, out=self._deposit_scratch,
                                                  count_max=self._count_max)
            cmax = self._count_max
        elif self.deposit_mode == 0 and absmax >= 0.0:
            # the same two deposit stages and the density functions on one GPU, in ONE binding call (max|px| came with
            # the statistics: no separate reduction pass over px)
            if self._q_scratch is None or self._q_scratch.numel() < 2 * xb * zb:
                self._q_scratch = torch.empty(2 * xb * zb, dtype=torch.int64, device=self.device)
            x_axis, z_axis = Axis.make(x_lo, x_hi, xb), Axis.make(z_lo, z_hi, zb)
            fields, scalars = ops.get_df(x, z, px, x_axis, z_axis, absmax, window, self.filter_order, self.velocity_threhold,
                                         self._q_scratch, self._deposit_scratch, self._count_max)
            self._current = _Record(fields, scalars, x_axis, z_axis, t, sigma_x, sigma_z, xmean, zmean)
            self.t = t
            return
        else:
            count, vxsum = ops.deposit_cic(x, z, px, xb, x_lo, x_hi, zb, z_lo, z_hi, mode=self.deposit_mode,
                                           out=self._deposit_scratch)
        x_axis, z_axis = Axis.make(x_lo, x_hi, xb), Axis.make(z_lo, z_hi, zb)
        fields, scalars = ops.make_df(count, vxsum, x_axis, z_axis, window, self.filter_order, self.velocity_threhold,
                                      count_max=cmax)
        self._current = _Record(fields, scalars, x_axis, z_axis, t, sigma_x, sigma_z, xmean, zmean)
        self.t = t

    # host-visible views of the current record (parity/debug; each access copies device -> host)
    def _field(self, k):
        return self._current.fields[k].cpu().numpy()

    density = property(lambda self: self._field(0))
    density_x = property(lambda self: self._field(1))
    density_z = property(lambda self: self._field(2))
    vx = property(lambda self: self._field(3))
    vx_x = property(lambda self: self._field(4))

    @property
    def x_grids(self):
        a = self._current.x_axis
        return np.linspace(a.start, a.stop, a.n)

    @property
    def z_grids(self):
        a = self._current.z_axis
        return np.linspace(a.start, a.stop, a.n)

    # ------------------------------------------------------------------------------- log
    def append_DF(self):
        rec = self._current
        self.DF_log.append(rec)
        self.time_log.append(rec.t)
        self.sigma_x_log.append(rec.sigma_x)
        self.sigma_z_log.append(rec.sigma_z)
        self.end_time = rec.t

    def pop_left_DF(self, new_start_time):
        while self.start_time < new_start_time:                        # deposit.py:265-270
            self.DF_log.popleft()
            self.time_log.popleft()
            self.sigma_x_log.popleft()
            self.sigma_z_log.popleft()
            self.start_time = self.time_log[0]
        while self.interp_start < new_start_time:                      # deposit.py:273-280
            self.time_interp.popleft()
            self._head = (self._head + 1) % self._ring.shape[0]
            self.interp_start = self.time_interp[0]

    def pop_right_DF(self):
        self.DF_log.pop()
        self.time_log.pop()
        self.sigma_x_log.pop()
        self.sigma_z_log.pop()
        self.end_time = self.time_log[-1]

    def pop_right_interpolant(self):
        """Drop the newest slice of both the raw log and the interpolant (the ring slot is simply
        reused by the next push).  Lets a caller re-run a step from the same history state."""
        self.pop_right_DF()
        self.time_interp.pop()

    # ------------------------------------------------------------------------------- ring
    def _slot_index(self, k):
        return (self._head + k) % self._ring.shape[0]

    def _slot(self, k):
        return self._ring[self._slot_index(k)]

    def _ensure_ring(self, X, Z, need):
        ring = self._ring
        if ring is None or ring.shape[1] != X or ring.shape[2] != Z or ring.shape[0] < need:
            cap = self._capacity_for(need)
            # A rebuild changes the slice shape almost every time (29 rebuilds in the 133 steps of the bundled chicane), and
            # a fresh device allocation of up to several GB costs about a millisecond each way: the ring is a VIEW of one
            # storage block that only ever grows.
            elems = _lib.VOXEL_DOUBLES if self.precision == "fp64" else _lib.VOXEL_FLOATS
            words = cap * X * Z * elems
            if self._ring_storage is None or self._ring_storage.numel() < words:
                self._ring = None      # release the old block before allocating the new one
                self._ring_storage = None
                self._ring_storage = ops.new_slices((words // elems + 1,), self.precision, self.device).view(-1)
            self._ring = self._ring_storage[:words].view(cap, X, Z, elems)
            sup_words = cap * X * 2
            if self._support_storage is None or self._support_storage.numel() < sup_words:
                self._support_storage = torch.empty(sup_words + 2 * X, dtype=torch.int32, device=self.device)
            self._support = self._support_storage[:sup_words].view(cap, X, 2)
            self._support[..., 0] = torch.iinfo(torch.int32).max
            self._support[..., 1] = -1
            self._head = 0
            return True
        return False

    @staticmethod
    def _capacity_for(need):
        """Ring slots for a window of `need` slices: a quarter of headroom (at least 4 slots), not a power of two --
        a 2000 x 2000 slice is 192 MB, so a 130-slice window must not allocate 512 slots."""
        need = max(int(need), 1)
        return max(16, need + max(4, need // 4))

    def _grow_ring(self):
        """Ring full on a non-rebuild push: enlarge it, keep the window order.  The window is copied as its two
        contiguous segments (head .. end of the ring, start .. head), so the peak is old ring + new ring."""
        old, T = self._ring, len(self.time_interp) - 1
        cap_old = old.shape[0]
        cap = self._capacity_for(T + 1 + max(4, T // 4))
        storage = torch.empty(cap * old[0].numel() + old[0].numel(), dtype=old.dtype, device=self.device)
        new = storage[:cap * old[0].numel()].view((cap,) + tuple(old.shape[1:]))
        sup_storage = torch.empty(cap * old.shape[1] * 2 + 2 * old.shape[1], dtype=torch.int32, device=self.device)
        sup = sup_storage[:cap * old.shape[1] * 2].view(cap, old.shape[1], 2)
        sup[..., 0] = torch.iinfo(torch.int32).max
        sup[..., 1] = -1
        first = min(T, cap_old - self._head)                 # slices from head to the end of the old ring
        new[:first].copy_(old[self._head:self._head + first])
        sup[:first].copy_(self._support[self._head:self._head + first])
        if T > first:                                        # the wrapped part
            new[first:T].copy_(old[:T - first])
            sup[first:T].copy_(self._support[:T - first])
        self._ring, self._support, self._head = new, sup, 0
        self._ring_storage, self._support_storage = storage, sup_storage

    def _regrid(self, rec: _Record, idx):
        """Re-grid one logged record into ring slot `idx` and refresh that slot's row support."""
        ops.history_regrid(rec.fields, rec.x_axis, rec.z_axis, self._x_axis_interp, self._z_axis_interp,
                           rec.scalars[4:5], self._ring[idx], self._support[idx])

    def append_interpolant(self, formation_length, n_formation_length):
        start_point = max(0, self.end_time - n_formation_length * formation_length)    # deposit.py:313
        if self._ring is not None:
            self.pop_left_DF(new_start_time=start_point)
        else:                                      # very first call: nothing in the interpolant yet
            while self.start_time < start_point:
                self.DF_log.popleft(); self.time_log.popleft()
                self.sigma_x_log.popleft(); self.sigma_z_log.popleft()
                self.start_time = self.time_log[0]
        rec = self._current
        if (self.sigma_x_interp and self.sigma_z_interp
                and 2 > rec.sigma_x / self.sigma_x_interp > 1 / 2
                and 2 > rec.sigma_z / self.sigma_z_interp > 1 / 2):
            self.time_interp.append(rec.t)
            if len(self.time_interp) > self._ring.shape[0]:
                self._grow_ring()
            self._regrid(rec, self._slot_index(len(self.time_interp) - 1))
            return False
        # rebuild (deposit.py:339-390)
        print("start reinterpolation. number of slice", str(len(self.time_log)))
        max_sx, min_sx = max(self.sigma_x_log), min(self.sigma_x_log)
        max_sz, min_sz = max(self.sigma_z_log), min(self.sigma_z_log)
        xbins = int(500 * (max_sx / min_sx))
        zbins = int(500 * (max_sz / min_sz))
        if isinstance(self.upper_limit, int):
            xbins, zbins = min(xbins, self.upper_limit), min(zbins, self.upper_limit)
        print("xbins =", xbins, " zbins = ", zbins)
        self.sigma_x_interp, self.sigma_z_interp = max_sx, max_sz
        xa = Axis.make(rec.xmean - 5 * max_sx, rec.xmean + 5 * max_sx, xbins)
        za = Axis.make(rec.zmean - 5 * max_sz, rec.zmean + 5 * max_sz, zbins)
        self._x_axis_interp, self._z_axis_interp = xa, za
        self.x_grid_interp = np.linspace(xa.start, xa.stop, xbins)
        self.z_grid_interp = np.linspace(za.start, za.stop, zbins)
        self.time_interp = deque(self.time_log)
        self._ensure_ring(xbins, zbins, len(self.DF_log))
        self._head = 0
        for k, r in enumerate(self.DF_log):
            self._regrid(r, k)
        self.rebuilds += 1
        return True

    def build_interpolant(self):
        """Publish the metadata the wake kernel consumes (deposit.py:416-421).  The five
        np.array(deque) copies of the reference (deposit.py:422-426) have no counterpart: the ring
        already is the stack."""
        T = len(self.time_interp)
        if T < 2:
            raise ValueError("build_interpolant needs at least two time slices (delta_x would be 0/0)")
        self.min_x, self.max_x = self.time_interp[0], self.time_interp[-1]
        self.min_y, self.max_y = self.x_grid_interp[0], self.x_grid_interp[-1]
        self.min_z, self.max_z = self.z_grid_interp[0], self.z_grid_interp[-1]
        self.delta_x = (self.max_x - self.min_x) / (T - 1)
        self.delta_y = (self.max_y - self.min_y) / (self.x_grid_interp.shape[0] - 1)
        self.delta_z = (self.max_z - self.min_z) / (self.z_grid_interp.shape[0] - 1)
        self.history = ops.DeviceHistory(self._ring, self._head, T, float(self.min_x), float(self.min_y),
                                         float(self.min_z), float(self.delta_x), float(self.delta_y),
                                         float(self.delta_z), self._support)

    # lazily materialised host copies of the (T, X, Z) stacks, reference attribute names
    def _stack(self, k):
        T = len(self.time_interp)
        _, X, Z, _ = self._ring.shape
        out = np.empty((T, X, Z))
        for i in range(T):
            out[i] = ops.history_unpack(self._slot(i), X, Z)[k].cpu().numpy()
        return out

    data_density_interp = property(lambda self: self._stack(0))
    data_density_x_interp = property(lambda self: self._stack(1))
    data_density_z_interp = property(lambda self: self._stack(2))
    data_vx_interp = property(lambda self: self._stack(3))
    data_vx_x_interp = property(lambda self: self._stack(4))
