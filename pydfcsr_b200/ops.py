"""Functional wrappers over the C ABI (include/dfcsr_b200.h) taking torch CUDA tensors.

PyTorch is used for device memory and streams only; every computation below is one of the
hand-written sm_100a kernels in ``csrc/``.  There is no CPU fallback: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import Axis, check, lib

F64 = torch.float64


def _ptr(t: torch.Tensor | None):
    """Device address of a contiguous CUDA tensor as a plain int (ctypes converts ints and None for `c_void_p` parameters
    itself: no wrapper object per argument -- a lattice step passes ~50 pointers on the host's critical path)."""
    if t is None:
        return None
    if t.is_cuda and t.is_contiguous():
        return t.data_ptr()
    if not t.is_cuda:
        raise _lib.DfcsrError("pydfcsr_b200 kernels need CUDA tensors (there is no CPU fallback)")
    raise _lib.DfcsrError("tensor must be contiguous")


def _f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != F64:
        raise _lib.DfcsrError(f"{name} must be float64")
    return t


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """The current torch stream of the current device as a raw cudaStream_t (every kernel of the library is enqueued on
    it).  torch.cuda.current_stream() costs ~10 us of Python per call and torch.cuda.current_device() ~2 us -- a dozen calls
    per lattice step sit on the host's critical path between the statistics and the wake launch -- so the raw C accessors
    are used when torch has them."""
    if _raw_stream is not None:
        return _raw_stream(_raw_device() if _raw_device is not None else torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------
# A14 beam scalars
# ---------------------------------------------------------------------------------------------
_stats_ws: dict = {}


# the statistics travel through dfcsr_mirror_to_host (no copy engine); DFCSR_STATS_MIRROR=0 or a failing call (pinned
# memory not mapped) falls back to cudaMemcpyAsync
_mirror_ok = os.environ.get("DFCSR_STATS_MIRROR", "1") != "0"


class PendingStats:
    """Result of an asynchronous statistics pass: `get()` waits for the device (once) and returns the 16 doubles.
    The pinned buffer belongs to this object until it has been read (or the object dies); only then does it go back
    to the free list, so a result that is read late can never be overwritten by a later pass."""

    def __init__(self, host_buf, event, free_list, device_stats=None):
        self._buf, self._event, self._value, self._free = host_buf, event, None, free_list
        self.device_stats = device_stats      # the 16 doubles on the device (input of get_df_from_stats)

    def _release(self):
        if self._buf is not None:
            self._free.append(self._buf)
            self._buf = None

    def get(self) -> np.ndarray:
        if self._value is None:
            self._event.synchronize()
            self._value = self._buf.numpy().copy()
            self._release()
        return self._value

    def __del__(self):
        try:
            if self._buf is not None:
                self._event.synchronize()      # the device may still be writing into the buffer
                self._release()
        except Exception:
            pass


def _centre(values, n):
    if values is None:
        return None
    return (C.c_double * n)(*[float(v) for v in values])


def beam_stats_async(x: torch.Tensor, z: torch.Tensor, pz: torch.Tensor | None = None, px: torch.Tensor | None = None,
                     centre=None, shards=None) -> PendingStats:
    """Device reductions -> 16 doubles (indices: ``_lib.S_*``) copied to pinned host memory on the current stream.
    Nothing blocks here: the host can keep enqueueing work and call ``.get()`` when it needs the numbers.
    centre: (x, z, pz) about which the first pass accumulates (same on every rank; None = zeros).
    shards: a `distributed.ParticleShards` when x, z, ... are this rank's shard of a bunch distributed over ranks; the
    result is then the statistics of the WHOLE bunch, bit-identical on every rank and to the single-GPU pass."""
    _ptr(x), _ptr(z)          # raises for non-CUDA tensors before anything is allocated
    dev = x.device
    if dev not in _stats_ws:
        # [0] reduction workspace, [1] free list of pinned result buffers (grown on demand).  Each pass gets its own
        # device result vector, so two passes in flight on different streams never share one.
        _stats_ws[dev] = [torch.zeros(lib.dfcsr_beam_stats_workspace(), dtype=torch.uint8, device=dev), []]
    ws, free = _stats_ws[dev]
    host = free.pop() if free else torch.zeros(_lib.STATS_DOUBLES, dtype=F64).pin_memory()
    d_stats = torch.empty(_lib.STATS_DOUBLES, dtype=F64, device=dev)     # all 16 entries are written by the passes
    ctr = _centre(centre, 3)
    if shards is None:
        check(lib.dfcsr_beam_stats(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(pz), _ptr(px), x.numel(), ctr, _ptr(d_stats),
                                   _ptr(ws), _stream()), "dfcsr_beam_stats")
    else:
        for p in (0, 1):
            table, peer_ptrs = shards.stats_table(p)
            check(lib.dfcsr_beam_stats_partial(p, _ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(pz), _ptr(px), x.numel(),
                                               shards.n_total, shards.first_block, shards.n_blocks, ctr, _ptr(d_stats),
                                               _ptr(table), peer_ptrs, len(peer_ptrs) if peer_ptrs is not None else 0,
                                               _stream()), "dfcsr_beam_stats_partial")
            shards.exchange(table)
            check(lib.dfcsr_beam_stats_final(p, _ptr(table), shards.n_total, ctr, int(pz is not None), int(px is not None),
                                             _ptr(d_stats), _stream()), "dfcsr_beam_stats_final")
    global _mirror_ok
    if _mirror_ok:      # one-warp store into the mapped pinned buffer: no copy engine, nothing to queue behind
        _mirror_ok = lib.dfcsr_mirror_to_host(_ptr(d_stats), C.c_void_p(host.data_ptr()), _lib.STATS_DOUBLES, _stream()) == 0
    if not _mirror_ok:
        host.copy_(d_stats, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return PendingStats(host, ev, free, d_stats)


def beam_stats(x: torch.Tensor, z: torch.Tensor, pz: torch.Tensor | None = None, px: torch.Tensor | None = None,
               centre=None, shards=None) -> np.ndarray:
    """Device reductions -> 16 doubles on the host (indices: ``_lib.S_*``).  Synchronises."""
    return beam_stats_async(x, z, pz, px, centre, shards).get()


_cov_ws: dict = {}


class PendingCov:
    """Result of an asynchronous covariance pass: `get()` waits for the device (once) and returns (means[6], cov[6, 6])."""

    def __init__(self, host_buf, event, free_list):
        self._buf, self._event, self._value, self._free = host_buf, event, None, free_list

    def get(self):
        if self._value is None:
            self._event.synchronize()
            flat = self._buf.numpy().copy()
            self._free.append(self._buf)
            self._buf = None
            cov = np.zeros((6, 6))
            cov[np.triu_indices(6)] = flat[6:]
            cov = cov + np.triu(cov, 1).T
            self._value = (flat[:6].copy(), cov)
        return self._value

    def __del__(self):
        try:
            if self._buf is not None:
                self._event.synchronize()
                self._free.append(self._buf)
        except Exception:
            pass


def beam_cov_async(coords, centre=None, shards=None) -> PendingCov:
    """Enqueue the covariance pass over the six coordinate tensors (x, px, y, py, z, pz); np.cov normalisation (ddof = 1).
    One device pass + 27 doubles mirrored to pinned host memory; nothing blocks.  centre / shards: as for beam_stats_async."""
    dev = coords[0].device
    if dev not in _cov_ws:
        _cov_ws[dev] = [torch.zeros(lib.dfcsr_beam_cov_workspace(), dtype=torch.uint8, device=dev), []]
    ws, free = _cov_ws[dev]
    host = free.pop() if free else torch.zeros(27, dtype=F64).pin_memory()
    d_out = torch.empty(27, dtype=F64, device=dev)
    ptrs = [_ptr(_f64(c, "coords")) for c in coords]
    ctr = _centre(centre, 6)
    if shards is None:
        check(lib.dfcsr_beam_cov(*ptrs, coords[0].numel(), ctr, _ptr(d_out), _ptr(ws), _stream()), "dfcsr_beam_cov")
    else:
        table, peer_ptrs = shards.cov_table()
        check(lib.dfcsr_beam_cov_partial(*ptrs, coords[0].numel(), shards.n_total, shards.first_block, shards.n_blocks, ctr,
                                         _ptr(table), peer_ptrs, len(peer_ptrs) if peer_ptrs is not None else 0, _stream()),
              "dfcsr_beam_cov_partial")
        shards.exchange(table)
        check(lib.dfcsr_beam_cov_final(_ptr(table), shards.n_total, ctr, _ptr(d_out), _stream()), "dfcsr_beam_cov_final")
    global _mirror_ok
    if _mirror_ok:
        _mirror_ok = lib.dfcsr_mirror_to_host(_ptr(d_out), C.c_void_p(host.data_ptr()), 27, _stream()) == 0
    if not _mirror_ok:
        host.copy_(d_out, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return PendingCov(host, ev, free)


def beam_cov(coords, centre=None, shards=None) -> tuple[np.ndarray, np.ndarray]:
    """(means[6], cov[6, 6]); synchronises.  See beam_cov_async."""
    return beam_cov_async(coords, centre, shards).get()


# ---------------------------------------------------------------------------------------------
# K1 deposit
# ---------------------------------------------------------------------------------------------
def deposit_cic(x, z, px, nx, x_start, x_end, nz, z_start, z_end, mode=0, out=None):
    """(count, vxsum): the two CIC grids of deposit.py:172-182 in one pass.  mode 0/4/5: fixed-point, bit-reproducible
    (default); 1/2/3: fp64 atomics (include/dfcsr_b200.h)."""
    if out is None:
        out = torch.empty((2, nx, nz), dtype=F64, device=x.device)
    check(lib.dfcsr_deposit_cic(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(_f64(px, "px")), x.numel(),
                                nx, x_start, x_end, nz, z_start, z_end, _ptr(out[0]), _ptr(out[1]), mode,
                                _stream()), "dfcsr_deposit_cic")
    return out[0], out[1]


def deposit_cic_q(x, z, px, n_total, nx, x_start, x_end, nz, z_start, z_end, absmax_px, q_out):
    """Stage 1 of the fixed-point deposit (dfcsr_deposit_cic_q): this rank's particles into the (2, nx*nz) int64 buffer
    `q_out` at the scales of the whole bunch (n_total particles, absmax_px = max |px| over all of them)."""
    if q_out.dtype != torch.int64 or q_out.numel() < 2 * nx * nz or not q_out.is_contiguous():
        raise _lib.DfcsrError("q_out must be a contiguous int64 tensor with at least 2*nx*nz elements")
    check(lib.dfcsr_deposit_cic_q(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(_f64(px, "px")), x.numel(), int(n_total),
                                  nx, x_start, x_end, nz, z_start, z_end, float(absmax_px), _ptr(q_out), _stream()),
          "dfcsr_deposit_cic_q")
    return q_out


def deposit_cic_finish(peer_q_ptrs, n_total, nx, nz, absmax_px, out=None, device=None, count_max=None):
    """Stage 2 (dfcsr_deposit_cic_finish): sum the fixed-point buffers of all ranks (`peer_q_ptrs`: ctypes uint64 array of
    their addresses in this process) and convert to the fp64 (count, vxsum) grids.  count_max: optional 1-element int64
    CUDA tensor that receives max(count) (bit pattern of the double) for make_df."""
    if out is None:
        out = torch.empty((2, nx, nz), dtype=F64, device=device)
    check(lib.dfcsr_deposit_cic_finish(peer_q_ptrs, len(peer_q_ptrs), nx, nz, int(n_total), float(absmax_px),
                                       _ptr(out[0]), _ptr(out[1]), _ptr(count_max), _stream()), "dfcsr_deposit_cic_finish")
    return out[0], out[1]


def deposit_ngp(x, z, nx, x_start, x_end, nz, z_start, z_end):
    out = torch.empty((nx, nz), dtype=torch.int64, device=x.device)
    check(lib.dfcsr_deposit_ngp(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), x.numel(), nx, x_start, x_end,
                                nz, z_start, z_end, _ptr(out), _stream()), "dfcsr_deposit_ngp")
    return out


# ---------------------------------------------------------------------------------------------
# K2 density functions
# ---------------------------------------------------------------------------------------------
_sg_cache: dict = {}


def savgol_operators(window: int, order: int):
    """Host fp64 Savitzky-Golay operators for scipy's mode='interp' (see csrc/make_df.cu)."""
    key = (window, order)
    if key not in _sg_cache:
        if window % 2 != 1 or window < 1 or order >= window:
            raise ValueError("filter_window must be odd, positive and larger than filter_order")
        half = window // 2
        pos = np.arange(window, dtype=np.float64)
        pinv = np.linalg.pinv(np.vander(pos, order + 1, increasing=True))
        taps = (np.vander(np.array([float(half)]), order + 1, increasing=True) @ pinv)[0]
        lo = np.vander(pos[:half], order + 1, increasing=True) @ pinv if half else np.zeros((0, window))
        hi = np.vander(pos[window - half:], order + 1, increasing=True) @ pinv if half else np.zeros((0, window))
        _sg_cache[key] = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in (taps, lo, hi))
    return _sg_cache[key]


_df_ws: dict = {}
_df_need: dict = {}
_sg_dev: dict = {}


def make_df(count, vxsum, x_axis: Axis, z_axis: Axis, window: int, order: int, velocity_threshold: float,
            out: torch.Tensor | None = None, count_max: torch.Tensor | None = None):
    """(fields[5, nx, nz], scalars[8]) from the deposit grids (deposit.py:183-235).  count_max: the 1-element int64 CUDA
    tensor filled by deposit_cic_finish (saves the reduction launch)."""
    nx, nz = x_axis.n, z_axis.n
    dev = count.device
    need = _df_need.get((nx, nz))
    if need is None:
        need = _df_need[(nx, nz)] = lib.dfcsr_make_df_workspace(nx, nz)
    ws = _df_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = _df_ws[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((5, nx, nz), dtype=F64, device=dev)
    scalars = torch.empty(_lib.DF_SCALARS, dtype=F64, device=dev)
    key = (window, order, dev)
    if key not in _sg_dev:          # operators are uploaded once per (window, order, device)
        _sg_dev[key] = tuple(torch.from_numpy(a.reshape(-1).copy()).to(dev) if a.size else None
                             for a in savgol_operators(window, order))
    taps, lo, hi = _sg_dev[key]
    check(lib.dfcsr_make_df(_ptr(count), _ptr(vxsum), x_axis, z_axis, window, _ptr(taps), _ptr(lo), _ptr(hi),
                            float(velocity_threshold), _ptr(count_max), _ptr(out), _ptr(scalars), _ptr(ws), _stream()),
          "dfcsr_make_df")
    return out, scalars


def get_df(x, z, px, x_axis: Axis, z_axis: Axis, absmax_px, window: int, order: int, velocity_threshold: float,
           q_scratch: torch.Tensor, deposit_out: torch.Tensor, count_max: torch.Tensor):
    """deposit_cic_q + deposit_cic_finish + make_df of one GPU in ONE binding call (dfcsr_get_df): same kernels, same
    results, a third of the host time.  Returns (fields[5, nx, nz], scalars[8]); the deposit grids land in deposit_out."""
    nx, nz = x_axis.n, z_axis.n
    dev = x.device
    need = _df_need.get((nx, nz))
    if need is None:
        need = _df_need[(nx, nz)] = lib.dfcsr_make_df_workspace(nx, nz)
    ws = _df_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = _df_ws[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
    key = (window, order, dev)
    if key not in _sg_dev:
        _sg_dev[key] = tuple(torch.from_numpy(a.reshape(-1).copy()).to(dev) if a.size else None
                             for a in savgol_operators(window, order))
    taps, lo, hi = _sg_dev[key]
    if q_scratch.dtype != torch.int64 or q_scratch.numel() < 2 * nx * nz or tuple(deposit_out.shape) != (2, nx, nz):
        raise _lib.DfcsrError("get_df: q_scratch must hold 2*nx*nz int64 and deposit_out must be (2, nx, nz)")
    fields = torch.empty((5, nx, nz), dtype=F64, device=dev)
    scalars = torch.empty(_lib.DF_SCALARS, dtype=F64, device=dev)
    check(lib.dfcsr_get_df(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(_f64(px, "px")), x.numel(), x_axis, z_axis,
                           float(absmax_px), _ptr(q_scratch), _ptr(deposit_out[0]), _ptr(deposit_out[1]), _ptr(count_max),
                           window, _ptr(taps), _ptr(lo), _ptr(hi), float(velocity_threshold), _ptr(fields), _ptr(scalars),
                           _ptr(ws), _stream()), "dfcsr_get_df")
    return fields, scalars


def get_df_from_stats(x, z, px, d_stats: torch.Tensor, xlim: float, zlim: float, nx: int, nz: int, window: int, order: int,
                      velocity_threshold: float, q_scratch: torch.Tensor, deposit_out: torch.Tensor, count_max: torch.Tensor,
                      limits: torch.Tensor):
    """get_df enqueued BEFORE the host has the statistics (dfcsr_get_df_from_stats): grid limits and max|px| come from
    `d_stats`, the device vector of the statistics pass that is still in flight on this stream; `limits` (4 doubles on the
    device) receives the limits.  The grid shape is the caller's guess.  Returns (fields, scalars)."""
    dev = x.device
    need = _df_need.get((nx, nz))
    if need is None:
        need = _df_need[(nx, nz)] = lib.dfcsr_make_df_workspace(nx, nz)
    ws = _df_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = _df_ws[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
    key = (window, order, dev)
    if key not in _sg_dev:
        _sg_dev[key] = tuple(torch.from_numpy(a.reshape(-1).copy()).to(dev) if a.size else None
                             for a in savgol_operators(window, order))
    taps, lo, hi = _sg_dev[key]
    if q_scratch.dtype != torch.int64 or q_scratch.numel() < 2 * nx * nz or tuple(deposit_out.shape) != (2, nx, nz):
        raise _lib.DfcsrError("get_df_from_stats: q_scratch must hold 2*nx*nz int64 and deposit_out must be (2, nx, nz)")
    fields = torch.empty((5, nx, nz), dtype=F64, device=dev)
    scalars = torch.empty(_lib.DF_SCALARS, dtype=F64, device=dev)
    check(lib.dfcsr_get_df_from_stats(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(_f64(px, "px")), x.numel(), _ptr(d_stats),
                                      float(xlim), float(zlim), nx, nz, _ptr(limits), _ptr(q_scratch), _ptr(deposit_out[0]),
                                      _ptr(deposit_out[1]), _ptr(count_max), window, _ptr(taps), _ptr(lo), _ptr(hi),
                                      float(velocity_threshold), _ptr(fields), _ptr(scalars), _ptr(ws), _stream()),
          "dfcsr_get_df_from_stats")
    return fields, scalars


def _df_scratch(nx, nz, window, order, dev):
    """(workspace, taps, edge_lo, edge_hi) of make_df for this grid and filter on this device (cached)."""
    need = _df_need.get((nx, nz))
    if need is None:
        need = _df_need[(nx, nz)] = lib.dfcsr_make_df_workspace(nx, nz)
    ws = _df_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = _df_ws[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
    key = (window, order, dev)
    if key not in _sg_dev:
        _sg_dev[key] = tuple(torch.from_numpy(a.reshape(-1).copy()).to(dev) if a.size else None
                             for a in savgol_operators(window, order))
    return (ws,) + _sg_dev[key]


def get_df_from_stats_sharded(x, z, px, d_stats: torch.Tensor, xlim: float, zlim: float, nx: int, nz: int, window: int,
                              order: int, velocity_threshold: float, shards, deposit_out: torch.Tensor,
                              count_max: torch.Tensor, limits: torch.Tensor):
    """get_df_from_stats for a bunch sharded over ranks: limits kernel, this rank's fixed-point deposit, the exact integer
    sum over the ranks (peer mappings + one device-side barrier, or an NCCL all-reduce) fused with the conversion, density
    functions -- all with limits / max|px| from the device statistics, nothing waits for the host.  Collective."""
    dev = x.device
    ws, taps, lo, hi = _df_scratch(nx, nz, window, order, dev)
    st = _stream()
    check(lib.dfcsr_df_limits(_ptr(d_stats), float(xlim), float(zlim), _ptr(limits), st), "dfcsr_df_limits")
    q, ptrs = shards.q_buffer(nx * nz)
    check(lib.dfcsr_deposit_cic_q_dev(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(_f64(px, "px")), x.numel(),
                                      int(shards.n_total), nx, nz, _ptr(limits), _ptr(d_stats), _ptr(q), st),
          "dfcsr_deposit_cic_q_dev")
    shards.reduce_q(q)
    check(lib.dfcsr_deposit_cic_finish_dev(ptrs, len(ptrs), nx, nz, int(shards.n_total), _ptr(d_stats), _ptr(deposit_out[0]),
                                           _ptr(deposit_out[1]), _ptr(count_max), _stream()), "dfcsr_deposit_cic_finish_dev")
    fields = torch.empty((5, nx, nz), dtype=F64, device=dev)
    scalars = torch.empty(_lib.DF_SCALARS, dtype=F64, device=dev)
    check(lib.dfcsr_make_df_dev(_ptr(deposit_out[0]), _ptr(deposit_out[1]), nx, nz, _ptr(limits), window, _ptr(taps), _ptr(lo),
                                _ptr(hi), float(velocity_threshold), _ptr(count_max), _ptr(fields), _ptr(scalars), _ptr(ws),
                                _stream()), "dfcsr_make_df_dev")
    return fields, scalars


# ---------------------------------------------------------------------------------------------
# 2-D Savitzky-Golay operator (SGolay_filter.py:3-81)
# ---------------------------------------------------------------------------------------------
_sg2_cache: dict = {}


def sgolay2d_kernels(window_size: int, order: int) -> np.ndarray:
    """(3, w, w) host kernels: smoothing, 'col' and 'row' derivative stencils, signs as the reference passes
    them to fftconvolve (SGolay_filter.py:68-81).  Same ValueErrors as the reference (SGolay_filter.py:10-14)."""
    key = (window_size, order)
    if key not in _sg2_cache:
        if window_size % 2 == 0:
            raise ValueError("window_size must be odd")
        n_terms = (order + 1) * (order + 2) / 2.0
        if window_size ** 2 < n_terms:
            raise ValueError("order is too high for the window size")
        off = np.arange(-(window_size // 2), window_size // 2 + 1, dtype=np.float64)
        u = np.repeat(off, window_size)          # offset along axis 0 of every window sample
        v = np.tile(off, window_size)            # offset along axis 1
        design = np.stack([u ** (k - n) * v ** n for k in range(order + 1) for n in range(k + 1)], axis=1)
        fit = np.linalg.pinv(design)
        ker = np.zeros((3, window_size, window_size))
        ker[0] = fit[0].reshape(window_size, window_size)
        if order >= 1:
            ker[1] = -fit[1].reshape(window_size, window_size)
            ker[2] = -fit[2].reshape(window_size, window_size)
        _sg2_cache[key] = ker
    return _sg2_cache[key]


def sgolay2d(z: torch.Tensor, window_size: int, order: int, derivative=None):
    """sgolay2d(z, window_size, order, derivative) of the reference on a CUDA float64 (rows, cols) tensor:
    smoothed array for derivative=None, one array for 'col' / 'row', a pair for 'both'."""
    pick = {None: [0], "col": [1], "row": [2], "both": [1, 2]}
    if derivative not in pick:
        raise ValueError("derivative must be None, 'col', 'row' or 'both'")
    if derivative is not None and order < 1:
        raise ValueError("derivatives need order >= 1")
    z = _f64(z, "z")
    rows, cols = z.shape
    key = (window_size, order, derivative, z.device)
    if key not in _sg2_cache:       # stencils are uploaded once per (window, order, derivative, device)
        ker = sgolay2d_kernels(window_size, order)[pick[derivative]]
        _sg2_cache[key] = torch.from_numpy(np.ascontiguousarray(ker)).to(z.device)
    d_ker = _sg2_cache[key]
    out = torch.empty((d_ker.shape[0], rows, cols), dtype=F64, device=z.device)
    check(lib.dfcsr_sgolay2d(_ptr(z), rows, cols, window_size, _ptr(d_ker), d_ker.shape[0], _ptr(out), _stream()),
          "dfcsr_sgolay2d")
    return (out[0], out[1]) if derivative == "both" else out[0]


# ---------------------------------------------------------------------------------------------
# K3 history
# ---------------------------------------------------------------------------------------------
def voxel_format(t: torch.Tensor) -> int:
    """dfcsr_voxel_format of a slice/ring tensor: fp64 voxels are (.., 6) doubles, fp32 voxels (.., 8) floats."""
    if t.dtype == torch.float64 and t.shape[-1] == _lib.VOXEL_DOUBLES:
        return _lib.VOXEL_F64
    if t.dtype == torch.float32 and t.shape[-1] == _lib.VOXEL_FLOATS:
        return _lib.VOXEL_F32
    raise _lib.DfcsrError(f"not a voxel tensor: dtype {t.dtype}, last dimension {t.shape[-1]}")


def new_slices(shape, precision, device) -> torch.Tensor:
    """Allocate voxel storage: shape + (6,) float64 for 'fp64', shape + (8,) float32 for 'fp32'."""
    if precision == "fp64":
        return torch.empty(tuple(shape) + (_lib.VOXEL_DOUBLES,), dtype=torch.float64, device=device)
    if precision == "fp32":
        return torch.empty(tuple(shape) + (_lib.VOXEL_FLOATS,), dtype=torch.float32, device=device)
    raise ValueError("precision must be 'fp64' or 'fp32'")


def history_regrid(fields, src_x: Axis, src_z: Axis, dst_x: Axis, dst_z: Axis, fill_vx_x, slice_out, support_out=None):
    """fill_vx_x: host float, or a 1-element CUDA tensor (read on the device, no host sync).  support_out: the slot's
    (X, 2) int32 row-support table, filled by the same kernel."""
    dev_fill = fill_vx_x if isinstance(fill_vx_x, torch.Tensor) else None
    if support_out is not None and (support_out.dtype != torch.int32 or tuple(support_out.shape) != (dst_x.n, 2)
                                    or not support_out.is_contiguous()):
        raise _lib.DfcsrError("row support must be a contiguous (X, 2) int32 tensor")
    check(lib.dfcsr_history_regrid(_ptr(fields), src_x, src_z, dst_x, dst_z,
                                   0.0 if dev_fill is not None else float(fill_vx_x), _ptr(dev_fill),
                                   voxel_format(slice_out), _ptr(slice_out), _ptr(support_out), _stream()), "dfcsr_history_regrid")
    return slice_out


def history_pack(fields, slice_out=None, precision="fp64"):
    _, X, Z = fields.shape
    if slice_out is None:
        slice_out = new_slices((X, Z), precision, fields.device)
    check(lib.dfcsr_history_pack(_ptr(_f64(fields, "fields")), X, Z, voxel_format(slice_out), _ptr(slice_out), _stream()),
          "dfcsr_history_pack")
    return slice_out


def new_row_support(cap, X, device) -> torch.Tensor:
    """(cap, X, 2) int32 row hulls of the non-zero density voxels (dfcsr_history.d_row_support), all rows empty."""
    sup = torch.empty((cap, X, 2), dtype=torch.int32, device=device)
    sup[..., 0] = torch.iinfo(torch.int32).max
    sup[..., 1] = -1
    return sup


def history_row_support(slice_in, support_out):
    """Hull [z_lo, z_hi] of the voxels with non-zero density / density gradient for every row of one slice."""
    X, Z = slice_in.shape[0], slice_in.shape[1]
    if support_out.dtype != torch.int32 or tuple(support_out.shape) != (X, 2) or not support_out.is_contiguous():
        raise _lib.DfcsrError("row support must be a contiguous (X, 2) int32 tensor")
    check(lib.dfcsr_history_row_support(_ptr(slice_in), X, Z, voxel_format(slice_in), _ptr(support_out), _stream()),
          "dfcsr_history_row_support")
    return support_out


def history_unpack(slice_in, X, Z):
    out = torch.empty((5, X, Z), dtype=F64, device=slice_in.device)
    check(lib.dfcsr_history_unpack(_ptr(slice_in), X, Z, voxel_format(slice_in), _ptr(out), _stream()), "dfcsr_history_unpack")
    return out


# ---------------------------------------------------------------------------------------------
# K4 wake
# ---------------------------------------------------------------------------------------------
@dataclass
class DeviceLattice:
    """Lattice tables uploaded once per run (lattice.py:136-143)."""
    table: torch.Tensor       # (ns, 6)
    rho: torch.Tensor
    distance: torch.Tensor
    min_s: float
    delta_s: float

    @classmethod
    def upload(cls, coords, n_vec, tau_vec, rho, distance, min_s, delta_s, device):
        tab = np.concatenate([np.asarray(coords, dtype=np.float64), np.asarray(n_vec, dtype=np.float64),
                              np.asarray(tau_vec, dtype=np.float64)], axis=1)
        return cls(torch.from_numpy(np.ascontiguousarray(tab)).to(device),
                   torch.from_numpy(np.ascontiguousarray(rho, dtype=np.float64)).to(device),
                   torch.from_numpy(np.ascontiguousarray(distance, dtype=np.float64)).to(device),
                   float(min_s), float(delta_s))

    def view(self) -> _lib.Lattice:
        v = self.__dict__.get("_view")            # the tables never change after upload: build the struct once
        if v is None:
            v = self.__dict__["_view"] = _lib.Lattice(self.table.data_ptr(), self.table.shape[0], self.rho.numel(),
                                                      self.min_s, self.delta_s, self.rho.data_ptr(), self.distance.data_ptr())
        return v


@dataclass
class DeviceHistory:
    """Device-resident (t', x, z) history ring: 48-byte fp64 voxels, or 32-byte fp32 voxels in the
    optional mixed-precision mode."""
    ring: torch.Tensor        # (cap, X, Z, 6) float64 or (cap, X, Z, 8) float32
    head: int
    T: int
    min_t: float
    min_x: float
    min_z: float
    delta_t: float
    delta_x: float
    delta_z: float
    support: torch.Tensor | None = None   # (cap, X, 2) int32 row hulls of the non-zero density voxels, or None

    def view(self) -> _lib.History:
        v = self.__dict__.get("_view")            # one struct per published history (a wake step asks for it 2-3 times)
        if v is not None:
            return v
        cap, X, Z, elems = self.ring.shape
        if self.support is not None and (self.support.dtype != torch.int32 or tuple(self.support.shape) != (cap, X, 2)
                                         or not self.support.is_contiguous()):
            raise _lib.DfcsrError("row support must be a contiguous (cap, X, 2) int32 tensor")
        v = self.__dict__["_view"] = _lib.History(
            self.ring.data_ptr(), X * Z * elems, cap, self.head, self.T, X, Z, voxel_format(self.ring),
            self.min_t, self.min_x, self.min_z, self.delta_t, self.delta_x, self.delta_z,
            None if self.support is None else self.support.data_ptr())
        return v

    @classmethod
    def from_stacks(cls, stacks, min_t, min_x, min_z, delta_t, delta_x, delta_z, device, cap=None, head=0,
                    precision="fp64", row_support=True):
        """Import five host (T, X, Z) arrays in dfcsr_field order (oracle / golden histories)."""
        T, X, Z = stacks[0].shape
        cap = cap or T
        ring = new_slices((cap, X, Z), precision, device).zero_()
        support = new_row_support(cap, X, device) if row_support else None
        for k in range(T):
            fields = torch.from_numpy(np.ascontiguousarray(np.stack([s[k] for s in stacks]))).to(device)
            history_pack(fields, ring[(head + k) % cap])
            if support is not None:
                history_row_support(ring[(head + k) % cap], support[(head + k) % cap])
        return cls(ring, head, T, float(min_t), float(min_x), float(min_z), float(delta_t), float(delta_x),
                   float(delta_z), support)


def wake_params(t, sigma_x, sigma_z, slope0, mean_x, formation_window, csr_scaling, nx, nz, skip="auto") -> _lib.WakeParams:
    """skip: zero-density skipping policy, 'auto' (on for sparse history grids), 'on' or 'off' (dfcsr_skip_mode)."""
    return _lib.WakeParams(float(t), float(sigma_x), float(sigma_z), float(slope0), float(mean_x),
                           float(formation_window), float(csr_scaling), int(nx), int(nz), _lib.SKIP_MODES[skip], 0)


def wake_mesh(hist: DeviceHistory, lat: DeviceLattice, wp: _lib.WakeParams, xmesh, zmesh, first=0, count=None,
              out=None, counters=None):
    """(dE_dct, x_kick) for mesh points [first, first+count) (CSR.py:397-451)."""
    count = xmesh.numel() - first if count is None else count
    if out is None:
        out = torch.empty((2, max(count, 1)), dtype=F64, device=xmesh.device)
    hv, lv = hist.view(), lat.view()
    check(lib.dfcsr_wake_mesh(C.byref(hv), C.byref(lv), C.byref(wp), _ptr(_f64(xmesh, "xmesh")),
                              _ptr(_f64(zmesh, "zmesh")), first, count, _ptr(out[0]), _ptr(out[1]),
                              _ptr(counters), _stream()), "dfcsr_wake_mesh")
    return out[0][:count], out[1][:count]


def wake_grid(hist: DeviceHistory, lat: DeviceLattice, wp: _lib.WakeParams, x_axis: Axis, z_axis: Axis, slope, intercept,
              first=0, count=None, out=None, counters=None):
    """Like wake_mesh, with the observation mesh generated on the device from CSR_xrange_transformed /
    CSR_zrange and the chirp line (CSR.py:380-389): nothing is built or uploaded on the host."""
    n = x_axis.n * z_axis.n
    count = n - first if count is None else count
    if out is None:
        out = torch.empty((2, max(count, 1)), dtype=F64, device=hist.ring.device)
    hv, lv = hist.view(), lat.view()
    check(lib.dfcsr_wake_grid(C.byref(hv), C.byref(lv), C.byref(wp), x_axis, z_axis, float(slope), float(intercept),
                              first, count, _ptr(out[0]), _ptr(out[1]), _ptr(counters), _stream()), "dfcsr_wake_grid")
    return out[0][:count], out[1][:count]


def wake_grid_peers(hist: DeviceHistory, lat: DeviceLattice, wp: _lib.WakeParams, x_axis: Axis, z_axis: Axis, slope,
                    intercept, first, count, peer_ptrs, counters=None, stride=1):
    """K4 fused with the exchange (dfcsr_wake_grid_peers): mesh points first, first + stride, ... (count of them) are
    computed here and stored into every rank's (2, N) grid; `peer_ptrs` = ctypes array of the grids' addresses as mapped in this process."""
    hv, lv = hist.view(), lat.view()
    check(lib.dfcsr_wake_grid_peers(C.byref(hv), C.byref(lv), C.byref(wp), x_axis, z_axis, float(slope), float(intercept),
                                    int(first), int(count), int(stride), peer_ptrs, len(peer_ptrs), _ptr(counters), _stream()),
          "dfcsr_wake_grid_peers")


def wake_xgroup_plan(hist: DeviceHistory, wp: _lib.WakeParams, x_axis: Axis, z_axis: Axis) -> _lib.XGroupPlan:
    """Plan of the x-group mapping for this step (dfcsr_wake_xgroup_plan); `n_groups == 0`: use wake_grid."""
    hv = hist.view()
    plan = _lib.XGroupPlan()
    check(lib.dfcsr_wake_xgroup_plan(C.byref(hv), C.byref(wp), x_axis, z_axis, C.byref(plan)), "dfcsr_wake_xgroup_plan")
    return plan


_XGROUP_WS: dict = {}


def xgroup_workspace(device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch of the x-group kernel, one per device, grown on demand (the kernel leaves its ticket
    words zero, so the block is reusable from launch to launch on one stream)."""
    ws = _XGROUP_WS.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = _XGROUP_WS[device] = torch.zeros(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
    return ws


def wake_grid_xgroups(hist: DeviceHistory, lat: DeviceLattice, wp: _lib.WakeParams, x_axis: Axis, z_axis: Axis, slope,
                      intercept, plan: _lib.XGroupPlan, group_first=0, group_count=None, group_stride=1, out=None,
                      peer_ptrs=None, counters=None):
    """K4 with one lane per observation point (dfcsr_wake_grid_xgroups): groups group_first, group_first + stride, ...
    are computed here; results land at their mesh index of the FULL (2, N) `out` and / or of every peer grid."""
    n = x_axis.n * z_axis.n
    if group_count is None:
        group_count = (plan.n_groups - group_first + group_stride - 1) // group_stride
    dev = hist.ring.device
    if out is None and peer_ptrs is None:
        # a launch over all groups writes every mesh point: no fill needed
        whole = group_first == 0 and group_stride == 1 and group_count == plan.n_groups
        out = (torch.empty if whole else torch.zeros)((2, n), dtype=F64, device=dev)
    ws = xgroup_workspace(dev, group_count * plan.workspace_bytes_per_group)
    hv, lv = hist.view(), lat.view()
    check(lib.dfcsr_wake_grid_xgroups(C.byref(hv), C.byref(lv), C.byref(wp), x_axis, z_axis, float(slope), float(intercept),
                                      int(group_first), int(group_count), int(group_stride),
                                      _ptr(out[0]) if out is not None else None, _ptr(out[1]) if out is not None else None,
                                      peer_ptrs, len(peer_ptrs) if peer_ptrs is not None else 0, _ptr(ws), ws.numel(),
                                      _ptr(counters), _stream()), "dfcsr_wake_grid_xgroups")
    return (out[0], out[1]) if out is not None else None


def wake_uses_skipping(hist: DeviceHistory, wp: _lib.WakeParams) -> bool:
    """Whether a wake launch on this history with these beam scalars selects zero-density skipping."""
    hv = hist.view()
    rc = lib.dfcsr_wake_uses_skipping(C.byref(hv), C.byref(wp))
    if rc < 0:
        check(rc, "dfcsr_wake_uses_skipping")
    return rc == 1


def wake_point_debug(hist: DeviceHistory, lat: DeviceLattice, wp: _lib.WakeParams, s: float, x: float):
    """Integrand arrays of one point, region by region (get_CSR_wake(debug=True), CSR.py:571-600)."""
    cap = 5 * wp.nx * wp.nz
    iz = torch.zeros(cap, dtype=F64, device=hist.ring.device)
    ix = torch.zeros(cap, dtype=F64, device=hist.ring.device)
    regions = np.zeros((4, 6))
    nreg = C.c_int32(0)
    hv, lv = hist.view(), lat.view()
    check(lib.dfcsr_wake_point_debug(C.byref(hv), C.byref(lv), C.byref(wp), float(s), float(x), _ptr(iz), _ptr(ix),
                                     cap, regions.ctypes.data_as(C.c_void_p), C.byref(nreg), _stream()),
          "dfcsr_wake_point_debug")
    out = []
    base = 0
    izh, ixh = iz.cpu().numpy(), ix.cpu().numpy()
    for r in range(nreg.value):
        n_x, n_s = int(regions[r, 2]), int(regions[r, 5])
        out.append(dict(xp=np.linspace(regions[r, 0], regions[r, 1], n_x), sp=np.linspace(regions[r, 3], regions[r, 4], n_s),
                        integrand_z=izh[base:base + n_x * n_s].reshape(n_x, n_s),
                        integrand_x=ixh[base:base + n_x * n_s].reshape(n_x, n_s)))
        base += n_x * n_s
    return out


# ---------------------------------------------------------------------------------------------
# K5 kick
# ---------------------------------------------------------------------------------------------
def apply_kick(x, z, px, pz, slope, intercept, dE, kick, x_axis: Axis, z_axis: Axis, step_size, init_energy,
               transverse_on=True):
    """In-place px/pz update (beams.py:108-131)."""
    check(lib.dfcsr_apply_kick(_ptr(_f64(x, "x")), _ptr(_f64(z, "z")), _ptr(_f64(px, "px")), _ptr(_f64(pz, "pz")),
                               x.numel(), float(slope), float(intercept), _ptr(_f64(dE, "dE")), _ptr(_f64(kick, "kick")),
                               x_axis, z_axis, float(step_size), float(init_energy), int(bool(transverse_on)),
                               _stream()), "dfcsr_apply_kick")


# ---------------------------------------------------------------------------------------------
# particle transport (beams.py:101-106)
# ---------------------------------------------------------------------------------------------
def track_element(coords, element: _lib.Element, p0c: float, mc2: float) -> None:
    """Transport the six coordinate tensors (x, px, y, py, z, pz) through one element, in place (dfcsr_track_element)."""
    ptrs = [_ptr(_f64(c, "coords")) for c in coords]
    check(lib.dfcsr_track_element(*ptrs, coords[0].numel(), C.byref(element), float(p0c), float(mc2), _stream()),
          "dfcsr_track_element")



def track_linear(coords, matrix) -> None:
    """v <- M v in place for the six coordinate tensors (x, px, y, py, z, pz); `matrix` is a host (6, 6) array."""
    m = np.ascontiguousarray(matrix, dtype=np.float64).reshape(36)
    ptrs = [_ptr(_f64(c, "coords")) for c in coords]
    check(lib.dfcsr_track_linear(*ptrs, coords[0].numel(), m.ctypes.data_as(C.POINTER(C.c_double)), _stream()),
          "dfcsr_track_linear")


# ---------------------------------------------------------------------------------------------
# diagnostics
# ---------------------------------------------------------------------------------------------
def selftest_sqrt(n: int, seed: int = 0, lo_exp: float = -60.0, hi_exp: float = 8.0):
    """Bitwise comparison of K4's fused sqrt / reciprocal-sqrt with the CUDA library's sqrt.rn.f64 and rsqrt on
    n pseudo-random doubles (binary exponents in [lo_exp, hi_exp)): returns (sqrt mismatches, rsqrt mismatches)."""
    out = (C.c_uint64 * 2)(0, 0)
    check(lib.dfcsr_selftest_sqrt(int(n), int(seed), float(lo_exp), float(hi_exp), out, _stream()), "dfcsr_selftest_sqrt")
    return int(out[0]), int(out[1])
