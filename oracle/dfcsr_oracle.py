"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy + numba, fp64) of the pyDFCSR hot path.

This file is the *oracle* for the CUDA path in ``pydfcsr_b200``: it is imported by
``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline legs, never by the
product.  Every function cites the reference lines (``/root/reference/pyDFCSR_2D/...``) whose
behaviour it restates.  It is written from the algorithm's description in SURVEY.md §3/§8, not
copied: different decomposition (explicit edge operators for Savitzky-Golay, explicit region lists
for the quadrature, one fused 5-field gather), same arithmetic.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` runs the unmodified reference
(imported through ``oracle/refstub.py``) beside every function below on seeded inputs in the build
container, and ``tests/golden/*.npz`` (made by ``oracle/make_golden.py`` from the reference) pin it
on the GPU box where the reference does not exist.
"""
from __future__ import annotations

import math
from collections import deque
from dataclasses import dataclass, field

import numpy as np

try:  # numba mirrors the reference's own acceleration of the particle/gather loops
    from numba import njit
except Exception:  # pragma: no cover - numba is in the image; keep a pure-python escape hatch
    def njit(*a, **k):
        def wrap(f):
            return f
        return wrap if not (a and callable(a[0])) else a[0]


# --------------------------------------------------------------------------------------------
# A1  CIC deposit (deposit.py:42-87) and the NGP extension (no reference code)
# --------------------------------------------------------------------------------------------
@njit(cache=False)
def _cic2d_loop(q1, q2, w, out, s1, inv1, s2, inv2):
    n1, n2 = out.shape
    for p in range(q1.shape[0]):
        c1 = (q1[p] - s1) * inv1
        c2 = (q2[p] - s2) * inv2
        i = int(math.floor(c1))
        j = int(math.floor(c2))
        a_lo = 1.0 - (c1 - i)
        b_lo = 1.0 - (c2 - j)
        wp = w[p]
        if 0 <= i < n1:
            if 0 <= j < n2:
                out[i, j] += wp * a_lo * b_lo
            if 0 <= j + 1 < n2:
                out[i, j + 1] += wp * a_lo * (1.0 - b_lo)
        if 0 <= i + 1 < n1:
            if 0 <= j < n2:
                out[i + 1, j] += wp * (1.0 - a_lo) * b_lo
            if 0 <= j + 1 < n2:
                out[i + 1, j + 1] += wp * (1.0 - a_lo) * (1.0 - b_lo)


def cic_deposit_2d(q1, q2, w, nb1, start1, end1, nb2, start2, end2):
    """Cloud-in-cell histogram; bin spacing is (end-start)/nbins (deposit.py:55-58), each of the
    four corner updates is guarded separately (deposit.py:76-85)."""
    out = np.zeros((nb1, nb2))
    inv1 = 1.0 / ((end1 - start1) / nb1)
    inv2 = 1.0 / ((end2 - start2) / nb2)
    _cic2d_loop(np.ascontiguousarray(q1, dtype=np.float64), np.ascontiguousarray(q2, dtype=np.float64),
                np.ascontiguousarray(w, dtype=np.float64), out, float(start1), inv1, float(start2), inv2)
    return out


def ngp_deposit_2d(q1, q2, nb1, start1, end1, nb2, start2, end2):
    """Nearest-grid-point counts (int64).  NOT in the reference: index rule
    i = floor((q-start)/spacing + 0.5) with the CIC spacing; +1 iff both indices are in range."""
    inv1 = 1.0 / ((end1 - start1) / nb1)
    inv2 = 1.0 / ((end2 - start2) / nb2)
    i = np.floor((np.asarray(q1, dtype=np.float64) - start1) * inv1 + 0.5).astype(np.int64)
    j = np.floor((np.asarray(q2, dtype=np.float64) - start2) * inv2 + 0.5).astype(np.int64)
    ok = (i >= 0) & (i < nb1) & (j >= 0) & (j < nb2)
    out = np.zeros((nb1, nb2), dtype=np.int64)
    np.add.at(out, (i[ok], j[ok]), 1)
    return out


# --------------------------------------------------------------------------------------------
# A2  separable Savitzky-Golay smoothing with scipy's mode='interp' edge fit
#     (deposit.py:195-199,215-224 -> scipy/signal/_savitzky_golay.py:244-258,261)
# --------------------------------------------------------------------------------------------
def savgol_operators(window: int, order: int):
    """FIR taps for the interior and the two (window//2 x window) polynomial edge operators.

    scipy fits a degree-``order`` least-squares polynomial to the first/last ``window`` samples
    and evaluates it on the first/last ``window//2`` positions; both the taps and the edge fit are
    linear maps of the data that depend on (window, order) only."""
    if window % 2 != 1 or window < 1:
        raise ValueError("window must be a positive odd integer")
    if order >= window:
        raise ValueError("order must be less than window")
    half = window // 2
    pos = np.arange(window, dtype=np.float64)
    vand = np.vander(pos, order + 1, increasing=True)          # (window, order+1)
    pinv = np.linalg.pinv(vand)                                 # (order+1, window)
    centre = np.vander(np.array([float(half)]), order + 1, increasing=True)
    taps = (centre @ pinv)[0]                                   # symmetric smoothing taps
    edge_lo = np.vander(pos[:half], order + 1, increasing=True) @ pinv           # rows 0..half-1
    edge_hi = np.vander(pos[window - half:], order + 1, increasing=True) @ pinv  # rows n-half..n-1
    return taps, edge_lo, edge_hi


def savgol_axis(a, window, order, axis):
    taps, edge_lo, edge_hi = savgol_operators(window, order)
    a = np.moveaxis(np.asarray(a, dtype=np.float64), axis, 0)
    n = a.shape[0]
    half = window // 2
    if n < window:
        raise ValueError("window longer than the axis")
    out = np.empty_like(a)
    acc = np.zeros_like(a[half:n - half])
    for k in range(window):
        acc += taps[k] * a[k:n - window + 1 + k]
    out[half:n - half] = acc
    out[:half] = np.tensordot(edge_lo, a[:window], axes=(1, 0))
    out[n - half:] = np.tensordot(edge_hi, a[n - window:], axes=(1, 0))
    return np.moveaxis(out, 0, axis)


def savgol_separable(a, window, order):
    """axis 0 first, then axis 1 (deposit.py:195-196)."""
    return savgol_axis(savgol_axis(a, window, order, 0), window, order, 1)


# --------------------------------------------------------------------------------------------
# (f)4  true 2-D Savitzky-Golay filter (SGolay_filter.py:3-81; dead code in the reference's run loop,
#       deposit.py:187,194,227,232 — kept as an optional operator)
# --------------------------------------------------------------------------------------------
def sgolay2d_kernels(window: int, order: int):
    """The three convolution kernels the reference hands to ``fftconvolve`` (SGolay_filter.py:68-81):
    [0] smoothing, [1] minus the d/d(axis 0) fit coefficient ('col'), [2] minus the d/d(axis 1)
    coefficient ('row').  Least-squares fit of all monomials a^p b^q, p+q <= order, on the window."""
    if window % 2 == 0:
        raise ValueError("window_size must be odd")
    if window ** 2 < (order + 1) * (order + 2) / 2.0:
        raise ValueError("order is too high for the window size")
    half = window // 2
    a, b = np.meshgrid(np.arange(-half, half + 1, dtype=np.float64), np.arange(-half, half + 1, dtype=np.float64),
                       indexing="ij")                       # a: axis 0 offset (slow), b: axis 1 offset
    a, b = a.ravel(), b.ravel()
    # column order of the reference: total degree k ascending, then the axis-1 power n = 0..k  (SGolay_filter.py:24)
    cols = [(a ** (k - n)) * (b ** n) for k in range(order + 1) for n in range(k + 1)]
    pinv = np.linalg.pinv(np.stack(cols, axis=1))
    smooth = pinv[0].reshape(window, window)
    if order < 1:
        return np.stack([smooth, np.zeros_like(smooth), np.zeros_like(smooth)])
    return np.stack([smooth, -pinv[1].reshape(window, window), -pinv[2].reshape(window, window)])


def sgolay2d_pad(z, half: int):
    """Border extension of SGolay_filter.py:36-65: odd reflection with an absolute value, '-' on the
    low sides and '+' on the high sides; the top-right and bottom-left corners are built from the
    already padded right / bottom bands (the reference's own asymmetry)."""
    z = np.asarray(z, dtype=np.float64)
    H, W = z.shape
    h = half
    if h == 0:
        return z.copy()
    if H < 2 * h + 1 or W < 2 * h + 1:
        raise ValueError("array smaller than the window")
    Z = np.zeros((H + 2 * h, W + 2 * h))
    Z[h:H + h, h:W + h] = z
    top, bot, left, right = z[0, :], z[-1, :], z[:, :1], z[:, -1:]
    Z[:h, h:W + h] = top - np.abs(z[h:0:-1, :] - top)
    Z[H + h:, h:W + h] = bot + np.abs(z[H - 2:H - 2 - h:-1, :] - bot)
    Z[h:H + h, :h] = left - np.abs(z[:, h:0:-1] - left)
    Z[h:H + h, W + h:] = right + np.abs(z[:, W - 2:W - 2 - h:-1] - right)
    Z[:h, :h] = z[0, 0] - np.abs(z[h:0:-1, h:0:-1] - z[0, 0])
    Z[H + h:, W + h:] = z[-1, -1] + np.abs(z[H - 2:H - 2 - h:-1, W - 2:W - 2 - h:-1] - z[-1, -1])
    band = Z[h, W + h:]                                                     # right band of row 0
    Z[:h, W + h:] = band - np.abs(Z[2 * h:h:-1, W + h:] - band)
    band = Z[H + h:, h:h + 1]                                               # bottom band of column 0
    Z[H + h:, :h] = band - np.abs(Z[H + h:, 2 * h:h:-1] - band)
    return Z


def sgolay2d(z, window: int, order: int, derivative=None):
    """sgolay2d(z, window_size, order, derivative) of SGolay_filter.py:3 with the 'valid' convolution
    written as a direct sum (the reference uses an FFT convolution: same numbers to ~1e-16 of the
    largest term)."""
    ker = sgolay2d_kernels(window, order)
    Z = sgolay2d_pad(z, window // 2)
    H, W = np.asarray(z).shape

    def conv(k):
        out = np.zeros((H, W))
        for a in range(window):
            for b in range(window):
                out += k[a, b] * Z[window - 1 - a:window - 1 - a + H, window - 1 - b:window - 1 - b + W]
        return out

    if derivative is None:
        return conv(ker[0])
    if derivative == "col":
        return conv(ker[1])
    if derivative == "row":
        return conv(ker[2])
    if derivative == "both":
        return conv(ker[1]), conv(ker[2])
    raise ValueError("derivative must be None, 'col', 'row' or 'both'")


# --------------------------------------------------------------------------------------------
# A3  np.gradient with coordinate arrays (deposit.py:212-213; numpy/lib/_function_base_impl.py)
# --------------------------------------------------------------------------------------------
def gradient_axis(f, coords, axis):
    """Second-order interior differences, first-order one-sided edges.  numpy switches to the
    uniform formula only when *all* node differences are bit-identical; linspace nodes usually
    are not, so both branches are restated."""
    f = np.moveaxis(np.asarray(f, dtype=np.float64), axis, 0)
    h = np.diff(np.asarray(coords, dtype=np.float64))
    out = np.empty_like(f)
    shape = (-1,) + (1,) * (f.ndim - 1)
    if (h == h[0]).all():
        out[1:-1] = (f[2:] - f[:-2]) / (2.0 * h[0])
    else:
        h_lo = h[:-1]
        h_hi = h[1:]
        a = (-(h_hi) / (h_lo * (h_lo + h_hi))).reshape(shape)
        b = ((h_hi - h_lo) / (h_lo * h_hi)).reshape(shape)
        c = (h_lo / (h_hi * (h_lo + h_hi))).reshape(shape)
        out[1:-1] = a * f[:-2] + b * f[1:-1] + c * f[2:]
    out[0] = (f[1] - f[0]) / h[0]
    out[-1] = (f[-1] - f[-2]) / h[-1]
    return np.moveaxis(out, 0, axis)


def trapz_axis0_then_1(f, xg, zg):
    """np.trapz(np.trapz(f, xg, axis=0), zg) (deposit.py:201)."""
    dx = np.diff(xg)
    inner = (dx[:, None] * (f[1:] + f[:-1]) / 2.0).sum(axis=0)
    dz = np.diff(zg)
    return float((dz * (inner[1:] + inner[:-1]) / 2.0).sum())


# --------------------------------------------------------------------------------------------
# A4  DF_tracker.get_DF (deposit.py:145-245)
# --------------------------------------------------------------------------------------------
@dataclass
class DepositConfig:
    """deposit.py:132-143 (defaults and the reference's spelling ``velocity_threhold``)."""
    xbins: int = 100
    zbins: int = 100
    xlim: float = 5
    zlim: float = 5
    filter_order: int = 0
    filter_window: int = 0
    velocity_threhold: float = 5
    upper_limit: object = None


@dataclass
class DensityFunctions:
    x_grids: np.ndarray
    z_grids: np.ndarray
    density: np.ndarray
    vx: np.ndarray
    density_x: np.ndarray
    density_z: np.ndarray
    vx_x: np.ndarray
    t: float
    sigma_x: float
    sigma_z: float
    xmean: float
    zmean: float


def choose_deposit_grid(x, z, cfg: DepositConfig, sigma_x=None):
    """deposit.py:157-167: the YAML grid is honoured only for a strongly tilted bunch; note the
    slice test uses |z| (not |z - mean z|)."""
    sigma_x = np.std(x) if sigma_x is None else sigma_x
    sigma_z = np.std(z)
    central = np.abs(z) < 0.1 * sigma_z
    frac = sigma_x / np.std(x[central])
    if frac > 5:
        return cfg.xbins, cfg.zbins, cfg.filter_window
    return 100, 100, 5


def make_density_functions(x, z, px, t, cfg: DepositConfig) -> DensityFunctions:
    x = np.asarray(x, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    px = np.asarray(px, dtype=np.float64)
    sigma_x, sigma_z = np.std(x), np.std(z)
    xmean, zmean = np.mean(x), np.mean(z)
    xb, zb, window = choose_deposit_grid(x, z, cfg, sigma_x)
    x_lo, x_hi = xmean - cfg.xlim * sigma_x, xmean + cfg.xlim * sigma_x
    z_lo, z_hi = zmean - cfg.zlim * sigma_z, zmean + cfg.zlim * sigma_z
    xg = np.linspace(x_lo, x_hi, xb)
    zg = np.linspace(z_lo, z_hi, zb)
    count = cic_deposit_2d(x, z, np.ones(x.shape), xb, x_lo, x_hi, zb, z_lo, z_hi)
    vx = cic_deposit_2d(x, z, px, xb, x_lo, x_hi, zb, z_lo, z_hi)
    thr = np.max(count) / cfg.velocity_threhold                     # deposit.py:183 (raw counts)
    dense = count > thr
    vx[dense] /= count[dense]
    density = savgol_separable(count, window, cfg.filter_order)
    vx = savgol_separable(vx, window, cfg.filter_order)
    density = density / trapz_axis0_then_1(density, xg, zg)
    vx[density <= thr] = 0.0                                        # deposit.py:204 (same thr)
    density_x = gradient_axis(density, xg, 0)
    density_z = gradient_axis(density, zg, 1)
    vx_x = gradient_axis(vx, xg, 0)
    density_x = savgol_separable(density_x, window, cfg.filter_order)
    density_z = savgol_separable(density_z, window, cfg.filter_order)
    vx_x = savgol_separable(vx_x, window, cfg.filter_order)
    thr2 = np.max(density) / cfg.velocity_threhold * 8              # deposit.py:233
    vx_x[density < thr2] = np.mean(vx_x[density > thr2])
    return DensityFunctions(xg, zg, density, vx, density_x, density_z, vx_x, t,
                            float(sigma_x), float(sigma_z), float(xmean), float(zmean))


# --------------------------------------------------------------------------------------------
# A7  bilinear re-gridding == scipy RegularGridInterpolator(linear, bounds_error=False, fill)
#     (deposit.py:296-309; scipy/interpolate/_rgi.py:446-477,635-642)
# --------------------------------------------------------------------------------------------
def _rgi_axis(grid, q):
    idx = np.clip(np.searchsorted(grid, q, side="right") - 1, 0, grid.size - 2)
    frac = (q - grid[idx]) / (grid[idx + 1] - grid[idx])
    outside = (q < grid[0]) | (q > grid[-1])
    return idx, frac, outside


def regrid_bilinear(src, xg, zg, xq, zq, fill=0.0):
    """Tensor-product query (meshgrid(xq, zq, indexing='ij')) of the linear interpolant."""
    ix, fx, ox = _rgi_axis(np.asarray(xg), np.asarray(xq))
    iz, fz, oz = _rgi_axis(np.asarray(zg), np.asarray(zq))
    ix_, iz_ = ix[:, None], iz[None, :]
    fx_, fz_ = fx[:, None], fz[None, :]
    val = (src[ix_, iz_] * (1 - fx_) * (1 - fz_) + src[ix_, iz_ + 1] * (1 - fx_) * fz_
           + src[ix_ + 1, iz_] * fx_ * (1 - fz_) + src[ix_ + 1, iz_ + 1] * fx_ * fz_)
    val[ox[:, None] | oz[None, :]] = fill
    return val


def interp_bilinear_points(src, xg, zg, xq, zq, fill=0.0):
    """Point-wise query of the same interpolant (beams.py:111-112,118-119)."""
    ix, fx, ox = _rgi_axis(np.asarray(xg), np.asarray(xq))
    iz, fz, oz = _rgi_axis(np.asarray(zg), np.asarray(zq))
    val = (src[ix, iz] * (1 - fx) * (1 - fz) + src[ix, iz + 1] * (1 - fx) * fz
           + src[ix + 1, iz] * fx * (1 - fz) + src[ix + 1, iz + 1] * fx * fz)
    val[ox | oz] = fill
    return val


# --------------------------------------------------------------------------------------------
# A8  history log, sliding window and re-gridding policy (deposit.py:247-280,312-426)
# --------------------------------------------------------------------------------------------
FIELDS = ("density", "density_x", "density_z", "vx", "vx_x")


@dataclass
class HistoryStack:
    """What build_interpolant publishes (deposit.py:416-426): five (T,X,Z) stacks + metadata.
    Axis naming follows the consumer: 'x' is retarded time, 'y' transverse, 'z' longitudinal."""
    data: dict
    min_x: float
    min_y: float
    min_z: float
    delta_x: float
    delta_y: float
    delta_z: float

    @property
    def shape(self):
        return self.data["density"].shape


class HistoryOracle:
    def __init__(self, cfg: DepositConfig):
        self.cfg = cfg
        self.raw = deque()            # DensityFunctions per logged step
        self.start_time = 0.0
        self.end_time = 0.0
        self.sigma_x_interp = None
        self.sigma_z_interp = None
        self.x_grid_interp = None
        self.z_grid_interp = None
        self.time_interp = deque()
        self.slices = {k: deque() for k in FIELDS}
        self.interp_start = 0
        self.rebuilds = 0

    def append(self, df: DensityFunctions):            # deposit.py:247-256
        self.raw.append(df)
        self.end_time = df.t
        self.current = df

    def _resample(self, df):
        out = {}
        for name in FIELDS:
            src = getattr(df, name)
            fill = float(np.mean(src)) if name == "vx_x" else 0.0       # deposit.py:332,384
            out[name] = regrid_bilinear(src, df.x_grids, df.z_grids, self.x_grid_interp,
                                        self.z_grid_interp, fill)
        return out

    def push(self, formation_length, n_formation_length):
        """append_interpolant (deposit.py:312-390).  Returns True when the stack was rebuilt."""
        window_start = max(0, self.end_time - n_formation_length * formation_length)
        while self.start_time < window_start:                            # deposit.py:265-270
            self.raw.popleft()
            self.start_time = self.raw[0].t
        while self.interp_start < window_start:                          # deposit.py:273-280
            for name in FIELDS:
                self.slices[name].popleft()
            self.time_interp.popleft()
            self.interp_start = self.time_interp[0]
        df = self.current
        keep = (self.sigma_x_interp and self.sigma_z_interp
                and 2 > df.sigma_x / self.sigma_x_interp > 1 / 2
                and 2 > df.sigma_z / self.sigma_z_interp > 1 / 2)
        if keep:
            self.time_interp.append(df.t)
            new = self._resample(df)
            for name in FIELDS:
                self.slices[name].append(new[name])
            return False
        sx = [d.sigma_x for d in self.raw]
        sz = [d.sigma_z for d in self.raw]
        nx = int(500 * (max(sx) / min(sx)))
        nz = int(500 * (max(sz) / min(sz)))
        if isinstance(self.cfg.upper_limit, int):                        # deposit.py:359-361
            nx, nz = min(nx, self.cfg.upper_limit), min(nz, self.cfg.upper_limit)
        self.sigma_x_interp, self.sigma_z_interp = max(sx), max(sz)
        self.x_grid_interp = np.linspace(df.xmean - 5 * self.sigma_x_interp,
                                         df.xmean + 5 * self.sigma_x_interp, nx)
        self.z_grid_interp = np.linspace(df.zmean - 5 * self.sigma_z_interp,
                                         df.zmean + 5 * self.sigma_z_interp, nz)
        self.slices = {k: deque() for k in FIELDS}
        self.time_interp = deque(d.t for d in self.raw)
        for d in self.raw:
            new = self._resample(d)
            for name in FIELDS:
                self.slices[name].append(new[name])
        self.rebuilds += 1
        return True

    def stack(self) -> HistoryStack:                                     # deposit.py:395-426
        t0, t1 = self.time_interp[0], self.time_interp[-1]
        nt = len(self.time_interp)
        with np.errstate(divide="ignore", invalid="ignore"):
            dt = np.float64(t1 - t0) / np.float64(nt - 1)               # nt==1 -> nan, like numpy
        xg, zg = self.x_grid_interp, self.z_grid_interp
        return HistoryStack({k: np.array(self.slices[k]) for k in FIELDS},
                            t0, xg[0], zg[0], float(dt),
                            (xg[-1] - xg[0]) / (xg.shape[0] - 1), (zg[-1] - zg[0]) / (zg.shape[0] - 1))


# --------------------------------------------------------------------------------------------
# A5/A6  uniform-grid gathers with the reference's truncation edge rules
#        (interp3D.py:18-66, interp1D.py:13-36)
# --------------------------------------------------------------------------------------------
@njit(cache=False)
def _axis_cell(u, n):
    """(i0, i1, frac, ok): int() truncates toward zero, the last node clamps, everything else
    outside [0, n) is rejected; NaN fails every comparison and is rejected too."""
    if not (u > -1.0e18 and u < 1.0e18):
        return 0, 0, 0.0, False
    i0 = int(u)
    i1 = i0 if i0 == n - 1 else i0 + 1
    return i0, i1, u - i0, (i0 >= 0 and i1 < n)


@njit(cache=False)
def _gather5(tq, yq, zq, f0, f1, f2, f3, f4, m0, m1, m2, d0, d1, d2, out):
    nt, ny, nz = f0.shape
    for p in range(tq.shape[0]):
        a0, a1, ad, aok = _axis_cell((tq[p] - m0) / d0, nt)
        b0, b1, bd, bok = _axis_cell((yq[p] - m1) / d1, ny)
        c0, c1, cd, cok = _axis_cell((zq[p] - m2) / d2, nz)
        if not (aok and bok and cok):
            for k in range(5):
                out[k, p] = 0.0
            continue
        for k in range(5):
            f = f0 if k == 0 else (f1 if k == 1 else (f2 if k == 2 else (f3 if k == 3 else f4)))
            e00 = f[a0, b0, c0] * (1 - ad) + f[a1, b0, c0] * ad
            e01 = f[a0, b0, c1] * (1 - ad) + f[a1, b0, c1] * ad
            e10 = f[a0, b1, c0] * (1 - ad) + f[a1, b1, c0] * ad
            e11 = f[a0, b1, c1] * (1 - ad) + f[a1, b1, c1] * ad
            g0 = e00 * (1 - bd) + e10 * bd
            g1 = e01 * (1 - bd) + e11 * bd
            out[k, p] = g0 * (1 - cd) + g1 * cd


def gather_history(hist: HistoryStack, tq, yq, zq):
    """Five trilinear gathers at once; returns (5, n) in FIELDS order."""
    tq = np.ascontiguousarray(tq, dtype=np.float64)
    out = np.empty((5, tq.shape[0]))
    d = hist.data
    _gather5(tq, np.ascontiguousarray(yq, dtype=np.float64), np.ascontiguousarray(zq, dtype=np.float64),
             d["density"], d["density_x"], d["density_z"], d["vx"], d["vx_x"],
             float(hist.min_x), float(hist.min_y), float(hist.min_z),
             float(hist.delta_x), float(hist.delta_y), float(hist.delta_z), out)
    return out


def interp3d(xq, yq, zq, data, min_x, min_y, min_z, delta_x, delta_y, delta_z):
    """Single-field entry with the reference's signature (interp3D.py:19)."""
    h = HistoryStack({k: data for k in FIELDS}, min_x, min_y, min_z, delta_x, delta_y, delta_z)
    return gather_history(h, np.atleast_1d(xq), np.atleast_1d(yq), np.atleast_1d(zq))[0]


@njit(cache=False)
def _gather1(q, table, m, d, out):
    n = table.shape[0]
    for p in range(q.shape[0]):
        i0, i1, fr, ok = _axis_cell((q[p] - m) / d, n)
        out[p] = table[i0] * (1 - fr) + table[i1] * fr if ok else 0.0


def interp1d(q, table, min_x, delta_x):
    q = np.ascontiguousarray(np.atleast_1d(q), dtype=np.float64)
    out = np.empty(q.shape[0])
    _gather1(q, np.ascontiguousarray(table, dtype=np.float64), float(min_x), float(delta_x), out)
    return out


# --------------------------------------------------------------------------------------------
# A15  lattice tables consumed by the integrand (lattice.py:4-110,136-143)
# --------------------------------------------------------------------------------------------
@dataclass
class LatticeTables:
    coords: np.ndarray      # (Ns, 2) reference orbit
    n_vec: np.ndarray       # (Ns, 2) normal
    tau_vec: np.ndarray     # (Ns, 2) tangent
    min_s: float
    delta_s: float
    rho: np.ndarray         # (Ne,) curvature per element
    distance: np.ndarray    # (Ne,) cumulative element end positions

    def curvature(self, sp):
        """Piecewise-constant curvature, zero past the last element (CSR.py:651-656)."""
        idx = np.searchsorted(self.distance, sp, side="right")
        rho_ext = np.concatenate([self.rho, [0.0]])
        return rho_ext[idx]


def reference_orbit(elements, n_sample=2000):
    """Reference orbit tables from a list of (type, L, angle) tuples (lattice.py:4-110).
    Returns LatticeTables.  Element bookkeeping (``> distance`` test, per-sample rotation about
    the instantaneous centre) follows the reference so the tables agree to rounding."""
    lengths = np.array([e[1] for e in elements], dtype=np.float64)
    distance = np.zeros(len(elements))
    rho = np.zeros(len(elements))
    for k, (kind, length, angle) in enumerate(elements):
        distance[k] = length if k == 0 else length + distance[k - 1]
        if kind == "dipole":
            rho[k] = angle / length
    s = np.linspace(0, distance[-1], n_sample)
    coords = np.zeros((n_sample, 2))
    tau = np.zeros((n_sample, 2))
    nrm = np.zeros((n_sample, 2))
    theta = 0
    tau[0] = (np.cos(theta), np.sin(theta))
    nrm[0] = (np.sin(theta), -1 * np.cos(theta))
    cur = 0
    for k in range(1, n_sample):
        if s[k] > distance[cur]:
            cur += 1
        ds = s[k] - s[k - 1]
        kind, length, angle = elements[cur]
        if kind == "dipole":
            phi = ds / length * angle
            rad = length / angle
            cx = coords[k - 1, 0] - rad * np.sin(theta)
            cy = coords[k - 1, 1] + rad * np.cos(theta)
            coords[k, 0] = cx + rad * np.sin(phi + theta)
            coords[k, 1] = cy - rad * np.cos(phi + theta)
        else:
            phi = 0
            coords[k, 0] = coords[k - 1, 0] + ds * np.cos(theta)
            coords[k, 1] = coords[k - 1, 1] + ds * np.sin(theta)
        theta += phi
        tau[k] = (np.cos(theta), np.sin(theta))
        nrm[k] = (np.sin(theta), -1 * np.cos(theta))
    del lengths
    return LatticeTables(coords, nrm, tau, float(s[0]), float((s[-1] - s[0]) / (s.shape[0] - 1)), rho, distance)


# --------------------------------------------------------------------------------------------
# A9-A12  observation mesh, region set-up, integrand, nested trapezoid (CSR.py:361-782)
# --------------------------------------------------------------------------------------------
@dataclass
class WakeScalars:
    """Beam/run scalars read by get_CSR_wake (CSR.py:456-467,539; CSR.py:80)."""
    t: float                 # beam.position
    sigma_x: float
    sigma_z: float
    slope0: float            # polyfit(z, x, 1)[0]
    mean_x: float
    formation_window: float  # n_formation_length * formation_length
    csr_scaling: float       # 8.98755e3 * charge
    nx: int = 200            # integration_params.xbins
    nz: int = 200            # integration_params.zbins


def observation_mesh(x, z, slope, sigma_z, mean_z, xlim, zlim, xbins, zbins):
    """get_CSR_mesh (CSR.py:361-394): x-major flattening of (xrange_T x zrange), then the chirp
    line is added back."""
    x_t = x - np.polyval(slope, z)
    zr = np.linspace(mean_z - zlim * sigma_z, mean_z + zlim * sigma_z, zbins)
    xr = np.linspace(np.mean(x_t) - xlim * np.std(x_t), np.mean(x_t) + xlim * np.std(x_t), xbins)
    xm, zm = np.meshgrid(xr, zr, indexing="ij")
    zm = zm.ravel()
    return xm.ravel() + np.polyval(slope, zm), zm, xr, zr


def wake_regions(s, x, sc: WakeScalars):
    """Rectangles of the (x', s') quadrature as tuples (x_lo, x_hi, n_x, s_lo, s_hi, n_s), in the
    order the reference accumulates them (CSR.py:480-553,577-585).  Bounds may be reversed or
    degenerate; they are kept signed."""
    sx, sz, tan_t = sc.sigma_x, sc.sigma_z, sc.slope0
    x0 = (s - sc.t) * tan_t
    if abs(tan_t) <= 1:
        s2, s3, s4 = s - 500 * sz, s - 20 * sz, s + 5 * sz
        s1 = max(0, s2 - sc.formation_window)
        wide = (x0 - 20 * sx, x0 + 20 * sx, 2 * sc.nx)
        narrow = (x0 - 10 * sx, x0 + 10 * sx, sc.nx)
        return [wide + (s1, s2, sc.nz), narrow + (s2, s3, sc.nz), narrow + (s3, s4, sc.nz)]
    if tan_t > 0:
        tan_a = -2 * tan_t / (1 - tan_t ** 2)
        d = (10 * sx + sc.mean_x - x) / tan_a
        near = (x + 0.1 * sx, x + 10 * sx, sc.nx)
        core = (x - 3 * sx, x + 0.1 * sx, sc.nx)
    else:
        tan_a = 2 * tan_t / (1 - tan_t ** 2)
        d = -(sc.mean_x - x - 10 * sx) / tan_a
        near = (x - 10 * sx, x - 1 * sx, sc.nx)
        core = (x - 1 * sx, x + 3 * sx, sc.nx)
    s4 = s + 3 * sz
    s3 = max(0, s - d)
    s2 = s3 - 200 * sz
    s1 = max(0, s2 - sc.formation_window)
    band = (x0 - 5 * sx, x0 + 5 * sx, sc.nx)
    wide = (x0 - 20 * sx, x0 + 20 * sx, 2 * sc.nx)
    return [wide + (s1, s2, sc.nz), band + (s2, s3, sc.nz), near + (s3, s4, sc.nz), core + (s3, s4, sc.nz)]


def wake_integrand(s, x, sc: WakeScalars, lat: LatticeTables, hist: HistoryStack, xp, sp):
    """Longitudinal and transverse integrands on flat sample arrays (CSR.py:605-782)."""
    t = sc.t
    vx_obs = gather_history(hist, np.array([t]), np.array([x]), np.array([s - t]))[3, 0]

    def orbit(q):
        return (interp1d(q, lat.coords[:, 0], lat.min_s, lat.delta_s), interp1d(q, lat.coords[:, 1], lat.min_s, lat.delta_s),
                interp1d(q, lat.n_vec[:, 0], lat.min_s, lat.delta_s), interp1d(q, lat.n_vec[:, 1], lat.min_s, lat.delta_s),
                interp1d(q, lat.tau_vec[:, 0], lat.min_s, lat.delta_s), interp1d(q, lat.tau_vec[:, 1], lat.min_s, lat.delta_s))

    X0, Y0, nx0, ny0, tx0, ty0 = (v[0] for v in orbit(np.array([s])))
    X1, Y1, nx1, ny1, tx1, ty1 = orbit(sp)
    rx = X0 - X1 + x * nx0 - xp * nx1
    ry = Y0 - Y1 + x * ny0 - xp * ny1
    r = np.sqrt(rx ** 2 + ry ** 2)
    kappa = lat.curvature(sp)
    t_ret = t - r
    rho_r, rho_x_r, rho_z_r, vx_r, vxx_r = gather_history(hist, t_ret, xp, sp - t_ret)
    scale = 1 + xp * kappa
    vel_x, vel_y = tx0 + vx_obs * nx0, ty0 + vx_obs * ny0
    velr_x, velr_y = tx1 + vx_r * nx1, ty1 + vx_r * ny1
    grad_x = rho_x_r * nx1 + rho_z_r / scale * tx1
    grad_y = rho_x_r * ny1 + rho_z_r / scale * ty1
    dot = vel_x * velr_x + vel_y * velr_y
    num1 = scale * ((vel_x - dot * velr_x) * grad_x + (vel_y - dot * velr_y) * grad_y)
    num2 = -scale * dot * rho_r * vxx_r
    i_z = num1 / r + num2 / r
    q1 = rx * (nx0 - nx1) + ry * (ny0 - ny1)
    q2 = nx0 * tx1 + ny0 * ty1
    d_rho = -(velr_x * grad_x + velr_y * grad_y) - rho_r * vxx_r
    i_x = scale * q1 / (r * r * r) * rho_r + scale * q1 / (r * r) * d_rho + (-scale * q2 / r * d_rho)
    return i_z, i_x


def _trapz2(f, xn, sn):
    inner = (np.diff(xn)[:, None] * (f[1:] + f[:-1]) / 2.0).sum(axis=0)
    return (np.diff(sn) * (inner[1:] + inner[:-1]) / 2.0).sum()


def wake_point(s, x, sc: WakeScalars, lat: LatticeTables, hist: HistoryStack):
    """(dE_dct, x_kick) at one observation point (CSR.py:454-602): per region a nested trapezoid,
    inner over x' (axis 0), outer over s'; longitudinal carries a minus sign."""
    de = 0.0
    kick = 0.0
    terms = []
    for (xa, xb, n_x, sa, sb, n_s) in wake_regions(s, x, sc):
        xn = np.linspace(xa, xb, n_x)
        sn = np.linspace(sa, sb, n_s)
        xm, sm = np.meshgrid(xn, sn, indexing="ij")
        with np.errstate(divide="ignore", invalid="ignore"):
            i_z, i_x = wake_integrand(s, x, sc, lat, hist, xm.ravel(), sm.ravel())
        terms.append((-sc.csr_scaling * _trapz2(i_z.reshape(xm.shape), xn, sn),
                      sc.csr_scaling * _trapz2(i_x.reshape(xm.shape), xn, sn)))
    for a, b in terms:
        de, kick = de + a, kick + b
    return de, kick


def wake_mesh(xmesh, zmesh, sc: WakeScalars, lat: LatticeTables, hist: HistoryStack, first=0, count=None):
    """calculate_2D_CSR over a contiguous block of the flat mesh (CSR.py:404-415,434-445)."""
    count = len(xmesh) - first if count is None else count
    de = np.zeros(count)
    kick = np.zeros(count)
    for k in range(count):
        de[k], kick[k] = wake_point(sc.t + zmesh[first + k], xmesh[first + k], sc, lat, hist)
    return de, kick


def split_counts(n, parts):
    """The reference's MPI block split (CSR.py:121-125, test/test_mpi.py:14-17)."""
    ave, res = divmod(n, parts)
    count = [ave + 1 if p < res else ave for p in range(parts)]
    displ = [sum(count[:p]) for p in range(parts)]
    return count, displ


# --------------------------------------------------------------------------------------------
# A13/A14  kick application and beam scalars (beams.py:88-131,137-215)
# --------------------------------------------------------------------------------------------
def beam_scalars(x, z):
    slope = np.polyfit(z, x, deg=1)
    return dict(sigma_x=np.std(x), sigma_z=np.std(z), mean_x=np.mean(x), mean_z=np.mean(z), slope=slope)


def apply_kick(x, z, px, pz, de_dct, x_kick, xrange_t, zrange, step_size, init_energy, transverse_on=True):
    """beams.py:108-131: energy and angle kicks from bilinear samples of the wake grids at
    (x - polyval(slope, z), z); outside the mesh the kick is zero."""
    slope = np.polyfit(z, x, deg=1)
    x_t = x - np.polyval(slope, z)
    pz_new = pz + interp_bilinear_points(step_size * de_dct * 1e6 / init_energy, xrange_t, zrange, x_t, z, 0.0)
    px_new = px
    if transverse_on:
        px_new = px + interp_bilinear_points(step_size * x_kick * 1e6 / init_energy, xrange_t, zrange, x_t, z, 0.0)
    return px_new, pz_new


# --------------------------------------------------------------------------------------------
# (f)2  Twiss / dispersion statistics logged every step (twiss.py:2-71, CSR.py:837-859)
# --------------------------------------------------------------------------------------------
def twiss_from_coords(coords, p0c, mc2):
    """Per plane: the 3x3 sample covariance (np.cov, ddof = 1) of (q, p, delta), dispersion removed,
    then alpha, beta, gamma, emittance, eta, eta' and the normalised emittance (twiss.py:29-71)."""
    x, px, y, py, _z, pz = (np.asarray(c, dtype=np.float64) for c in coords)
    out = {}
    for plane, (q, p) in (("x", (x, px)), ("y", (y, py))):
        s = np.cov(np.stack([q, p, pz]))
        d2, qd, pd = s[2, 2], s[0, 2], s[1, 2]
        e_beta, e_gamma, e_alpha = s[0, 0] - qd * qd / d2, s[1, 1] - pd * pd / d2, -s[0, 1] + qd * pd / d2
        emit = math.sqrt(e_beta * e_gamma - e_alpha * e_alpha)
        vals = {"alpha": e_alpha / emit, "beta": e_beta / emit, "gamma": e_gamma / emit, "emit": emit,
                "eta": qd / d2, "etap": pd / d2, "norm_emit": emit * p0c / mc2}
        out.update({f"{k}_{plane}": v for k, v in vals.items()})
    return out


# --------------------------------------------------------------------------------------------
# Zero-density skipping of the CUDA wake kernel, as an executable specification (no reference
# counterpart: the reference evaluates every sample; every term of its integrands carries rho' or
# grad rho' (CSR.py:732-775), so samples whose eight voxels hold none add exactly 0).
# tests/ check (i) that property on the unmodified reference, (ii) that the rules below never drop a
# sample that can contribute, on many geometries; the GPU tests check the kernel bitwise.
# --------------------------------------------------------------------------------------------
def row_support(hist: HistoryStack):
    """(T, X, 2) int array: per slice and transverse row the hull [z_lo, z_hi] of the voxels with non-zero
    density, d(density)/dx or d(density)/dz; (INT32_MAX, -1) for an empty row (dfcsr_history_row_support)."""
    nzv = (hist.data["density"] != 0) | (hist.data["density_x"] != 0) | (hist.data["density_z"] != 0)
    T, X, Z = nzv.shape
    rows = nzv.any(axis=2)
    lo = np.where(rows, nzv.argmax(axis=2), np.iinfo(np.int32).max)
    hi = np.where(rows, Z - 1 - nzv[:, :, ::-1].argmax(axis=2), -1)
    return np.stack([lo, hi], axis=2).astype(np.int64)


def skip_plan(s, x, xp, s_lo, s_hi, n_s, sc: WakeScalars, lat: LatticeTables, hist: HistoryStack, support=None):
    """What the kernel does for ONE x' node of one rectangle (wake.cu, kSupport path).  Returns
    (swept, gathered): boolean arrays over the n_s s' nodes -- `swept` = nodes inside the 32-node blocks the
    coarse bracket keeps, `gathered` = swept nodes that pass the exact in-grid test and the per-slice hull test."""
    T, X, Z = hist.shape
    sup = row_support(hist) if support is None else support
    sn = np.linspace(s_lo, s_hi, n_s)
    gathered = np.zeros(n_s, bool)
    swept = np.zeros(n_s, bool)
    uy = (xp - hist.min_y) / hist.delta_y
    if not (uy > -1 and uy < X):
        return swept, gathered
    y0 = int(uy)
    y1 = y0 if y0 == X - 1 else y0 + 1

    def orbit(q, tab):
        return interp1d(q, tab, lat.min_s, lat.delta_s)
    obs = np.array([s])
    cx = orbit(obs, lat.coords[:, 0])[0] - orbit(sn, lat.coords[:, 0]) + x * orbit(obs, lat.n_vec[:, 0])[0]
    cy = orbit(obs, lat.coords[:, 1])[0] - orbit(sn, lat.coords[:, 1]) + x * orbit(obs, lat.n_vec[:, 1])[0]
    rx, ry = cx - xp * orbit(sn, lat.n_vec[:, 0]), cy - xp * orbit(sn, lat.n_vec[:, 1])
    t_ret = sc.t - np.sqrt(rx ** 2 + ry ** 2)
    ut = (t_ret - hist.min_x) / hist.delta_x
    uz = ((sn - t_ret) - hist.min_z) / hist.delta_z
    # per-slice pair hulls of the two rows: {lo - 1, hi} for cell t0 = hull(t0) U hull(t0 + 1)
    cell = np.empty((T, 2), np.int64)
    for t in range(T):
        t1 = t if t == T - 1 else t + 1
        lo = min(sup[t, y0, 0], sup[t, y1, 0], sup[t1, y0, 0], sup[t1, y1, 0])
        hi = max(sup[t, y0, 1], sup[t, y1, 1], sup[t1, y0, 1], sup[t1, y1, 1])
        cell[t] = (lo - 1 if lo != np.iinfo(np.int32).max else lo, hi)
    band_lo, band_hi = cell[:, 0].min(), cell[:, 1].max()
    # coarse pass: 32 samples spread over the rectangle, classes below / near / above the band
    m = (n_s + 31) >> 5
    jc = np.minimum(np.arange(32) * m, n_s - 1)
    kmax = float(np.max(np.abs(lat.rho))) if len(lat.rho) else 0.0
    step = abs(s_hi - s_lo) / (n_s - 1) if n_s > 1 else 0.0
    delta = 2.0 + 2.0 * m * step * abs(xp) * kmax / hist.delta_z
    cuz = uz[jc]
    cls = np.zeros(32, np.int64)
    with np.errstate(invalid="ignore"):
        cls[(band_lo > band_hi) | (cuz < band_lo - delta)] = 1
        cls[(cls == 0) & (cuz >= band_hi + 1.0 + delta)] = 2
    j_lo, j_hi = 0, n_s - 1
    lead = 0
    if cls[0] != 0:
        while lead < 32 and cls[lead] == cls[0]:
            lead += 1
    trail = 0
    if cls[31] != 0:
        while trail < 32 and cls[31 - trail] == cls[31]:
            trail += 1
    if lead == 32:
        return swept, gathered
    if lead > 0:
        j_lo = max(j_lo, min((lead - 1) * m, n_s - 1) & ~31)
    if trail > 0:
        j_hi = min(j_hi, min((32 - trail) * m, n_s - 1))
    for j0 in range(j_lo, j_hi + 1, 32):
        swept[j0:j0 + 32] = True
    with np.errstate(invalid="ignore"):
        inside = (ut > -1) & (ut < T) & (uz > -1) & (uz < Z)
    t0 = np.where(inside, ut, 0).astype(np.int64)
    z0 = np.where(inside, uz, 0).astype(np.int64)
    gathered = swept & inside & (z0 >= cell[t0, 0]) & (z0 <= cell[t0, 1])
    return swept, gathered
