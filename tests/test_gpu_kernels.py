"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle on the same
seeded inputs.  Tolerances: fp64 wakes/kicks 1e-10 relative (BASELINE.json north_star); integer
NGP counts and the re-gridding arithmetic bit-exact."""
import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from tests import scenario

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda:0")


def _up(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_beam_stats(dev, tilt):
    from pydfcsr_b200 import _lib, ops, synth
    b = synth.gaussian_bunch(300_001, seed=3, tilt=tilt)
    x, z, pz = b[0], b[4], b[5]
    st = ops.beam_stats(_up(x, dev), _up(z, dev), _up(pz, dev))
    slope = np.polyfit(z, x, 1)
    xt = x - np.polyval(slope, z)
    sl = np.abs(z) < 0.1 * np.std(z)
    exp = {_lib.S_MEAN_X: np.mean(x), _lib.S_MEAN_Z: np.mean(z), _lib.S_SIGMA_X: np.std(x), _lib.S_SIGMA_Z: np.std(z),
           _lib.S_SLOPE: slope[0], _lib.S_INTERCEPT: slope[1], _lib.S_SIGMA_XT: np.std(xt),
           _lib.S_SLICE_SIGMA_X: np.std(x[sl]), _lib.S_SLICE_COUNT: sl.sum(), _lib.S_SIGMA_PZ: np.std(pz),
           _lib.S_MEAN_PZ: np.mean(pz), _lib.S_N: x.size}
    for k, v in exp.items():
        assert abs(st[k] - v) <= 1e-11 * max(abs(v), 1e-30) + 1e-24, (k, st[k], v)
    assert abs(st[_lib.S_MEAN_XT]) < 1e-12 * np.std(x)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("shape", [(100, 100), (37, 53), (300, 200)])
def test_deposit_cic(dev, mode, shape):
    from pydfcsr_b200 import ops, synth
    b = synth.gaussian_bunch(200_000, seed=5)
    x, px, z = b[0], b[1], b[4]
    nx, nz = shape
    # a grid narrower than the bunch so that edge guards and far-away particles are exercised
    xs, xe = np.mean(x) - 2.5 * np.std(x), np.mean(x) + 3.0 * np.std(x)
    zs, ze = np.mean(z) - 5 * np.std(z), np.mean(z) + 5 * np.std(z)
    cnt, vxs = ops.deposit_cic(_up(x, dev), _up(z, dev), _up(px, dev), nx, xs, xe, nz, zs, ze, mode=mode)
    ref_c = O.cic_deposit_2d(x, z, np.ones_like(x), nx, xs, xe, nz, zs, ze)
    ref_v = O.cic_deposit_2d(x, z, px, nx, xs, xe, nz, zs, ze)
    assert _rel(cnt.cpu().numpy(), ref_c) < 1e-12
    assert _rel(vxs.cpu().numpy(), ref_v) < 1e-11   # signed weights: compare against max |.|


def test_deposit_cic_empty_and_single(dev):
    import torch
    from pydfcsr_b200 import ops
    e = torch.empty(0, dtype=torch.float64, device=dev)
    cnt, vxs = ops.deposit_cic(e, e, e, 8, 0.0, 1.0, 8, 0.0, 1.0)
    assert float(cnt.abs().sum()) == 0.0 and float(vxs.abs().sum()) == 0.0
    one = _up([0.5], dev)
    cnt, _ = ops.deposit_cic(one, one, one, 8, 0.0, 1.0, 8, 0.0, 1.0)
    ref = O.cic_deposit_2d(np.array([0.5]), np.array([0.5]), np.ones(1), 8, 0.0, 1.0, 8, 0.0, 1.0)
    assert np.array_equal(cnt.cpu().numpy(), ref)


@pytest.mark.parametrize("n", [3_000, 400_000])
def test_deposit_cic_fixed_point_is_reproducible_and_tight(dev, n):
    """The default (64-bit fixed-point) deposit: integer adds commute, so two runs agree bit for bit, and the
    quantisation stays at the 1e-14 level of the largest cell.  Zero and non-finite weights are handled."""
    import torch
    from pydfcsr_b200 import ops, synth
    b = synth.gaussian_bunch(n, seed=11, tilt=1.5)
    x, px, z = b[0], b[1], b[4]
    args = (300, np.mean(x) - 4 * np.std(x), np.mean(x) + 4 * np.std(x), 300, np.mean(z) - 4 * np.std(z), np.mean(z) + 4 * np.std(z))
    dx, dz, dp = _up(x, dev), _up(z, dev), _up(px, dev)
    c1, v1 = (t.clone() for t in ops.deposit_cic(dx, dz, dp, *args))
    c2, v2 = ops.deposit_cic(dx, dz, dp, *args)
    assert torch.equal(c1, c2) and torch.equal(v1, v2)
    ref_c = O.cic_deposit_2d(x, z, np.ones_like(x), *args)
    ref_v = O.cic_deposit_2d(x, z, px, *args)
    assert _rel(c1.cpu().numpy(), ref_c) < 2e-13 and _rel(v1.cpu().numpy(), ref_v) < 2e-13
    # a permutation of the particles gives the same bits (the fp64-atomic modes only agree to rounding)
    perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    c3, v3 = ops.deposit_cic(dx[perm].contiguous(), dz[perm].contiguous(), dp[perm].contiguous(), *args)
    if n >= 65536:      # same tile decomposition: the per-CTA rounding to the global scale depends on who owns what
        assert _rel(c3.cpu().numpy(), c1.cpu().numpy()) < 1e-13
    else:
        assert torch.equal(c3, c1) and torch.equal(v3, v1)
    c0, v0 = ops.deposit_cic(dx, dz, torch.zeros_like(dp), *args)
    assert torch.equal(c0, c1) and float(v0.abs().max()) == 0.0
    bad = dp.clone()
    bad[n // 2] = float("inf")
    cb, vb = ops.deposit_cic(dx, dz, bad, *args)
    assert torch.equal(cb, c1) and bool(torch.isnan(vb).all())


def test_deposit_ngp_bit_exact(dev):
    from pydfcsr_b200 import ops, synth
    b = synth.gaussian_bunch(500_000, seed=7)
    x, z = b[0], b[4]
    args = (64, np.mean(x) - 3 * np.std(x), np.mean(x) + 3 * np.std(x), 128, np.mean(z) - 3 * np.std(z), np.mean(z) + 3 * np.std(z))
    got = ops.deposit_ngp(_up(x, dev), _up(z, dev), *args).cpu().numpy()
    ref = O.ngp_deposit_2d(x, z, *args)
    assert got.dtype == np.int64 and np.array_equal(got, ref)
    assert got.sum() < x.size            # some particles fall outside the 3-sigma grid


@pytest.mark.parametrize("tilt,order,window", [(0.0, 1, 9), (2.5, 1, 9), (2.5, 2, 9), (2.5, 0, 5), (0.0, 0, 0)])
def test_make_df(dev, tilt, order, window):
    from pydfcsr_b200 import ops, synth
    from pydfcsr_b200._lib import Axis
    b = synth.gaussian_bunch(200_000, seed=11, tilt=tilt)
    x, px, z = b[0], b[1], b[4]
    cfg = O.DepositConfig(xbins=64, zbins=96, filter_order=order, filter_window=window, velocity_threhold=1000)
    df = O.make_density_functions(x, z, px, 0.0, cfg)
    nx, nz = df.density.shape
    win = window if (nx, nz) == (64, 96) else 5
    xs, xe, zs, ze = df.x_grids[0], df.x_grids[-1], df.z_grids[0], df.z_grids[-1]
    cnt = O.cic_deposit_2d(x, z, np.ones_like(x), nx, xs, xe, nz, zs, ze)
    vxs = O.cic_deposit_2d(x, z, px, nx, xs, xe, nz, zs, ze)
    fields, scal = ops.make_df(_up(cnt, dev), _up(vxs, dev), Axis.make(xs, xe, nx), Axis.make(zs, ze, nz), win, order, 1000)
    f = fields.cpu().numpy()
    for k, name in enumerate(O.FIELDS):
        assert _rel(f[k], getattr(df, name)) < 1e-11, name
    assert abs(float(scal[4]) - np.mean(df.vx_x)) <= 1e-11 * np.max(np.abs(df.vx_x))


def test_history_regrid_bit_exact(dev):
    from pydfcsr_b200 import ops
    from pydfcsr_b200._lib import Axis
    import torch
    rng = np.random.default_rng(0)
    src = rng.normal(size=(5, 41, 57))
    sx, sz = np.linspace(-1.3e-4, 2.1e-4, 41), np.linspace(-9e-4, 1.1e-3, 57)
    # destination wider than the source on one side, inside on the other, and hitting source nodes exactly
    dx, dz = np.linspace(-2.0e-4, 2.1e-4, 333), np.linspace(-9e-4, 0.9e-3, 250)
    out = torch.empty((333, 250, 6), dtype=torch.float64, device=dev)
    ops.history_regrid(_up(src, dev), Axis.make(sx[0], sx[-1], 41), Axis.make(sz[0], sz[-1], 57),
                       Axis.make(dx[0], dx[-1], 333), Axis.make(dz[0], dz[-1], 250), 0.125, out)
    got = out.cpu().numpy()
    for k in range(5):
        ref = O.regrid_bilinear(src[k], sx, sz, dx, dz, 0.125 if k == 4 else 0.0)
        assert np.array_equal(got[..., k], ref), k
    assert not got[..., 5].any()
    back = ops.history_unpack(out, 333, 250).cpu().numpy()
    assert np.array_equal(back, np.moveaxis(got[..., :5], -1, 0))


def _device_problem(sc, dev, nx, nz):
    from pydfcsr_b200 import ops
    st, lat = sc["stack"], sc["lattice"]
    hist = ops.DeviceHistory.from_stacks([st.data[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z,
                                         st.delta_x, st.delta_y, st.delta_z, dev, cap=st.shape[0] + 3, head=2)
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, dev)
    wp = ops.wake_params(nx=nx, nz=nz, **sc["wake_scalars"])
    osc = O.WakeScalars(nx=nx, nz=nz, **sc["wake_scalars"])
    return hist, dlat, wp, osc


@pytest.mark.parametrize("tilt", [0.0, 2.5, -2.5])
def test_wake_mesh_matches_oracle(dev, tilt):
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    nx = nz = 50
    hist, dlat, wp, osc = _device_problem(sc, dev, nx, nz)
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 5, 7)
    import torch
    cnt = torch.zeros(3, dtype=torch.int64, device=dev)
    de, kick = ops.wake_mesh(hist, dlat, wp, _up(xm, dev), _up(zm, dev), counters=cnt)
    ref_de, ref_kick = O.wake_mesh(xm, zm, osc, sc["lattice"], sc["stack"])
    assert _rel(de.cpu().numpy(), ref_de) < TOL
    assert _rel(kick.cpu().numpy(), ref_kick) < TOL
    big = np.abs(ref_de) > 1e-3 * np.max(np.abs(ref_de))
    assert np.max(np.abs(de.cpu().numpy()[big] / ref_de[big] - 1)) < 1e-9
    n_in, n_all, n_gat = (int(v) for v in cnt.cpu())
    assert n_all == xm.size * (4 if abs(tilt) <= 1 else 5) * nx * nz and 0 < n_gat <= n_in < n_all
    # block split (CSR.py:121-125): any contiguous block gives the same numbers
    de2, _ = ops.wake_mesh(hist, dlat, wp, _up(xm, dev), _up(zm, dev), first=11, count=9)
    assert np.array_equal(de2.cpu().numpy(), de.cpu().numpy()[11:20])


@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_wake_point_debug_integrands(dev, tilt):
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    nx = nz = 24
    hist, dlat, wp, osc = _device_problem(sc, dev, nx, nz)
    s = sc["pos"] + 0.3 * sc["wake_scalars"]["sigma_z"]
    x = 0.4 * sc["wake_scalars"]["sigma_x"]
    got = ops.wake_point_debug(hist, dlat, wp, s, x)
    regions = O.wake_regions(s, x, osc)
    assert len(got) == len(regions)
    for g, (xa, xb, n_x, sa, sb, n_s) in zip(got, regions):
        xn, sn = np.linspace(xa, xb, n_x), np.linspace(sa, sb, n_s)
        assert np.array_equal(g["xp"], xn) and np.array_equal(g["sp"], sn)
        xm, sm = np.meshgrid(xn, sn, indexing="ij")
        iz, ix = O.wake_integrand(s, x, osc, sc["lattice"], sc["stack"], xm.ravel(), sm.ravel())
        assert _rel(g["integrand_z"].ravel(), iz) < 1e-11
        assert _rel(g["integrand_x"].ravel(), ix) < 1e-11


def test_apply_kick(dev):
    from pydfcsr_b200 import ops, synth
    from pydfcsr_b200._lib import Axis
    b = synth.gaussian_bunch(200_000, seed=13, tilt=0.7)
    x, px, z, pz = b[0], b[1], b[4], b[5]
    rng = np.random.default_rng(2)
    slope = np.polyfit(z, x, 1)
    xt = x - np.polyval(slope, z)
    xr = np.linspace(np.mean(xt) - 3 * np.std(xt), np.mean(xt) + 3 * np.std(xt), 10)
    zr = np.linspace(np.mean(z) - 3 * np.std(z), np.mean(z) + 3 * np.std(z), 30)
    de, kick = rng.normal(size=(10, 30)), rng.normal(size=(10, 30))
    ref_px, ref_pz = O.apply_kick(x, z, px, pz, de, kick, xr, zr, 0.1, 5e9, True)
    dpx, dpz = _up(px, dev), _up(pz, dev)
    ops.apply_kick(_up(x, dev), _up(z, dev), dpx, dpz, slope[0], slope[1], _up(de, dev), _up(kick, dev),
                   Axis.make(xr[0], xr[-1], 10), Axis.make(zr[0], zr[-1], 30), 0.1, 5e9, True)
    assert _rel(dpz.cpu().numpy() - pz, ref_pz - pz) < 1e-10
    assert _rel(dpx.cpu().numpy() - px, ref_px - px) < 1e-10
    assert np.count_nonzero(ref_pz == pz) > 0     # particles outside the 3-sigma mesh get no kick
    assert np.array_equal((dpz.cpu().numpy() == pz), (ref_pz == pz))


@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_wake_fp32_storage_mode(dev, tilt):
    """Optional mixed-precision mode (BASELINE.json north_star: 'an optional fp32 mode within 1e-4'):
    the history is stored and blended in fp32, geometry / indices / algebra / quadrature stay fp64.
    Gate 1e-4 relative to the mesh maximum; the measured error is ~1e-7."""
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    st, lat = sc["stack"], sc["lattice"]
    nx = nz = 50
    hist32 = ops.DeviceHistory.from_stacks([st.data[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z,
                                           st.delta_x, st.delta_y, st.delta_z, dev, cap=8, head=5, precision="fp32")
    assert hist32.ring.dtype == __import__("torch").float32 and hist32.ring.shape[-1] == 8
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, dev)
    wp = ops.wake_params(nx=nx, nz=nz, **sc["wake_scalars"])
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 5, 7)
    de, kick = ops.wake_mesh(hist32, dlat, wp, _up(xm, dev), _up(zm, dev))
    ref_de, ref_kick = O.wake_mesh(xm, zm, O.WakeScalars(nx=nx, nz=nz, **sc["wake_scalars"]), lat, st)
    e1, e2 = _rel(de.cpu().numpy(), ref_de), _rel(kick.cpu().numpy(), ref_kick)
    assert e1 < 1e-4 and e2 < 1e-4, (e1, e2)
    assert e1 < 5e-6 and e2 < 5e-6, (e1, e2)          # what the mode actually delivers
    # pack/unpack round trip of fp32 voxels equals a float32 cast of the fields
    back = ops.history_unpack(hist32.ring[5], st.shape[1], st.shape[2]).cpu().numpy()
    assert np.array_equal(back[0], st.data["density"][0].astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("shape,window,order", [((100, 100), 5, 2), ((300, 300), 9, 1), ((64, 512), 9, 2), ((33, 65), 25, 2),
                                                ((97, 31), 7, 3), ((3, 3), 3, 1), ((40, 40), 1, 0)])
def test_sgolay2d(dev, shape, window, order):
    """2-D Savitzky-Golay operator (SGolay_filter.py:3-81) against the oracle: every border region of the
    extension, tiles that are not multiples of 32, the largest window."""
    from pydfcsr_b200 import ops
    rng = np.random.default_rng(window + shape[0])
    z = rng.normal(size=shape) + np.linspace(0, 3, shape[1])[None, :] ** 2 - np.linspace(-1, 2, shape[0])[:, None]
    dz = _up(z, dev)
    assert _rel(ops.sgolay2d(dz, window, order).cpu().numpy(), O.sgolay2d(z, window, order)) < 1e-13
    if order >= 1:
        col, row = ops.sgolay2d(dz, window, order, "both")
        ref = O.sgolay2d(z, window, order, "both")
        assert _rel(col.cpu().numpy(), ref[0]) < 1e-13 and _rel(row.cpu().numpy(), ref[1]) < 1e-13
        # a plane a*i + b*j is reproduced exactly by the fit: derivative stencils return its slopes
        i, j = np.meshgrid(np.arange(shape[0], dtype=float), np.arange(shape[1], dtype=float), indexing="ij")
        col, row = ops.sgolay2d(_up(0.25 * i - 1.5 * j, dev), window, order, "both")
        h = window // 2
        inner = (slice(h, shape[0] - h), slice(h, shape[1] - h))
        if shape[0] > 2 * h:
            assert np.allclose(col.cpu().numpy()[inner], 0.25, atol=1e-11) and np.allclose(row.cpu().numpy()[inner], -1.5, atol=1e-11)


def test_sgolay2d_errors(dev):
    from pydfcsr_b200 import _lib, ops
    z = _up(np.zeros((8, 8)), dev)
    with pytest.raises(_lib.DfcsrError):
        ops.sgolay2d(z, 9, 1)                   # array smaller than the window
    with pytest.raises(ValueError):
        ops.sgolay2d(z, 4, 1)
    with pytest.raises(ValueError):
        ops.sgolay2d(z, 5, 1, "diag")


def test_beam_cov_and_twiss(dev):
    """6x6 covariance in one device pass (np.cov normalisation) and the Twiss statistics built on it
    (twiss.py:2-71), for a bunch far from the origin (the shifted sums must not lose the small variances)."""
    import torch
    from pydfcsr_b200 import Beam, ops, synth
    b = synth.gaussian_bunch(400_003, seed=13, tilt=1.2)
    b[0] += 0.02 * b[5] + 3.0e-3             # dispersion and a 50-sigma offset
    b[4] += 1.0e-2
    coords = [_up(c, dev) for c in b]
    mean, cov = ops.beam_cov(coords)
    ref_cov = np.cov(b)
    scale = np.sqrt(np.outer(np.diag(ref_cov), np.diag(ref_cov)))
    assert np.max(np.abs(cov - ref_cov) / scale) < 1e-10
    assert np.max(np.abs(mean - b.mean(axis=1)) / np.sqrt(np.diag(ref_cov))) < 1e-10
    m2, c2 = ops.beam_cov(coords)
    assert np.array_equal(cov, c2) and np.array_equal(mean, m2)            # deterministic reduction
    beam = Beam({"style": "array", "coords": torch.stack(coords), "charge": 1e-9, "energy": 5.0e9}, device=dev)
    want = O.twiss_from_coords(b, 5.0e9, 0.51099895e6)
    got = beam.twiss
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-8 * abs(v), (k, got[k], v)


def test_track_linear_kernel_matches_host_maps(dev):
    """The device transfer-map kernel against the host expressions of pydfcsr_b200/tracking.py for every
    element type of the stand-in tracker (drift, bend with edges, both quadrupole signs)."""
    from pydfcsr_b200 import synth, tracking
    b = synth.gaussian_bunch(50_001, seed=17, tilt=0.3)
    elements = [tracking.Drift(0.37), tracking.SBend(L=0.5002, G=0.0483 / 0.5002, E1=0.0, E2=0.0483),
                tracking.SBend(L=0.1, G=-0.0483 / 0.5002, E1=-0.0483, E2=0.0, FRINGE_AT="entrance_end"),
                tracking.Quadrupole(L=0.2, K1=1.7), tracking.Quadrupole(L=0.2, K1=-0.9), tracking.Sextupole(L=0.1, K2=3.0)]
    for el in elements:
        want = tracking.track_linear(tuple(b), el)
        dev_coords = tuple(_up(c, dev) for c in b)
        got = tracking.track_linear(dev_coords, el)
        assert all(g.data_ptr() == d.data_ptr() for g, d in zip(got, dev_coords))          # in place
        for k in range(6):
            scale = max(np.max(np.abs(want[k])), 1e-300)
            assert np.max(np.abs(got[k].cpu().numpy() - want[k])) <= 1e-14 * scale, (type(el).__name__, k)
        m = tracking.linear_matrix(el)
        assert abs(np.linalg.det(m) - 1.0) < 1e-12                                           # symplectic maps


# K4 with every zero-density skipping policy (dfcsr_wake_params.skip_mode): each must meet the same gate.
@pytest.mark.parametrize("skip", ["auto", "on", "off"])
@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_wake_skip_modes_match_oracle(dev, skip, tilt):
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    nx, nz = 50, 45           # odd node counts: the last 32-node block of a rectangle is partly empty
    hist, dlat, wp, osc = _device_problem(sc, dev, nx, nz)
    wp = ops.wake_params(nx=nx, nz=nz, skip=skip, **sc["wake_scalars"])
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 5, 7)
    import torch
    cnt = torch.zeros(3, dtype=torch.int64, device=dev)
    de, kick = ops.wake_mesh(hist, dlat, wp, _up(xm, dev), _up(zm, dev), counters=cnt)
    ref_de, ref_kick = O.wake_mesh(xm, zm, osc, sc["lattice"], sc["stack"])
    assert _rel(de.cpu().numpy(), ref_de) < TOL
    assert _rel(kick.cpu().numpy(), ref_kick) < TOL
    n_in, n_all, n_gat = (int(v) for v in cnt.cpu())
    assert n_all == xm.size * (4 if abs(tilt) <= 1 else 5) * nx * nz and 0 < n_gat <= n_in < n_all
    uses = ops.wake_uses_skipping(hist, wp)
    if skip != "auto" or abs(tilt) > 1:
        assert uses == (skip != "off")          # AUTO: the chirp-band branch (|slope| > 1) always counts as sparse
    assert n_gat == n_in if not uses else n_gat <= n_in
    if uses and tilt != 0.0:
        assert n_gat < n_in                     # the tilted bunch fills a band of its grid only
    # run-to-run bitwise reproducible
    de2, kick2 = ops.wake_mesh(hist, dlat, wp, _up(xm, dev), _up(zm, dev))
    assert torch.equal(de, de2) and torch.equal(kick, kick2)


def test_skipping_policy_switch_is_bitwise_neutral(dev):
    """The AUTO policy turns skipping on when the history grid is more than 1.5x the +-5 sigma box of the bunch
    (wake.cu: wants_skipping).  Straddle that boundary by scaling sigma_x in the wake scalars: the policy must flip,
    and on each side the wakes must be bitwise those of the other setting forced (skipping never changes a bit)."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=0.0)
    st = sc["stack"]
    hist, dlat, _wp, _ = _device_problem(sc, dev, 40, 40)
    base = dict(sc["wake_scalars"])
    area = (st.shape[1] * st.delta_y) * (st.shape[2] * st.delta_z)
    sx_edge = area / (1.5 * 100.0 * base["sigma_z"])          # grid_area == 1.5 * (10 sx)(10 sz)
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 4, 5)
    seen = []
    for f in (0.98, 1.02):
        scal = dict(base, sigma_x=sx_edge * f)
        res = {}
        for skip in ("auto", "on", "off"):
            wp = ops.wake_params(nx=40, nz=40, skip=skip, **scal)
            res[skip] = (ops.wake_uses_skipping(hist, wp),) + tuple(t.clone() for t in ops.wake_mesh(hist, dlat, wp, _up(xm, dev), _up(zm, dev)))
        assert res["on"][0] and not res["off"][0]
        seen.append(res["auto"][0])
        for k in (1, 2):
            assert torch.equal(res["auto"][k], res["on"][k]) and torch.equal(res["auto"][k], res["off"][k])
        assert float(res["auto"][1].abs().max()) > 0
    assert seen == [True, False]          # smaller bunch -> sparse grid -> skipping; larger bunch -> not


def test_fused_exchange_kernel_on_one_gpu(dev):
    """dfcsr_wake_grid_peers (K4 with the all-gather fused in) with the 'peers' being three grids of THIS GPU: three
    launches, one per block of the reference's split rule, must leave in every grid exactly the bits of the serial
    launch.  Covers the N > 1 kernel path on a single-GPU box (the cross-process mapping itself is covered by
    tests/nccl_worker.py and by the parity record of every multi-GPU bench line)."""
    import ctypes as C
    import torch
    from pydfcsr_b200 import ops
    from pydfcsr_b200._lib import Axis
    from pydfcsr_b200.distributed import split_counts
    sc = scenario.chicane_entry(tilt=0.0)
    hist, dlat, wp, _ = _device_problem(sc, dev, 40, 40)
    s = sc["scalars"]
    x, z = sc["coords"][0], sc["coords"][4]
    _, _, xr, zr = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 5, 7)
    xa, za = Axis.make(xr[0], xr[-1], 5), Axis.make(zr[0], zr[-1], 7)
    slope, icpt = float(s["slope"][0]), float(s["slope"][1])
    de, kick = ops.wake_grid(hist, dlat, wp, xa, za, slope, icpt)
    n, world = 35, 3
    grids = [torch.full((2, n), float("nan"), dtype=torch.float64, device=dev) for _ in range(world)]
    ptrs = (C.c_uint64 * world)(*[g.data_ptr() for g in grids])
    count, displ = split_counts(n, world)
    for r in range(world):
        ops.wake_grid_peers(hist, dlat, wp, xa, za, slope, icpt, first=displ[r], count=count[r], peer_ptrs=ptrs)
    for g in grids:
        assert torch.equal(g[0], de) and torch.equal(g[1], kick)
    # the same points dealt out round-robin (mesh point k to rank k mod 3), the default of CSR2D's fused exchange
    for g in grids:
        g.fill_(float("nan"))
    for r in range(world):
        ops.wake_grid_peers(hist, dlat, wp, xa, za, slope, icpt, first=r, count=(n - r + world - 1) // world, stride=world,
                            peer_ptrs=ptrs)
    for g in grids:
        assert torch.equal(g[0], de) and torch.equal(g[1], kick)


def test_fused_sqrt_is_bitwise_the_library_sqrt(dev):
    """t_ret = t - r must use the correctly rounded square root: the fused sqrt/rsqrt of the
    trimmed K4 kernels is compared bit for bit with sqrt.rn.f64 / rsqrt on 6e8 random arguments."""
    from pydfcsr_b200 import ops
    for seed, (lo, hi) in enumerate([(-60.0, 8.0), (-2.0, 2.0), (-1000.0, 1000.0)]):
        bad_r, bad_y = ops.selftest_sqrt(200_000_000, seed=seed + 1, lo_exp=lo, hi_exp=hi)
        assert bad_r == 0 and bad_y == 0, (lo, hi, bad_r, bad_y)


@pytest.mark.parametrize("tilt", [0.0, 2.5, -2.5])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_zero_density_skipping_is_exact(dev, tilt, precision):
    """dfcsr_history.d_row_support: samples whose eight voxels carry no density and no density gradient add exactly
    0 (every integrand term has a factor rho' or grad rho', CSR.py:732-775), so K4 skips them without loading the
    history.  The result must be BITWISE the one computed without the support table, and on a tilted beam most of
    the in-grid samples must actually be skipped."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    st, lat = sc["stack"], sc["lattice"]
    nx = nz = 50
    args = ([st.data[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z, st.delta_x, st.delta_y, st.delta_z, dev)
    h_on = ops.DeviceHistory.from_stacks(*args, cap=st.shape[0] + 3, head=2, precision=precision)
    h_off = ops.DeviceHistory.from_stacks(*args, cap=st.shape[0] + 3, head=2, precision=precision, row_support=False)
    assert h_on.support is not None and h_off.support is None
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, dev)
    wp = ops.wake_params(nx=nx, nz=nz, skip="on", **sc["wake_scalars"])   # on for every slope (AUTO: only sparse grids)
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 5, 7)
    c_on = torch.zeros(3, dtype=torch.int64, device=dev)
    c_off = torch.zeros(3, dtype=torch.int64, device=dev)
    de1, k1 = ops.wake_mesh(h_on, dlat, wp, _up(xm, dev), _up(zm, dev), counters=c_on)
    de0, k0 = ops.wake_mesh(h_off, dlat, wp, _up(xm, dev), _up(zm, dev), counters=c_off)
    assert torch.equal(de1, de0) and torch.equal(k1, k0)
    on, off = [int(v) for v in c_on.cpu()], [int(v) for v in c_off.cpu()]
    assert on[0] <= off[0] and on[1] == off[1] and off[2] == off[0] and 0 < on[2] <= on[0]
    if tilt != 0.0:
        assert on[2] < 0.9 * off[0], (on, off)   # the tilted bunch fills a band of the history grid (72 % here)
    # the support table is what the oracle's stack says: hull of the voxels with rho, rho_x or rho_z != 0
    cap, head = h_on.ring.shape[0], 2
    sup = h_on.support.cpu().numpy()
    fld = [st.data[k] for k in O.FIELDS[:3]]
    if precision == "fp32":
        fld = [f.astype(np.float32) for f in fld]
    nzm = (fld[0] != 0) | (fld[1] != 0) | (fld[2] != 0)
    for k in range(st.shape[0]):
        rows = nzm[k].any(axis=1)
        lo = np.where(rows, nzm[k].argmax(axis=1), np.iinfo(np.int32).max)
        hi = np.where(rows, nzm.shape[2] - 1 - nzm[k][:, ::-1].argmax(axis=1), -1)
        assert np.array_equal(sup[(head + k) % cap, :, 0], lo) and np.array_equal(sup[(head + k) % cap, :, 1], hi)


def test_zero_density_skipping_empty_and_full_rows(dev):
    """Degenerate supports: an all-zero history gives exactly zero wakes without gathering anything; a history
    without a single zero voxel skips nothing."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=0.0)
    st, lat = sc["stack"], sc["lattice"]
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, dev)
    wp = ops.wake_params(nx=20, nz=33, skip="on", **sc["wake_scalars"])
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 3, 4)
    for fill in (0.0, 1.0):
        stacks = [np.full(st.shape, fill) for _ in O.FIELDS]
        h = ops.DeviceHistory.from_stacks(stacks, st.min_x, st.min_y, st.min_z, st.delta_x, st.delta_y, st.delta_z, dev)
        cnt = torch.zeros(3, dtype=torch.int64, device=dev)
        de, kick = ops.wake_mesh(h, dlat, wp, _up(xm, dev), _up(zm, dev), counters=cnt)
        n_in, _, n_gat = (int(v) for v in cnt.cpu())
        if fill == 0.0:
            assert n_gat == 0 and n_in == 0 and not de.any() and not kick.any()      # every x' node is dropped up front
        else:
            assert n_gat == n_in > 0 and bool(de.abs().max() > 0)


# ---- particle shards: statistics, covariance and deposit must not depend on how the particles are distributed ----
def _emulated_shards(n, world):
    """(first_block, n_blocks, lo, hi) of every rank of a `world`-rank job, as distributed.ParticleShards computes them."""
    from pydfcsr_b200 import _lib
    from pydfcsr_b200.distributed import shard_blocks
    chunk = int(_lib.lib.dfcsr_stat_chunk(n))
    out = []
    for r in range(world):
        f, c = shard_blocks(r, world)
        out.append((f, c, min(f * chunk, n), min((f + c) * chunk, n)))
    assert out[0][2] == 0 and out[-1][3] == n and all(a[3] == b[2] for a, b in zip(out, out[1:]))
    return out


@pytest.mark.parametrize("n", [1_000_003, 4097])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_statistics_and_covariance_are_independent_of_the_sharding(dev, n, world):
    """dfcsr_beam_stats / dfcsr_beam_cov on one GPU against the same bunch cut into `world` shards whose chunk totals
    meet in one table (what N ranks do over NVLink peer memory): every one of the 14 statistics and 27 covariance
    entries must be BITWISE equal."""
    import ctypes as C
    import torch
    from pydfcsr_b200 import _lib, ops, synth
    from pydfcsr_b200.ops import _ptr
    b = synth.gaussian_bunch(n, seed=11, tilt=0.7)
    b[0] += 3.0e-3
    coords = [_up(b[k], dev) for k in range(6)]
    x, px, z, pz = coords[0], coords[1], coords[4], coords[5]
    centre = [float(b[0][0]), float(b[4][0]), float(b[5][0])]
    centre6 = [float(b[k][0]) for k in range(6)]
    ref = ops.beam_stats(x, z, pz, px, centre=centre)
    rmean, rcov = ops.beam_cov(coords, centre=centre6)
    assert ref[_lib.S_ABSMAX_PX] == float(np.max(np.abs(b[1]))) and ref[_lib.S_N] == n
    table = torch.zeros(1024 * 8, dtype=torch.float64, device=dev)
    ctab = torch.zeros(1024 * 27, dtype=torch.float64, device=dev)
    d_stats = torch.zeros(16, dtype=torch.float64, device=dev)
    d_cov = torch.zeros(27, dtype=torch.float64, device=dev)
    ctr = (C.c_double * 3)(*centre)
    ctr6 = (C.c_double * 6)(*centre6)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shards = _emulated_shards(n, world)
    for p in (0, 1):
        table.fill_(float("nan"))            # every row must be rewritten by its owner
        for f, c, lo, hi in shards:
            _lib.check(_lib.lib.dfcsr_beam_stats_partial(p, _ptr(x[lo:hi]), _ptr(z[lo:hi]), _ptr(pz[lo:hi]), _ptr(px[lo:hi]),
                                                         hi - lo, n, f, c, ctr, _ptr(d_stats), _ptr(table), None, 0, st))
        _lib.check(_lib.lib.dfcsr_beam_stats_final(p, _ptr(table), n, ctr, 1, 1, _ptr(d_stats), st))
    for f, c, lo, hi in shards:
        _lib.check(_lib.lib.dfcsr_beam_cov_partial(*[_ptr(q[lo:hi]) for q in coords], hi - lo, n, f, c, ctr6, _ptr(ctab), None, 0, st))
    _lib.check(_lib.lib.dfcsr_beam_cov_final(_ptr(ctab), n, ctr6, _ptr(d_cov), st))
    got = d_stats.cpu().numpy()
    assert np.array_equal(got[:14], ref[:14]), (got[:14] - ref[:14])
    flat = d_cov.cpu().numpy()
    assert np.array_equal(flat[:6], rmean) and np.array_equal(flat[6:], rcov[np.triu_indices(6)])
    # and the numbers are the right ones (np.std / np.polyfit / np.cov in fp64 on the host)
    assert abs(got[_lib.S_SIGMA_X] / np.std(b[0]) - 1) < 1e-11 and abs(got[_lib.S_SLOPE] / np.polyfit(b[4], b[0], 1)[0] - 1) < 1e-10
    assert np.allclose(rcov, np.cov(b), rtol=1e-9, atol=0)


@pytest.mark.parametrize("shape", [(100, 100), (300, 300), (64, 512)])
@pytest.mark.parametrize("world", [2, 8])
def test_fixed_point_deposit_is_independent_of_the_sharding(dev, shape, world):
    """dfcsr_deposit_cic (one GPU, all particles) against dfcsr_deposit_cic_q per shard + dfcsr_deposit_cic_finish over
    the shards' integer grids: bitwise equal, and equal to the serial oracle within the deposit gate."""
    import ctypes as C
    import torch
    from pydfcsr_b200 import ops, synth
    n = 600_011
    b = synth.gaussian_bunch(n, seed=5, tilt=2.5 if shape[0] != shape[1] else 0.0)
    x, z, px = _up(b[0], dev), _up(b[4], dev), _up(b[1], dev)
    nx, nz = shape
    lim = (float(np.mean(b[0]) - 5 * np.std(b[0])), float(np.mean(b[0]) + 5 * np.std(b[0])),
           float(np.mean(b[4]) - 5 * np.std(b[4])), float(np.mean(b[4]) + 5 * np.std(b[4])))
    c_ref, v_ref = (t.clone() for t in ops.deposit_cic(x, z, px, nx, lim[0], lim[1], nz, lim[2], lim[3]))
    amax = float(np.max(np.abs(b[1])))
    qs = []
    for f, c, lo, hi in _emulated_shards(n, world):
        q = torch.empty(2 * nx * nz, dtype=torch.int64, device=dev)
        ops.deposit_cic_q(x[lo:hi], z[lo:hi], px[lo:hi], n, nx, lim[0], lim[1], nz, lim[2], lim[3], amax, q)
        qs.append(q)
    ptrs = (C.c_uint64 * world)(*[q.data_ptr() for q in qs])
    c_got, v_got = ops.deposit_cic_finish(ptrs, n, nx, nz, amax, device=dev)
    assert torch.equal(c_got, c_ref) and torch.equal(v_got, v_ref)
    ref = O.cic_deposit_2d(b[0], b[4], np.ones(n), nx, lim[0], lim[1], nz, lim[2], lim[3])
    assert _rel(c_got.cpu().numpy(), ref) < 1e-12
    refv = O.cic_deposit_2d(b[0], b[4], b[1], nx, lim[0], lim[1], nz, lim[2], lim[3])
    assert _rel(v_got.cpu().numpy(), refv) < 1e-12


def test_track_element_kernel_reproduces_bmadx_known_answers(dev):
    """dfcsr_track_element (the device tracker CSR2D.run uses) on the reference's Bmad-X known answers
    (test/test_BmadX_tracking.ipynb cells 25, 28, 31; the numbers live in tests/test_tracking.py) to 1e-12, and on a
    whole bunch against the host restatement of the same maps for every element type and fringe variant."""
    import torch
    from pydfcsr_b200 import synth, tracking
    from tests.test_tracking import KNOWN, MC2, P0C
    for name, (element, want) in KNOWN.items():
        coords = tuple(torch.full((33,), 1e-3, dtype=torch.float64, device=dev) for _ in range(6))
        got = tracking.track_exact(coords, element, P0C, MC2)
        for g, w in zip(got, want):
            g = g.cpu().numpy()
            assert np.all(g == g[0]) and abs(g[0] - w) <= 1e-12 * abs(w), (name, g[0], w)
    b = synth.gaussian_bunch(50_001, seed=17, tilt=0.3)
    b[2] *= 3.0
    elements = [tracking.Drift(0.37), tracking.SBend(L=0.5002, G=0.0483 / 0.5002, E1=0.0, E2=0.0483),
                tracking.SBend(L=0.1, G=-0.0483 / 0.5002, E1=-0.0483, E2=0.0, FRINGE_AT="entrance_end"),
                tracking.SBend(L=0.1, G=0.3, E1=0.1, E2=0.2, FRINGE_AT="no_end"), tracking.SBend(L=0.2, G=0.0),
                tracking.Quadrupole(L=0.2, K1=1.7), tracking.Quadrupole(L=0.2, K1=-0.9, NUM_STEPS=4), tracking.Sextupole(L=0.1, K2=3.0)]
    for el in elements:
        want = tracking.track_exact(tuple(b), el, 5.0e9)
        dev_coords = tuple(_up(c, dev) for c in b)
        got = tracking.track_exact(dev_coords, el, 5.0e9)
        for k in range(6):
            scale = max(np.max(np.abs(want[k])), 1e-300)
            # FMA contraction on the device; z of a bend is a difference of O(L) path lengths (measured 1.4e-13)
            assert np.max(np.abs(got[k].cpu().numpy() - want[k])) <= 1e-12 * scale, (el, k)


def test_get_df_composite_is_bitwise_the_three_calls(dev):
    """dfcsr_get_df (one binding call) = dfcsr_deposit_cic_q + dfcsr_deposit_cic_finish + dfcsr_make_df: the same five
    kernels with the same arguments, so fields, scalars and deposit grids must be the same bits (deposit.py:145-245)."""
    import ctypes as C
    import torch
    from pydfcsr_b200 import ops
    from pydfcsr_b200._lib import Axis
    rng = np.random.default_rng(5)
    n = 200_000
    x, z, px = (_up(rng.normal(0, s, n), dev) for s in (1e-4, 2e-4, 1e-5))
    for (nx, nz, window, order) in ((100, 100, 5, 1), (64, 96, 9, 1), (300, 300, 9, 2)):
        xa, za = Axis.make(-5e-4, 5e-4, nx), Axis.make(-1e-3, 1e-3, nz)
        absmax = float(px.abs().max())
        q = torch.empty(2 * nx * nz, dtype=torch.int64, device=dev)
        cmax = torch.zeros(1, dtype=torch.int64, device=dev)
        dep = torch.empty((2, nx, nz), dtype=torch.float64, device=dev)
        ops.deposit_cic_q(x, z, px, n, nx, xa.start, xa.stop, nz, za.start, za.stop, absmax, q)
        c, v = ops.deposit_cic_finish((C.c_uint64 * 1)(q.data_ptr()), n, nx, nz, absmax, out=dep, count_max=cmax)
        f0, s0 = ops.make_df(c, v, xa, za, window, order, 1000.0, count_max=cmax)
        dep0 = dep.clone()
        q2 = torch.empty_like(q)
        cmax2 = torch.zeros_like(cmax)
        dep2 = torch.empty_like(dep)
        f1, s1 = ops.get_df(x, z, px, xa, za, absmax, window, order, 1000.0, q2, dep2, cmax2)
        assert torch.equal(dep2, dep0) and torch.equal(f1, f0) and torch.equal(s1, s0) and torch.equal(cmax2, cmax)
        assert float(f1[0].abs().max()) > 0
