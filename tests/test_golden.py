"""Golden vectors produced by the reference itself (oracle/make_golden.py, committed under
tests/golden/).  The CPU tests pin the oracle on them; the GPU tests pin the CUDA path on them."""
import os

import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from tests import scenario

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WAKE_CASES = ["notilt", "tilt_pos", "tilt_neg"]
DF_CASES = ["untilted", "tilted_o1", "tilted_o2"]


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def test_known_answer_trilinear():
    """The one known-answer vector the reference holds for this path: test/test_interpolation.ipynb
    cell 13 stores interp([2.1, 6.2, 8.3]) = 125.80469388 for f = 2x^3 + 3y^2 - z.  That stored output
    is the value on the 11 x 22 x 33 grid of scipy's RegularGridInterpolator documentation example
    (second documented point: [3.3, 5.2, 7.1] -> 146.30069388); on the notebook's later 200x200x100
    grid the same query gives 125.5422, so the stored cell output predates the grid edit."""
    x, y, z = np.linspace(1, 4, 11), np.linspace(4, 7, 22), np.linspace(7, 9, 33)
    xg, yg, zg = np.meshgrid(x, y, z, indexing="ij")
    data = 2 * xg ** 3 + 3 * yg ** 2 - zg
    v = O.interp3d(np.array([2.1, 3.3]), np.array([6.2, 5.2]), np.array([8.3, 7.1]), data, x[0], y[0], z[0],
                   np.mean(np.diff(x)), np.mean(np.diff(y)), np.mean(np.diff(z)))
    assert abs(v[0] - 125.80469388) < 5e-9 and abs(v[1] - 146.30069388) < 5e-9


def test_mpi_split_rule():
    g = np.load(os.path.join(G, "mpi_split.npz"))
    from pydfcsr_b200.distributed import split_counts
    for n, p in g["cases"]:
        for fn in (O.split_counts, split_counts):
            count, displ = fn(int(n), int(p))
            assert np.array_equal(count, g[f"count_{n}_{p}"]) and np.array_equal(displ, g[f"displ_{n}_{p}"])
            assert sum(count) == n


def _bunch(g, tilt=0.0):
    from pydfcsr_b200 import synth
    return synth.gaussian_bunch(int(g["n_particle"]), seed=int(g["seed"]), tilt=tilt)


def test_oracle_cic_golden():
    g = np.load(os.path.join(G, "cic.npz"))
    b = _bunch(g)
    xs, xe, zs, ze = g["bounds"]
    assert np.array_equal(O.cic_deposit_2d(b[0], b[4], np.ones_like(b[0]), 37, xs, xe, 53, zs, ze), g["count"])
    assert np.array_equal(O.cic_deposit_2d(b[0], b[4], b[1], 37, xs, xe, 53, zs, ze), g["vxsum"])


@pytest.mark.parametrize("case", DF_CASES)
def test_oracle_df_golden(case):
    g = np.load(os.path.join(G, f"df_{case}.npz"))
    b = _bunch(g, float(g["tilt"]))
    cfg = O.DepositConfig(**dict(scenario.DEPOSIT_CFG, filter_order=int(g["order"]), filter_window=int(g["window"])))
    df = O.make_density_functions(b[0], b[4], b[1], 0.25, cfg)
    assert np.array_equal(df.x_grids, g["x_grids"]) and np.array_equal(df.z_grids, g["z_grids"])
    for k in O.FIELDS:
        assert _rel(getattr(df, k), g[k]) < 1e-13, k


@pytest.mark.parametrize("case", WAKE_CASES)
def test_oracle_wake_golden(case):
    g = np.load(os.path.join(G, f"wake_{case}.npz"))
    sc = scenario.chicane_entry(tilt=float(g["tilt"]))
    st = sc["stack"]
    assert tuple(g["hist_shape"]) == st.shape
    assert np.array_equal(g["meta"], [st.min_x, st.min_y, st.min_z, st.delta_x, st.delta_y, st.delta_z])
    stride = int(g["hist_stride"])
    for k in O.FIELDS:
        assert _rel(st.data[k].ravel()[::stride], g[f"hist_{k}"]) < 1e-13, k
        assert abs(st.data[k].sum() - g[f"hist_{k}_sum"][0]) <= 1e-11 * g[f"hist_{k}_sum"][1]
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, xr, zr = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 5, 7)
    assert np.array_equal(xm, g["xmesh"]) and np.array_equal(zm, g["zmesh"])
    osc = O.WakeScalars(nx=50, nz=50, **sc["wake_scalars"])
    de, kick = O.wake_mesh(xm, zm, osc, sc["lattice"], st)
    assert _rel(de, g["dE_dct"].ravel()) < 1e-12 and _rel(kick, g["x_kick"].ravel()) < 1e-12
    # debug integrands of one point
    s_d, x_d = g["dbg_point"]
    osc24 = O.WakeScalars(nx=24, nz=24, **sc["wake_scalars"])
    iz, ix = [], []
    for (xa, xb, n_x, sa, sb, n_s) in O.wake_regions(s_d, x_d, osc24):
        xmm, smm = np.meshgrid(np.linspace(xa, xb, n_x), np.linspace(sa, sb, n_s), indexing="ij")
        a, b = O.wake_integrand(s_d, x_d, osc24, sc["lattice"], st, xmm.ravel(), smm.ravel())
        iz.append(a); ix.append(b)
    assert _rel(np.concatenate(iz), g["dbg_integrand_z"]) < 1e-13
    assert _rel(np.concatenate(ix), g["dbg_integrand_x"]) < 1e-13
    # kick
    px, pz = sc["coords"][1], sc["coords"][5]
    pxn, pzn = O.apply_kick(x, z, px, pz, g["dE_dct"], g["x_kick"], g["xrange"], g["zrange"], 0.1, 5e9, True)
    k = int(g["kick_stride"])
    assert np.array_equal(pzn[::k], g["pz_new"]) and np.array_equal(pxn[::k], g["px_new"])


# ------------------------------------------------------------------------------------------------- GPU
def _up(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.gpu
def test_gpu_cic_golden():
    from pydfcsr_b200 import ops
    g = np.load(os.path.join(G, "cic.npz"))
    b = _bunch(g)
    xs, xe, zs, ze = (float(v) for v in g["bounds"])
    for mode in (0, 1, 2, 4, 5):
        c, v = ops.deposit_cic(_up(b[0]), _up(b[4]), _up(b[1]), 37, xs, xe, 53, zs, ze, mode=mode)
        assert _rel(c.cpu().numpy(), g["count"]) < 1e-12 and _rel(v.cpu().numpy(), g["vxsum"]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("case", DF_CASES)
def test_gpu_df_golden(case):
    from pydfcsr_b200 import DF_tracker
    g = np.load(os.path.join(G, f"df_{case}.npz"))
    b = _bunch(g, float(g["tilt"]))
    trk = DF_tracker(dict(scenario.DEPOSIT_CFG, filter_order=int(g["order"]), filter_window=int(g["window"])), device="cuda:0")
    trk.get_DF(_up(b[0]), _up(b[4]), _up(b[1]), 0.25)
    assert _rel(trk.x_grids, g["x_grids"]) < 1e-13 and _rel(trk.z_grids, g["z_grids"]) < 1e-13
    for k in O.FIELDS:
        assert _rel(getattr(trk, k), g[k]) < 1e-10, k
    assert abs(float(trk._current.scalars[4]) - float(g["mean_vx_x"])) <= 1e-10 * np.max(np.abs(g["vx_x"]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", WAKE_CASES)
def test_gpu_wake_golden(case):
    """Whole CUDA path (deposit -> history -> wake -> kick) from regenerated particle batches against
    numbers the reference produced."""
    import torch
    from pydfcsr_b200 import DF_tracker, ops
    from pydfcsr_b200._lib import Axis
    g = np.load(os.path.join(G, f"wake_{case}.npz"))
    sc = scenario.chicane_entry(tilt=float(g["tilt"]))
    trk = DF_tracker(dict(scenario.DEPOSIT_CFG), device="cuda:0")
    for st in sc["steps"]:
        x, px, y, py, z, pz = st["coords"]
        trk.get_DF(_up(x), _up(z), _up(px), st["pos"])
        trk.append_DF()
        trk.append_interpolant(st["formation_length"], 1)
    trk.build_interpolant()
    assert (trk.history.T,) + tuple(trk._ring.shape[1:3]) == tuple(g["hist_shape"])
    stride = int(g["hist_stride"])
    for k in O.FIELDS:
        assert _rel(getattr(trk, f"data_{k}_interp").ravel()[::stride], g[f"hist_{k}"]) < 1e-10, k
    t, sx, sz, slope0, slope1, mx, mz, fl, scaling = g["scalars"]
    lat = sc["lattice"]
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, "cuda:0")
    wp = ops.wake_params(t, sx, sz, slope0, mx, fl, scaling, 50, 50)
    de, kick = ops.wake_mesh(trk.history, dlat, wp, _up(g["xmesh"]), _up(g["zmesh"]))
    assert _rel(de.cpu().numpy(), g["dE_dct"].ravel()) < 1e-10
    assert _rel(kick.cpu().numpy(), g["x_kick"].ravel()) < 1e-10
    wp24 = ops.wake_params(t, sx, sz, slope0, mx, fl, scaling, 24, 24)
    dbg = ops.wake_point_debug(trk.history, dlat, wp24, float(g["dbg_point"][0]), float(g["dbg_point"][1]))
    assert _rel(np.concatenate([r["integrand_z"].ravel() for r in dbg]), g["dbg_integrand_z"]) < 1e-10
    assert _rel(np.concatenate([r["integrand_x"].ravel() for r in dbg]), g["dbg_integrand_x"]) < 1e-10
    x, px, y, py, z, pz = sc["coords"]
    dpx, dpz = _up(px), _up(pz)
    ops.apply_kick(_up(x), _up(z), dpx, dpz, slope0, slope1, _up(g["dE_dct"]), _up(g["x_kick"]),
                   Axis.make(g["xrange"][0], g["xrange"][-1], 5), Axis.make(g["zrange"][0], g["zrange"][-1], 7), 0.1, 5e9, True)
    k = int(g["kick_stride"])
    assert _rel(dpz.cpu().numpy()[::k] - pz[::k], g["pz_new"] - pz[::k]) < 1e-10
    assert _rel(dpx.cpu().numpy()[::k] - px[::k], g["px_new"] - px[::k]) < 1e-10


# ------------------------------------------------------------------------- 2-D Savitzky-Golay (SGolay_filter.py)
def _sgolay_cases():
    from oracle.make_golden import sgolay2d_input
    g = np.load(os.path.join(G, "sgolay2d.npz"))
    z = sgolay2d_input()
    assert float(z.sum()) == float(g["checksum"])        # same seeded input as the generating run
    return z, g


def test_oracle_sgolay2d_golden():
    z, g = _sgolay_cases()
    for window, order in g["cases"]:
        window, order = int(window), int(order)
        assert _rel(O.sgolay2d(z, window, order), g[f"smooth_{window}_{order}"]) < 1e-13
        col, row = O.sgolay2d(z, window, order, "both")
        assert _rel(col, g[f"col_{window}_{order}"]) < 1e-13 and _rel(row, g[f"row_{window}_{order}"]) < 1e-13


@pytest.mark.gpu
def test_gpu_sgolay2d_golden():
    from pydfcsr_b200 import ops
    z, g = _sgolay_cases()
    dz = _up(z)
    for window, order in g["cases"]:
        window, order = int(window), int(order)
        assert _rel(ops.sgolay2d(dz, window, order).cpu().numpy(), g[f"smooth_{window}_{order}"]) < 1e-13
        col, row = ops.sgolay2d(dz, window, order, "both")
        assert _rel(col.cpu().numpy(), g[f"col_{window}_{order}"]) < 1e-13
        assert _rel(row.cpu().numpy(), g[f"row_{window}_{order}"]) < 1e-13
        assert _rel(ops.sgolay2d(dz, window, order, "row").cpu().numpy(), g[f"row_{window}_{order}"]) < 1e-13


@pytest.mark.parametrize("tilt,n_bend_steps", [(0.0, 4), (2.5, 4), (-2.5, 4), (2.5, 2), (-1.2, 3)])
def test_skip_plan_never_drops_a_contributing_sample(tilt, n_bend_steps):
    """The zero-density skipping rules of the wake kernel, restated in the oracle (row hulls + coarse s' bracket per
    x' node): over every rectangle of several observation points, a sample the plan does not gather must be one
    whose integrands are exactly zero in the full evaluation."""
    sc = scenario.chicane_entry(tilt=tilt, n_bend_steps=n_bend_steps)
    hs, lat = sc["stack"], sc["lattice"]
    osc = O.WakeScalars(nx=24, nz=70, **sc["wake_scalars"])
    sup = O.row_support(hs)
    ws = sc["wake_scalars"]
    skipped = total = 0
    for dz, dx in ((0.4, 0.3), (-2.0, 1.5), (2.5, -2.0)):
        s_obs = sc["pos"] + dz * ws["sigma_z"]
        x_obs = dx * ws["sigma_x"] + ws["slope0"] * dz * ws["sigma_z"]
        for (xa, xb, n_x, sa, sb, n_s) in O.wake_regions(s_obs, x_obs, osc):
            xn, sn = np.linspace(xa, xb, n_x), np.linspace(sa, sb, n_s)
            xm, sm = np.meshgrid(xn, sn, indexing="ij")
            with np.errstate(all="ignore"):
                iz, ix = O.wake_integrand(s_obs, x_obs, osc, lat, hs, xm.ravel(), sm.ravel())
            contributes = ((iz != 0) | (ix != 0)).reshape(n_x, n_s)
            for i in range(n_x):
                swept, gathered = O.skip_plan(s_obs, x_obs, xn[i], sa, sb, n_s, osc, lat, hs, support=sup)
                assert not np.any(contributes[i] & ~gathered), (tilt, dz, dx, i)
                skipped += int((~gathered).sum())
                total += n_s
    assert 0 < skipped < total
