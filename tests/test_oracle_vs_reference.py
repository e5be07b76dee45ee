"""Pins the oracle restatement against the UNMODIFIED reference imported from /root/reference
(build container only; skipped where the checkout is absent, e.g. on the GPU box)."""
import warnings

import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from oracle import refstub
from tests import scenario

pytestmark = pytest.mark.reference


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ref():
    warnings.simplefilter("ignore")
    return refstub.load_reference()


def test_trilinear_and_linear_gathers(ref):
    from pyDFCSR_2D.interp1D import interpolate1D
    from pyDFCSR_2D.interp3D import interpolate3D
    rng = np.random.default_rng(0)
    data = rng.normal(size=(4, 7, 9))
    # queries straddling every edge rule: (-1,0) extrapolation, clamp at n-1, outside, NaN
    q = np.concatenate([rng.uniform(-1.5, 10.5, 4000), [np.nan, 3.0, 6.0, 8.0, -0.3, -1.0]])
    t, y, z = q * 0.4, np.roll(q, 1) * 0.7, np.roll(q, 2)
    a = O.interp3d(t, y, z, data, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0)
    b = interpolate3D(t, y, z, data, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0)
    assert np.array_equal(a, b)
    tab = rng.normal(size=31)
    assert np.array_equal(O.interp1d(q * 3, tab, 0.5, 0.9), interpolate1D(q * 3, tab, 0.5, 0.9))


def test_cic(ref):
    from pyDFCSR_2D.deposit import histogram_cic_2d
    rng = np.random.default_rng(1)
    x, z, w = rng.normal(size=50_000), rng.normal(size=50_000), rng.normal(size=50_000)
    a = O.cic_deposit_2d(x, z, w, 33, -2.0, 2.5, 41, -3.0, 3.0)
    b = histogram_cic_2d(x, z, w, 33, -2.0, 2.5, 41, -3.0, 3.0)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("window,order", [(5, 0), (5, 1), (5, 2), (9, 1), (9, 2), (7, 3)])
def test_savgol_matches_scipy(window, order):
    from scipy.signal import savgol_filter
    a = np.random.default_rng(2).normal(size=(40, 55))
    ref_ = savgol_filter(savgol_filter(a, window, order, axis=0), window, order, axis=1)
    assert _rel(O.savgol_separable(a, window, order), ref_) < 5e-13


def test_gradient_matches_numpy():
    a = np.random.default_rng(3).normal(size=(30, 44))
    xg, zg = np.linspace(-1e-4, 3e-4, 30), np.linspace(-1e-3, 1e-3, 44)
    gx, gz = np.gradient(a, xg, zg)
    assert np.array_equal(O.gradient_axis(a, xg, 0), gx) and np.array_equal(O.gradient_axis(a, zg, 1), gz)
    xu = np.arange(30.0)       # bit-uniform spacing takes numpy's other branch
    assert np.array_equal(O.gradient_axis(a, xu, 0), np.gradient(a, xu, axis=0))


def test_regrid_matches_scipy():
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(4)
    src = rng.normal(size=(21, 17))
    xg, zg = np.linspace(-1, 2, 21), np.linspace(0, 5, 17)
    xq, zq = np.linspace(-1.5, 2, 64), np.linspace(0, 4.9, 50)
    X, Z = np.meshgrid(xq, zq, indexing="ij")
    ref_ = RegularGridInterpolator((xg, zg), src, method="linear", fill_value=0.3, bounds_error=False)((X, Z))
    assert np.array_equal(O.regrid_bilinear(src, xg, zg, xq, zq, 0.3), ref_)


@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_get_df(ref, tilt):
    from pyDFCSR_2D.deposit import DF_tracker
    from pydfcsr_b200 import synth
    b = synth.gaussian_bunch(150_000, seed=9, tilt=tilt)
    tr = DF_tracker(dict(scenario.DEPOSIT_CFG))
    tr.get_DF(b[0], b[4], b[1], 0.3)
    df = O.make_density_functions(b[0], b[4], b[1], 0.3, O.DepositConfig(**scenario.DEPOSIT_CFG))
    assert df.density.shape == tr.density.shape
    for k in ("x_grids", "z_grids"):
        assert np.array_equal(getattr(df, k), getattr(tr, k))
    for k in O.FIELDS:
        assert _rel(getattr(df, k), getattr(tr, k)) < 1e-13, k


@pytest.mark.parametrize("tilt", [0.0, -2.5])
def test_history_and_wake(ref, tilt, tmp_path):
    import yaml
    from pydfcsr_b200 import synth
    lat_yaml = str(tmp_path / "lat.yaml")
    with open(lat_yaml, "w") as fh:
        yaml.safe_dump(dict(synth.chicane_lattice_config()), fh, sort_keys=False)
    sc = scenario.chicane_entry(tilt=tilt)
    csr = refstub.make_reference_csr(lat_yaml, scenario.DEPOSIT_CFG, dict(n_formation_length=1, zbins=30, xbins=30),
                                     dict(xbins=3, zbins=4, xlim=3, zlim=3, workdir=str(tmp_path)))
    for st in sc["steps"]:
        x, px, y, py, z, pz = st["coords"]
        csr.DF_tracker.get_DF(x=x, z=z, px=px, t=st["pos"])
        csr.DF_tracker.append_DF()
        csr.DF_tracker.append_interpolant(formation_length=st["formation_length"], n_formation_length=1)
    csr.DF_tracker.build_interpolant()
    tr, hs = csr.DF_tracker, sc["stack"]
    for fld, attr in zip(O.FIELDS, ("data_density_interp", "data_density_x_interp", "data_density_z_interp",
                                    "data_vx_interp", "data_vx_x_interp")):
        assert _rel(hs.data[fld], getattr(tr, attr)) < 1e-13
    assert (hs.min_x, hs.min_y, hs.min_z, hs.delta_x, hs.delta_y, hs.delta_z) == \
           (tr.min_x, tr.min_y, tr.min_z, tr.delta_x, tr.delta_y, tr.delta_z)
    lat = sc["lattice"]
    assert np.array_equal(lat.coords, csr.lattice.coords) and np.array_equal(lat.n_vec, csr.lattice.n_vec)
    assert np.array_equal(lat.rho, csr.lattice.rho) and lat.delta_s == csr.lattice.delta_x
    x, px, y, py, z, pz = sc["coords"]
    csr.beam = refstub.FakeBeam(x, px, z, pz, sc["pos"])
    csr.CSR_scaling = 8.98755e3 * 1e-9
    csr.formation_length = sc["steps"][-1]["formation_length"]
    csr.get_CSR_mesh()
    s = sc["scalars"]
    xm, zm, xr, zr = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 3, 4)
    assert np.array_equal(xm, csr.CSR_xmesh) and np.array_equal(zm, csr.CSR_zmesh)
    csr.calculate_2D_CSR()
    de, kick = O.wake_mesh(xm, zm, O.WakeScalars(nx=30, nz=30, **sc["wake_scalars"]), lat, hs)
    assert _rel(de, csr.dE_dct.ravel()) < 1e-12 and _rel(kick, csr.x_kick.ravel()) < 1e-12


def test_product_lattice_matches_reference(ref, tmp_path):
    import yaml
    from pyDFCSR_2D.lattice import Lattice as RefLattice
    from pydfcsr_b200 import synth
    from pydfcsr_b200.lattice import Lattice
    path = str(tmp_path / "lat.yaml")
    with open(path, "w") as fh:
        yaml.safe_dump(dict(synth.chicane_lattice_config()), fh, sort_keys=False)
    a, b = Lattice({"lattice_input_file": path}), RefLattice({"lattice_input_file": path})
    for k in ("coords", "n_vec", "tau_vec", "rho", "distance", "nsep", "steps_per_element", "steps_record"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.delta_x == b.delta_x and a.total_steps == b.total_steps and a.step_size == b.step_size
    assert np.array_equal(a.CSR_steps_index, b.CSR_steps_index)


@pytest.mark.parametrize("shape,window,order", [((40, 50), 5, 2), ((64, 96), 9, 1), ((30, 30), 7, 3), ((11, 13), 11, 4),
                                                ((100, 100), 15, 2), ((9, 9), 3, 1), ((33, 65), 25, 2)])
def test_sgolay2d_matches_reference(ref, shape, window, order):
    """SGolay_filter.py:3-81 (border extension, kernel order and signs, 'valid' convolution)."""
    from pyDFCSR_2D.SGolay_filter import sgolay2d
    from pydfcsr_b200 import ops
    rng = np.random.default_rng(window)
    z = rng.normal(size=shape) + np.linspace(0, 3, shape[1])[None, :] ** 2 - np.linspace(-1, 2, shape[0])[:, None]
    half = window // 2
    import scipy.signal  # noqa: F401  (the reference file uses scipy.signal without importing the submodule)
    assert np.array_equal(O.sgolay2d_pad(z, half)[half:-half, half:-half], z)
    for d in (None, "col", "row"):
        assert _rel(O.sgolay2d(z, window, order, d), sgolay2d(z, window, order, d)) < 5e-14
    both_o, both_r = O.sgolay2d(z, window, order, "both"), sgolay2d(z, window, order, "both")
    assert _rel(both_o[0], both_r[0]) < 5e-14 and _rel(both_o[1], both_r[1]) < 5e-14
    # the product's host-side kernels are the oracle's
    assert np.allclose(ops.sgolay2d_kernels(window, order), O.sgolay2d_kernels(window, order), rtol=0, atol=1e-15)
    for bad in ((4, 1), (3, 4)):
        for fn in (lambda: sgolay2d(z, *bad), lambda: O.sgolay2d(z, *bad), lambda: ops.sgolay2d_kernels(*bad)):
            with pytest.raises(ValueError):
                fn()


def test_twiss_matches_reference(ref):
    """twiss.py:2-71 on a chirped, dispersed bunch."""
    from types import SimpleNamespace
    from pyDFCSR_2D.twiss import twiss_from_bmadx_particles
    from pydfcsr_b200 import synth
    b = synth.gaussian_bunch(50_000, seed=9, tilt=0.7)
    b[0] += 0.02 * b[5]                       # some dispersion
    p = SimpleNamespace(x=b[0], px=b[1], y=b[2], py=b[3], z=b[4], pz=b[5], p0c=5.0e9, mc2=0.51099895e6)
    want = twiss_from_bmadx_particles(p)
    got = O.twiss_from_coords(b, 5.0e9, 0.51099895e6)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-12 * abs(want[k]), k


@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_samples_without_density_contribute_exactly_zero_in_the_reference(ref, tilt, tmp_path):
    """The property K4's zero-density skipping rests on, checked on the UNMODIFIED reference:
    wherever the eight history voxels of a sample hold density = d(density)/dx = d(density)/dz = 0, the reference's own
    get_CSR_integrand (CSR.py:605-782) returns exactly 0 for both integrands -- every term carries rho' or grad rho'."""
    import yaml
    from pydfcsr_b200 import synth
    lat_yaml = str(tmp_path / "lat.yaml")
    with open(lat_yaml, "w") as fh:
        yaml.safe_dump(dict(synth.chicane_lattice_config()), fh, sort_keys=False)
    sc = scenario.chicane_entry(tilt=tilt)
    csr = refstub.make_reference_csr(lat_yaml, scenario.DEPOSIT_CFG, dict(n_formation_length=1, zbins=40, xbins=40),
                                     dict(xbins=3, zbins=4, xlim=3, zlim=3, workdir=str(tmp_path)))
    for st in sc["steps"]:
        x, px, y, py, z, pz = st["coords"]
        csr.DF_tracker.get_DF(x=x, z=z, px=px, t=st["pos"])
        csr.DF_tracker.append_DF()
        csr.DF_tracker.append_interpolant(formation_length=st["formation_length"], n_formation_length=1)
    csr.DF_tracker.build_interpolant()
    x, px, y, py, z, pz = sc["coords"]
    csr.beam = refstub.FakeBeam(x, px, z, pz, sc["pos"])
    csr.CSR_scaling = 8.98755e3 * 1e-9
    csr.formation_length = sc["steps"][-1]["formation_length"]
    hs, lat = sc["stack"], sc["lattice"]
    T, X, Z = hs.shape
    nzv = (hs.data["density"] != 0) | (hs.data["density_x"] != 0) | (hs.data["density_z"] != 0)
    osc = O.WakeScalars(nx=40, nz=40, **sc["wake_scalars"])
    s_obs = sc["pos"] + 0.4 * sc["wake_scalars"]["sigma_z"]
    x_obs = 0.3 * sc["wake_scalars"]["sigma_x"] + sc["wake_scalars"]["slope0"] * 0.4 * sc["wake_scalars"]["sigma_z"]
    n_zero = n_in = 0
    for (xa, xb, n_x, sa, sb, n_s) in O.wake_regions(s_obs, x_obs, osc):
        xm, sm = np.meshgrid(np.linspace(xa, xb, n_x), np.linspace(sa, sb, n_s), indexing="ij")
        with np.errstate(all="ignore"):
            iz, ix = csr.get_CSR_integrand(s=s_obs, t=sc["pos"], x=x_obs, xp=xm, sp=sm)
        # locate every sample with the reference's cell rule (interp3D.py:30-52) on the same geometry
        def orbit(q, tab):
            return O.interp1d(q, tab, lat.min_s, lat.delta_s)
        q = sm.ravel()
        rx = orbit(np.array([s_obs]), lat.coords[:, 0])[0] - orbit(q, lat.coords[:, 0]) \
            + x_obs * orbit(np.array([s_obs]), lat.n_vec[:, 0])[0] - xm.ravel() * orbit(q, lat.n_vec[:, 0])
        ry = orbit(np.array([s_obs]), lat.coords[:, 1])[0] - orbit(q, lat.coords[:, 1]) \
            + x_obs * orbit(np.array([s_obs]), lat.n_vec[:, 1])[0] - xm.ravel() * orbit(q, lat.n_vec[:, 1])
        t_ret = sc["pos"] - np.sqrt(rx ** 2 + ry ** 2)
        ut = (t_ret - hs.min_x) / hs.delta_x
        uy = (xm.ravel() - hs.min_y) / hs.delta_y
        uz = (q - t_ret - hs.min_z) / hs.delta_z
        inside = (ut > -1) & (ut < T) & (uy > -1) & (uy < X) & (uz > -1) & (uz < Z)
        t0, y0, z0 = (np.where(inside, u, 0).astype(np.int64) for u in (ut, uy, uz))
        t1, y1, z1 = np.minimum(t0 + 1, T - 1), np.minimum(y0 + 1, X - 1), np.minimum(z0 + 1, Z - 1)
        any_density = np.zeros(q.shape, bool)
        for a in (t0, t1):
            for b in (y0, y1):
                for c in (z0, z1):
                    any_density |= nzv[a, b, c]
        empty = inside & ~any_density
        n_zero += int(empty.sum())
        n_in += int(inside.sum())
        assert np.all(iz.ravel()[empty] == 0.0) and np.all(ix.ravel()[empty] == 0.0)
        assert np.all(iz.ravel()[~inside] == 0.0) and np.all(ix.ravel()[~inside] == 0.0)
    assert n_in > 0 and n_zero > 0
    if tilt:
        assert n_zero > 0.2 * n_in, (n_zero, n_in)         # a chirped bunch leaves most of its +-5 sigma grid empty
