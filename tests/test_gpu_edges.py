"""Edge cases of the wake kernel through the C ABI: degenerate quadrature meshes, empty blocks, the largest
supported integration mesh, 64-bit addressing of a multi-GB history, NaN/garbage tolerance, error returns."""
import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from tests import scenario

pytestmark = pytest.mark.gpu


def _up(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def _problem(tilt=0.0):
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    st, lat = sc["stack"], sc["lattice"]
    hist = ops.DeviceHistory.from_stacks([st.data[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z,
                                         st.delta_x, st.delta_y, st.delta_z, "cuda:0")
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, "cuda:0")
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 3, 4)
    return sc, hist, dlat, xm, zm


@pytest.mark.parametrize("nx,nz", [(1, 40), (40, 1), (1, 1), (2, 2), (7, 33), (33, 64)])
def test_degenerate_and_ragged_integration_meshes(nx, nz):
    """np.trapz over a single node is 0; node counts that are not multiples of the warp size leave
    partially filled sweeps."""
    from pydfcsr_b200 import ops
    sc, hist, dlat, xm, zm = _problem()
    wp = ops.wake_params(nx=nx, nz=nz, **sc["wake_scalars"])
    de, kick = ops.wake_mesh(hist, dlat, wp, _up(xm), _up(zm))
    ref_de, ref_kick = O.wake_mesh(xm, zm, O.WakeScalars(nx=nx, nz=nz, **sc["wake_scalars"]), sc["lattice"], sc["stack"])
    if nz == 1 or (nx == 1):
        assert not ref_de.any() and not de.cpu().numpy().any() and not kick.cpu().numpy().any()
    else:
        assert _rel(de.cpu().numpy(), ref_de) < 1e-10 and _rel(kick.cpu().numpy(), ref_kick) < 1e-10


def test_zero_formation_window_and_early_lattice():
    """formation_window = 0 collapses the far rectangle (s1 == s2); a beam 5 mm into the lattice makes the
    rectangles run backwards into negative s' (CSR.py:481,539: s2 is not clamped, s1 is)."""
    from pydfcsr_b200 import ops
    sc, hist, dlat, xm, zm = _problem()
    for changes in (dict(formation_window=0.0), dict(t=0.005)):
        scal = dict(sc["wake_scalars"], **changes)
        if "t" in changes:     # history times must still bracket t: shift the time axis with the beam
            shift = sc["wake_scalars"]["t"] - changes["t"]
            h2 = ops.DeviceHistory(hist.ring, hist.head, hist.T, hist.min_t - shift, hist.min_x, hist.min_z,
                                   hist.delta_t, hist.delta_x, hist.delta_z)
            st = sc["stack"]
            ost = O.HistoryStack(st.data, st.min_x - shift, st.min_y, st.min_z, st.delta_x, st.delta_y, st.delta_z)
        else:
            h2, ost = hist, sc["stack"]
        wp = ops.wake_params(nx=30, nz=30, **scal)
        de, kick = ops.wake_mesh(h2, dlat, wp, _up(xm), _up(zm))
        ref_de, ref_kick = O.wake_mesh(xm, zm, O.WakeScalars(nx=30, nz=30, **scal), sc["lattice"], ost)
        assert np.all(np.isfinite(ref_de))
        assert _rel(de.cpu().numpy(), ref_de) < 1e-10 and _rel(kick.cpu().numpy(), ref_kick) < 1e-10


def test_empty_block_and_error_returns():
    import torch
    from pydfcsr_b200 import _lib, ops
    sc, hist, dlat, xm, zm = _problem()
    wp = ops.wake_params(nx=8, nz=8, **sc["wake_scalars"])
    de, kick = ops.wake_mesh(hist, dlat, wp, _up(xm), _up(zm), first=5, count=0)
    assert de.numel() == 0 and kick.numel() == 0
    with pytest.raises(_lib.DfcsrError, match="shared memory"):
        ops.wake_mesh(hist, dlat, ops.wake_params(nx=8, nz=4000, **sc["wake_scalars"]), _up(xm), _up(zm))
    with pytest.raises(_lib.DfcsrError, match="at least one node"):
        ops.wake_mesh(hist, dlat, ops.wake_params(nx=0, nz=8, **sc["wake_scalars"]), _up(xm), _up(zm))
    bad = ops.DeviceHistory(hist.ring, 99, hist.T, hist.min_t, hist.min_x, hist.min_z, hist.delta_t, hist.delta_x, hist.delta_z)
    with pytest.raises(_lib.DfcsrError, match="ring geometry"):
        ops.wake_mesh(bad, dlat, wp, _up(xm), _up(zm))
    with pytest.raises(_lib.DfcsrError):
        ops.wake_mesh(hist, dlat, wp, torch.zeros(4, dtype=torch.float32, device="cuda"), _up(zm))   # wrong dtype
    # a NaN observation point poisons only itself (the reference returns NaN/0 there), its neighbours are untouched
    xm2 = xm.copy()
    xm2[3] = np.nan
    de2, _ = ops.wake_mesh(hist, dlat, wp, _up(xm2), _up(zm))
    de1, _ = ops.wake_mesh(hist, dlat, wp, _up(xm), _up(zm))
    keep = np.arange(xm.size) != 3
    assert torch.equal(de2[torch.from_numpy(keep).cuda()], de1[torch.from_numpy(keep).cuda()])


def test_largest_integration_mesh():
    """zbins = 640 is close to the shared-memory limit of the node table (3 x 640 x 72 B = 138 KB)."""
    from pydfcsr_b200 import ops
    sc, hist, dlat, xm, zm = _problem()
    wp = ops.wake_params(nx=16, nz=640, **sc["wake_scalars"])
    de, kick = ops.wake_mesh(hist, dlat, wp, _up(xm[:3]), _up(zm[:3]))
    ref_de, ref_kick = O.wake_mesh(xm[:3], zm[:3], O.WakeScalars(nx=16, nz=640, **sc["wake_scalars"]), sc["lattice"], sc["stack"])
    assert _rel(de.cpu().numpy(), ref_de) < 1e-10 and _rel(kick.cpu().numpy(), ref_kick) < 1e-10


def test_multi_gigabyte_history_addressing():
    """A 2000 x 2000 history (the reference's upper_limit) with 12 slots is 2.3 GB: voxel offsets exceed 2^31
    bytes and 2^28 doubles.  Analytic fields, ring wrapped so that the window straddles the end of the buffer."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry()
    lat, scal = sc["lattice"], sc["wake_scalars"]
    T, X, Z = 5, 2000, 2000
    sx, sz = scal["sigma_x"], scal["sigma_z"]
    xg = np.linspace(-5 * sx, 5 * sx, X)
    zg = np.linspace(-5 * sz, 5 * sz, Z)
    tg = scal["t"] - 0.1 * np.arange(T)[::-1]
    G = np.exp(-0.5 * (xg[:, None] / sx) ** 2 - 0.5 * (zg[None, :] / sz) ** 2) / (2 * np.pi * sx * sz)
    stacks = {"density": np.stack([G * (1 + 0.01 * k) for k in range(T)]),
              "density_x": np.stack([-(xg[:, None] / sx ** 2) * G for _ in range(T)]),
              "density_z": np.stack([-(zg[None, :] / sz ** 2) * G for _ in range(T)]),
              "vx": np.stack([1e-6 * np.tanh(xg[:, None] / sx) * np.ones_like(G) for _ in range(T)]),
              "vx_x": np.stack([1e-6 / sx / np.cosh(xg[:, None] / sx) ** 2 * np.ones_like(G) for _ in range(T)])}
    st = O.HistoryStack(stacks, tg[0], xg[0], zg[0], (tg[-1] - tg[0]) / (T - 1), (xg[-1] - xg[0]) / (X - 1), (zg[-1] - zg[0]) / (Z - 1))
    hist = ops.DeviceHistory.from_stacks([stacks[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z, st.delta_x,
                                         st.delta_y, st.delta_z, "cuda:0", cap=12, head=10)
    assert hist.ring.numel() * 8 > 2 ** 31
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, "cuda:0")
    xm = np.array([0.0, 0.4 * sx, -1.1 * sx, 2.0 * sx])
    zm = np.array([0.0, -0.7 * sz, 1.3 * sz, 2.5 * sz])
    wp = ops.wake_params(nx=40, nz=40, **scal)
    de, kick = ops.wake_mesh(hist, dlat, wp, _up(xm), _up(zm))
    ref_de, ref_kick = O.wake_mesh(xm, zm, O.WakeScalars(nx=40, nz=40, **scal), lat, st)
    assert _rel(de.cpu().numpy(), ref_de) < 1e-10 and _rel(kick.cpu().numpy(), ref_kick) < 1e-10
    del hist
    torch.cuda.empty_cache()
