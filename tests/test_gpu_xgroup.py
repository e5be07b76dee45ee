"""GPU parity tests of K4's x-group mapping (dfcsr_wake_grid_xgroups, csrc/wake_xgroup.cuh): one warp lane per
observation point of a mesh row.  Checked against the CPU oracle (1e-10, BASELINE.json north_star), against the
point-per-CTA kernel, and for the property the multi-GPU path rests on: ANY split of the groups over launches gives the
same bits.  Reference lines: CSR.py:397-451 (mesh loop), 454-602 (quadrature), 605-782 (integrand)."""
import ctypes as C

import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from tests import scenario
from tests.test_gpu_kernels import _device_problem, _rel

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def dev():
    import torch
    return torch.device("cuda:0")


def _problem(sc, dev, nx, nz):
    """Device problem with the skipping policy OFF: whether AUTO would pick skipping (and with it the point kernel) depends
    on how much of its grid the scenario's bunch fills; these tests are about the x-group mapping itself."""
    from pydfcsr_b200 import ops
    hist, dlat, _, osc = _device_problem(sc, dev, nx, nz)
    return hist, dlat, ops.wake_params(nx=nx, nz=nz, skip="off", **sc["wake_scalars"]), osc


def _mesh(sc, xbins, zbins):
    from pydfcsr_b200._lib import Axis
    s = sc["scalars"]
    x, z = sc["coords"][0], sc["coords"][4]
    xm, zm, xr, zr = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, xbins, zbins)
    return xm, zm, Axis.make(xr[0], xr[-1], xbins), Axis.make(zr[0], zr[-1], zbins), float(s["slope"][0]), float(s["slope"][1])


@pytest.mark.parametrize("tilt,xbins,zbins,nx,nz,gw", [(0.0, 32, 5, 40, 40, 32), (0.5, 48, 3, 50, 33, 32),
                                                          (-0.7, 70, 2, 24, 64, 32), (0.0, 128, 2, 40, 40, 32)])
def test_xgroup_matches_oracle_and_point_kernel(dev, tilt, xbins, zbins, nx, nz, gw):
    """Full and partial groups (48 = 32 + 16 lanes, 70 = 32 + 32 + 6, 128 = 4 x 32), straight and tilted bunches below the
    chirp-band switch, ragged integration meshes: oracle parity on every mesh point, agreement with the point kernel far
    below the gate, sample accounting identical to the point kernel's."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=tilt)
    hist, dlat, wp, osc = _problem(sc, dev, nx, nz)
    xm, zm, xa, za, slope, icpt = _mesh(sc, xbins, zbins)
    plan = ops.wake_xgroup_plan(hist, wp, xa, za)
    assert plan.group_points == gw and plan.n_groups == zbins * ((xbins + gw - 1) // gw) and plan.unit_nodes >= 1
    cnt = torch.zeros(3, dtype=torch.int64, device=dev)
    de, kick = ops.wake_grid_xgroups(hist, dlat, wp, xa, za, slope, icpt, plan=plan, counters=cnt)
    cnt_p = torch.zeros(3, dtype=torch.int64, device=dev)
    de_p, kick_p = ops.wake_grid(hist, dlat, wp, xa, za, slope, icpt, counters=cnt_p)
    ref_de, ref_kick = O.wake_mesh(xm, zm, osc, sc["lattice"], sc["stack"])
    assert _rel(de.cpu().numpy(), ref_de) < TOL and _rel(kick.cpu().numpy(), ref_kick) < TOL
    assert _rel(de.cpu().numpy(), de_p.cpu().numpy()) < 1e-12 and _rel(kick.cpu().numpy(), kick_p.cpu().numpy()) < 1e-12
    a, b = [int(v) for v in cnt.cpu()], [int(v) for v in cnt_p.cpu()]
    assert a[0] == b[0] and a[1] == b[1] == xbins * zbins * 4 * nx * nz and a[2] == a[0]
    # run to run: same bits (dynamic unit queue, fixed summation order)
    de2, kick2 = ops.wake_grid_xgroups(hist, dlat, wp, xa, za, slope, icpt, plan=plan)
    assert torch.equal(de, de2) and torch.equal(kick, kick2)


@pytest.mark.parametrize("xbins,zbins,gw", [(70, 5, 32), (112, 5, 32)])
def test_xgroup_result_is_independent_of_the_split(dev, xbins, zbins, gw):
    """Groups dealt out round-robin to 3 'ranks', in two contiguous blocks, one by one, and written through 'peer'
    grids (three grids of this GPU): always the bits of the single launch.  70 = 32 + 32 + 6 points per row, 112 = 3 x 32 + 16."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=0.0)
    hist, dlat, wp, _ = _problem(sc, dev, 40, 40)
    xm, zm, xa, za, slope, icpt = _mesh(sc, xbins, zbins)
    plan = ops.wake_xgroup_plan(hist, wp, xa, za)
    ng = zbins * ((xbins + gw - 1) // gw)
    assert plan.n_groups == ng and plan.group_points == gw
    n = xbins * zbins
    de, kick = ops.wake_grid_xgroups(hist, dlat, wp, xa, za, slope, icpt, plan=plan)
    splits = {"stride 3": [(r, None, 3) for r in range(3)],
              "two blocks": [(0, 5, 1), (5, ng - 5, 1)],
              "one by one": [(g, 1, 1) for g in range(ng)]}
    for label, parts in splits.items():
        out = torch.full((2, n), float("nan"), dtype=torch.float64, device=dev)
        for first, count, stride in parts:
            ops.wake_grid_xgroups(hist, dlat, wp, xa, za, slope, icpt, plan=plan, group_first=first, group_count=count,
                                  group_stride=stride, out=out)
        assert torch.equal(out[0], de) and torch.equal(out[1], kick), label
    grids = [torch.full((2, n), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    ptrs = (C.c_uint64 * 3)(*[g.data_ptr() for g in grids])
    for r in range(3):
        ops.wake_grid_xgroups(hist, dlat, wp, xa, za, slope, icpt, plan=plan, group_first=r, group_stride=3, peer_ptrs=ptrs)
    for g in grids:
        assert torch.equal(g[0], de) and torch.equal(g[1], kick)


def test_xgroup_plan_rules(dev):
    """The mapping applies only where its premise holds (the quadrature nodes of a point depend on s alone) and where the
    point kernel has no better tool; the plan is a function of the step's scalars and the whole mesh."""
    from pydfcsr_b200 import _lib, ops
    sc = scenario.chicane_entry(tilt=0.0)
    hist, dlat, wp, _ = _problem(sc, dev, 40, 40)
    _, _, xa, za, slope, icpt = _mesh(sc, 32, 4)
    assert ops.wake_xgroup_plan(hist, wp, xa, za).n_groups == 4
    _, _, xa_narrow, za_n, _, _ = _mesh(sc, 10, 30)                       # the bundled example's mesh: 10 of 32 lanes
    assert ops.wake_xgroup_plan(hist, wp, xa_narrow, za_n).n_groups == 0
    wp_on = ops.wake_params(nx=40, nz=40, skip="on", **sc["wake_scalars"])   # skipping requested: the point kernel serves it
    assert ops.wake_xgroup_plan(hist, wp_on, xa, za).n_groups == 0
    chirp = scenario.chicane_entry(tilt=2.5)                              # chirp band: x' nodes follow the point's x
    hist_c, dlat_c, wp_c, _ = _problem(chirp, dev, 40, 40)
    _, _, xa_c, za_c, _, _ = _mesh(chirp, 32, 4)
    assert abs(wp_c.slope0) > 1 and ops.wake_xgroup_plan(hist_c, wp_c, xa_c, za_c).n_groups == 0
    with pytest.raises(_lib.DfcsrError):
        ops.wake_grid_xgroups(hist_c, dlat_c, wp_c, xa_c, za_c, 0.0, 0.0, plan=ops.wake_xgroup_plan(hist, wp, xa, za))
    # larger meshes get larger units (fewer partial sums), never more than 8 nodes
    _, _, xa_big, za_big, _, _ = _mesh(sc, 64, 512)
    wp_big = ops.wake_params(nx=200, nz=200, skip="off", **sc["wake_scalars"])
    big = ops.wake_xgroup_plan(hist, wp_big, xa_big, za_big)
    assert big.group_points == 32 and big.n_groups == 1024 and 1 < big.unit_nodes <= 8 and big.max_units * big.unit_nodes >= 800


def test_xgroup_empty_quadrature_and_fp32(dev):
    """A history grid no x' node can reach gives exact zeros like the point kernel; the optional fp32 history stays
    inside its 1e-4 gate through this mapping too."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=0.0)
    st, lat = sc["stack"], sc["lattice"]
    xm, zm, xa, za, slope, icpt = _mesh(sc, 32, 3)
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, dev)
    wp = ops.wake_params(nx=30, nz=40, skip="off", **sc["wake_scalars"])
    far = ops.DeviceHistory.from_stacks([st.data[k] for k in O.FIELDS], st.min_x, st.min_y + 1.0, st.min_z,
                                        st.delta_x, st.delta_y, st.delta_z, dev)          # transverse axis 1 m away
    plan = ops.wake_xgroup_plan(far, wp, xa, za)
    assert plan.n_groups == 3
    de, kick = ops.wake_grid_xgroups(far, dlat, wp, xa, za, slope, icpt, plan=plan)
    de_p, kick_p = ops.wake_grid(far, dlat, wp, xa, za, slope, icpt)
    assert float(de.abs().max()) == 0.0 and float(kick.abs().max()) == 0.0
    assert float(de_p.abs().max()) == 0.0 and float(kick_p.abs().max()) == 0.0
    h32 = ops.DeviceHistory.from_stacks([st.data[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z,
                                        st.delta_x, st.delta_y, st.delta_z, dev, precision="fp32")
    plan = ops.wake_xgroup_plan(h32, wp, xa, za)
    de32, kick32 = ops.wake_grid_xgroups(h32, dlat, wp, xa, za, slope, icpt, plan=plan)
    osc = O.WakeScalars(nx=30, nz=40, **sc["wake_scalars"])
    ref_de, ref_kick = O.wake_mesh(xm, zm, osc, lat, st)
    assert _rel(de32.cpu().numpy(), ref_de) < 1e-4 and _rel(kick32.cpu().numpy(), ref_kick) < 1e-4


def test_csr2d_selects_the_mapping_per_step(dev):
    """CSR2D.calculate_2D_CSR uses the x-group kernel when the plan applies and the point kernel otherwise; forcing the
    point kernel gives the same grids to 1e-12."""
    import torch
    from pydfcsr_b200 import CSR2D, synth
    inp = {"input_beam": {"style": "synthetic", "n_particle": 100_000, "seed": 3},
           "input_lattice": {"lattice_config": synth.chicane_lattice_config()},
           "particle_deposition": dict(xbins=100, zbins=100, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                       velocity_threhold=1000, upper_limit=1000),
           "CSR_integration": dict(n_formation_length=1, zbins=48, xbins=40),
           "CSR_computation": dict(compute_CSR=1, apply_CSR=0, transverse_on=1, xbins=32, zbins=6, xlim=3, zlim=3,
                                   write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_test")}
    csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
    csr.skip_mode = "off"
    csr.run(stop_time=0.35)
    assert csr.last_wake_mapping == "xgroup"
    a = (csr.dE_dct.clone(), csr.x_kick.clone())
    csr.wake_mapping = "point"
    csr.calculate_2D_CSR()
    assert float((a[0] - csr.dE_dct).abs().max() / csr.dE_dct.abs().max()) < 1e-12
    assert float((a[1] - csr.x_kick).abs().max() / csr.x_kick.abs().max()) < 1e-12
    assert float(a[0].abs().max()) > 0


@pytest.mark.parametrize("T,zstep,xstep", [(1, 1, 1), (2, 40, 1), (3, 64, 50)])
def test_xgroup_small_histories(dev, T, zstep, xstep):
    """Single-slice windows (the slice pair is the same slice twice), z grids shorter than the warp's 16-node window and
    x grids of a few rows: the window clamps, the clamp cells and the extrapolation band keep the reference's rules
    (interp3D.py:30-64).  Oracle parity on a coarsened copy of the scenario's history."""
    import torch
    from pydfcsr_b200 import ops
    sc = scenario.chicane_entry(tilt=0.0)
    st, lat = sc["stack"], sc["lattice"]
    data = {k: np.ascontiguousarray(st.data[k][-T:, ::xstep, ::zstep]) for k in O.FIELDS}
    small = O.HistoryStack(data, st.min_x + (st.shape[0] - T) * st.delta_x, st.min_y, st.min_z,
                           st.delta_x, st.delta_y * xstep, st.delta_z * zstep)
    assert small.shape[0] == T and (zstep == 1 or small.shape[2] < 16)
    hist = ops.DeviceHistory.from_stacks([small.data[k] for k in O.FIELDS], small.min_x, small.min_y, small.min_z,
                                         small.delta_x, small.delta_y, small.delta_z, dev, cap=T + 2, head=1)
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, dev)
    nx, nz = 36, 40
    wp = ops.wake_params(nx=nx, nz=nz, skip="off", **sc["wake_scalars"])
    osc = O.WakeScalars(nx=nx, nz=nz, **sc["wake_scalars"])
    xm, zm, xa, za, slope, icpt = _mesh(sc, 32, 3)
    plan = ops.wake_xgroup_plan(hist, wp, xa, za)
    if plan.n_groups == 0:
        pytest.skip("the plan sends this history to the point kernel (spread criterion)")
    de, kick = ops.wake_grid_xgroups(hist, dlat, wp, xa, za, slope, icpt, plan=plan)
    ref_de, ref_kick = O.wake_mesh(xm, zm, osc, lat, small)
    assert float(np.abs(ref_de).max()) > 0
    assert _rel(de.cpu().numpy(), ref_de) < TOL and _rel(kick.cpu().numpy(), ref_kick) < TOL
