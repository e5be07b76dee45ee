"""GPU parity of the whole hot path through the reference-facing classes (DF_tracker, CSR2D, Beam):
the same particle batches go through the CUDA pipeline and through the CPU oracle pipeline."""
import os

import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from tests import scenario

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("tilt", [0.0, 2.5])
def test_df_tracker_history_matches_oracle(tilt):
    """get_DF -> append_DF -> append_interpolant -> build_interpolant over 6 logged steps."""
    import torch
    from pydfcsr_b200 import DF_tracker
    sc = scenario.chicane_entry(tilt=tilt)
    trk = DF_tracker(dict(scenario.DEPOSIT_CFG), device="cuda:0")
    for st in sc["steps"]:
        x, px, y, py, z, pz = st["coords"]
        trk.get_DF(torch.from_numpy(x).cuda(), torch.from_numpy(z).cuda(), torch.from_numpy(px).cuda(), st["pos"])
        for name in O.FIELDS:
            assert _rel(getattr(trk, name), getattr(st["df"], name)) < 1e-10, name
        assert np.array_equal(trk.x_grids.shape, st["df"].x_grids.shape)
        assert _rel(trk.x_grids, st["df"].x_grids) < 1e-13
        trk.append_DF()
        rebuilt = trk.append_interpolant(st["formation_length"], 1)
        assert rebuilt == st["rebuilt"]
    trk.build_interpolant()
    ref = sc["stack"]
    assert (trk.history.T,) + tuple(trk._ring.shape[1:3]) == ref.shape
    for name in O.FIELDS:
        assert _rel(getattr(trk, f"data_{name}_interp"), ref.data[name]) < 1e-10, name
    for a, b in ((trk.min_x, ref.min_x), (trk.min_y, ref.min_y), (trk.min_z, ref.min_z),
                 (trk.delta_x, ref.delta_x), (trk.delta_y, ref.delta_y), (trk.delta_z, ref.delta_z)):
        assert abs(a - b) <= 1e-12 * max(abs(b), 1e-30)


def test_history_window_pop_and_rebuild():
    """Sliding window (deposit.py:265-280) and the sigma-ratio rebuild (deposit.py:321-361): a bunch
    whose length shrinks by >2x forces a rebuild with a finer grid; a short window pops old slices."""
    import torch
    from pydfcsr_b200 import DF_tracker, synth
    cfg = dict(scenario.DEPOSIT_CFG)
    cfg["upper_limit"] = 700
    trk = DF_tracker(cfg, device="cuda:0")
    hist = O.HistoryOracle(O.DepositConfig(**cfg))
    b = synth.gaussian_bunch(60_000, seed=4)
    scale = [1.0, 0.9, 0.7, 0.45, 0.4, 0.38, 0.36, 0.35]
    for k, sc_ in enumerate(scale):
        x, px, z = b[0], b[1], b[4] * sc_
        t = 0.1 * k
        fl = float("inf") if k == 0 else 0.25
        trk.get_DF(torch.from_numpy(x).cuda(), torch.from_numpy(np.ascontiguousarray(z)).cuda(), torch.from_numpy(px).cuda(), t)
        trk.append_DF()
        got = trk.append_interpolant(fl, 1)
        hist.append(O.make_density_functions(x, z, px, t, hist.cfg))
        exp = hist.push(fl, 1)
        assert got == exp, k
        assert list(trk.time_interp) == list(hist.time_interp), k
    assert trk.rebuilds == hist.rebuilds >= 2
    trk.build_interpolant()
    ref = hist.stack()
    assert (trk.history.T,) + tuple(trk._ring.shape[1:3]) == ref.shape
    assert ref.shape[0] < len(scale) and ref.shape[2] <= 700
    for name in O.FIELDS:
        assert _rel(getattr(trk, f"data_{name}_interp"), ref.data[name]) < 1e-10, name


def test_csr2d_step_matches_oracle():
    """CSR2D end to end on the device (tracking stand-in, deposit, history, mesh, wake, kick) against
    the oracle fed with the device's particle batches."""
    import torch
    from pydfcsr_b200 import CSR2D, synth
    elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in synth.CHICANE_ELEMENTS]
    inp = {"input_beam": {"style": "synthetic", "n_particle": 100_000, "seed": 1},
           "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
           "particle_deposition": dict(scenario.DEPOSIT_CFG),
           "CSR_integration": dict(n_formation_length=1, zbins=40, xbins=40),
           "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, xbins=4, zbins=6, xlim=3, zlim=3,
                                   write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_test")}
    csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
    cfg = O.DepositConfig(**scenario.DEPOSIT_CFG)
    hist = O.HistoryOracle(cfg)
    lat = scenario.lattice_tables()
    assert np.array_equal(lat.coords, csr.lattice.coords) and np.array_equal(lat.tau_vec, csr.lattice.tau_vec)

    def oracle_log(fl):
        c = csr.beam.to_host()
        hist.append(O.make_density_functions(c[0], c[4], c[1], csr.beam.position, cfg))
        hist.push(fl, 1)
        return c

    oracle_log(float("inf"))
    # drive the same sequence as CSR2D.run but interleave the oracle after every tracking step
    from pydfcsr_b200 import tracking
    steps = [(tracking.Drift(0.1), 0.1)] + [(tracking.SBend(L=0.1, G=0.0483 / 0.5002, FRINGE_AT="no_end"), None)] * 3
    for el, fl in steps:
        if fl is None and csr.formation_length in (None, 0.1):
            csr.get_formation_length(R=0.5002 / 0.0483, sigma_z=5 * csr.beam.sigma_z)
        elif fl is not None:
            csr.formation_length = fl
        csr.beam.track(el, 0.1)
        c = oracle_log(csr.formation_length)
        pre_px, pre_pz = c[1].copy(), c[5].copy()
        csr.hot_path_step(apply=True, kick_length=0.1)
        s = O.beam_scalars(c[0], c[4])
        xm, zm, xr, zr = O.observation_mesh(c[0], c[4], s["slope"], s["sigma_z"], s["mean_z"], 3, 3, 4, 6)
        assert _rel(csr.CSR_xmesh, xm) < 1e-9 and _rel(csr.CSR_zmesh, zm) < 1e-12
        sc = O.WakeScalars(t=csr.beam.position, sigma_x=float(s["sigma_x"]), sigma_z=float(s["sigma_z"]),
                           slope0=float(s["slope"][0]), mean_x=float(s["mean_x"]), formation_window=csr.formation_length,
                           csr_scaling=8.98755e3 * 1e-9, nx=40, nz=40)
        de, kick = O.wake_mesh(xm, zm, sc, lat, hist.stack())
        assert _rel(csr.dE_dct.cpu().numpy().ravel(), de) < 1e-10
        assert _rel(csr.x_kick.cpu().numpy().ravel(), kick) < 1e-10
        px_new, pz_new = O.apply_kick(c[0], c[4], pre_px, pre_pz, de.reshape(4, 6), kick.reshape(4, 6), xr, zr, 0.1, 5e9, True)
        after = csr.beam.to_host()
        assert _rel(after[5] - pre_pz, pz_new - pre_pz) < 1e-9
        assert _rel(after[1] - pre_px, px_new - pre_px) < 1e-9
    # single-point entry and the debug hook agree with the mesh launch
    k = 9
    s_k = csr.beam.position + csr.CSR_zmesh[k]
    de_k, kick_k = csr.get_CSR_wake(s_k, csr.CSR_xmesh[k])
    assert abs(de_k - float(csr.dE_dct.ravel()[k])) <= 1e-12 * abs(de_k)
    dbg = csr.get_CSR_wake(s_k, csr.CSR_xmesh[k], debug=True)
    tot = sum(np.trapz(np.trapz(r["integrand_z"], r["xp"], axis=0), r["sp"]) for r in dbg)
    assert abs(-csr.CSR_scaling * tot - de_k) <= 1e-10 * abs(de_k)


def test_run_is_bit_reproducible():
    """Fixed-point deposit, ticketed reductions and fixed-order quadrature: two runs of the same input give
    the same bits for the wake grids and for the kicked particles (the reference's serial numba path is
    reproducible as well; fp64 atomics would not be)."""
    import torch
    from pydfcsr_b200 import CSR2D, synth
    elements = [(n, k, L, a, e1, e2, 2) for (n, k, L, a, e1, e2, _s) in synth.CHICANE_ELEMENTS[:3]]

    def run():
        inp = {"input_beam": {"style": "synthetic", "n_particle": 150_000, "seed": 4},
               "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
               "particle_deposition": dict(scenario.DEPOSIT_CFG),
               "CSR_integration": dict(n_formation_length=1, zbins=40, xbins=40),
               "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, xbins=5, zbins=7, xlim=3, zlim=3,
                                       write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_test")}
        csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
        csr.run()
        return csr.dE_dct.clone(), csr.x_kick.clone(), torch.stack(csr.beam.coords)

    a, b = run(), run()
    assert float(a[0].abs().max()) > 0
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_zero_density_skipping_is_bitwise_through_the_whole_chicane():
    """All 133 steps of the chicane (both quadrature branches, |slope| up to ~100, rebuilds, reversed rectangles early
    on) with the wake applied at every evaluation: the run with zero-density skipping (row hulls + coarse s' bracket,
    DESIGN.md section 4 (vi)) and the run without it (skip_mode OFF) must end with the SAME BITS in the last wake
    grids and in every particle coordinate -- any sample dropped by mistake would have kicked the beam differently."""
    import torch
    from pydfcsr_b200 import CSR2D, synth

    def run(skip):
        inp = {"input_beam": {"style": "synthetic", "n_particle": 200_000, "seed": 7},
               "input_lattice": {"lattice_config": synth.chicane_lattice_config()},
               "particle_deposition": dict(xbins=120, zbins=120, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                           velocity_threhold=1000, upper_limit=1000),
               "CSR_integration": dict(n_formation_length=1, zbins=70, xbins=60),
               "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, xbins=6, zbins=9, xlim=3, zlim=3,
                                       write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_test")}
        csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
        csr.wake_counters = torch.zeros(3, dtype=torch.int64, device="cuda:0")
        csr.skip_mode = skip
        csr.run()
        return csr.dE_dct.clone(), csr.x_kick.clone(), torch.stack(csr.beam.coords), [int(v) for v in csr.wake_counters.cpu()]

    on, off = run("auto"), run("off")
    assert float(on[0].abs().max()) > 0
    for u, v in zip(on[:3], off[:3]):
        assert torch.equal(u, v)
    # over the run, most in-grid samples were never gathered (the chirped bunch is a thin band of its grid)
    assert off[3][2] == off[3][0] and on[3][2] < 0.7 * off[3][0], (on[3], off[3])


def test_csr2d_full_chicane_shadowed_by_oracle():
    """The whole 133-step chicane through CSR2D.run() on the device, shadowed step by step by the CPU
    oracle fed with the device's particle batches: same grid-branch / window / rebuild decisions, and
    every wake (both quadrature branches: |slope| climbs to ~100 in the middle of the chicane, then
    falls back) within the 1e-10 gate.  Reduced sizes keep the CPU side to about a minute."""
    import torch
    from pydfcsr_b200 import CSR2D, synth
    dep = dict(xbins=48, zbins=64, xlim=5, zlim=5, filter_order=1, filter_window=9, velocity_threhold=1000, upper_limit=640)
    inp = {"input_beam": {"style": "synthetic", "n_particle": 60_000, "seed": 2},
           "input_lattice": {"lattice_config": synth.chicane_lattice_config()},
           "particle_deposition": dep,
           "CSR_integration": dict(n_formation_length=1, zbins=24, xbins=24),
           "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, xbins=3, zbins=4, xlim=3, zlim=3,
                                   write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_test")}
    csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
    trk = csr.DF_tracker
    cfg = O.DepositConfig(**dep)
    hist = O.HistoryOracle(cfg)
    lat = scenario.lattice_tables()
    log = dict(wakes=0, band=0, rebuilds=0, max_err=0.0, shapes=set())

    # the constructor has already logged the initial beam (CSR.py:75-78): mirror it
    c0 = csr.beam.to_host()
    hist.append(O.make_density_functions(c0[0], c0[4], c0[1], 0, cfg))
    hist.push(float("inf"), 1)

    get_df, push, wake = trk.get_DF, trk.append_interpolant, csr.calculate_2D_CSR

    def shadow_get_df(x, z, px, t, stats=None):
        get_df(x=x, z=z, px=px, t=t, stats=stats)
        df = O.make_density_functions(x.cpu().numpy(), z.cpu().numpy(), px.cpu().numpy(), t, cfg)
        assert df.density.shape == tuple(trk._current.fields.shape[1:]), "grid branch (deposit.py:160-167) differs"
        log["shapes"].add(df.density.shape)
        hist.append(df)

    def shadow_push(formation_length, n_formation_length):
        got = push(formation_length, n_formation_length)
        exp = hist.push(formation_length, n_formation_length)
        assert got == exp and list(trk.time_interp) == list(hist.time_interp), "window / rebuild policy differs"
        log["rebuilds"] += int(exp)
        return got

    def shadow_wake():
        wake()
        b = csr.beam
        sc = O.WakeScalars(t=b.position, sigma_x=b._sigma_x, sigma_z=b._sigma_z, slope0=b._slope[0], mean_x=b._mean_x,
                           formation_window=csr.integration_params.n_formation_length * csr.formation_length,
                           csr_scaling=csr.CSR_scaling, nx=24, nz=24)
        de, kick = O.wake_mesh(csr.CSR_xmesh, csr.CSR_zmesh, sc, lat, hist.stack())
        g_de, g_kick = csr.dE_dct.cpu().numpy().ravel(), csr.x_kick.cpu().numpy().ravel()
        assert np.all(np.isfinite(g_de)) and np.all(np.isfinite(g_kick))
        err = max(_rel(g_de, de), _rel(g_kick, kick))
        log["max_err"] = max(log["max_err"], err)
        log["wakes"] += 1
        log["band"] += int(abs(b._slope[0]) > 1)
        assert err < 1e-10, (b.position, b._slope[0], err)

    trk.get_DF, trk.append_interpolant, csr.calculate_2D_CSR = shadow_get_df, shadow_push, shadow_wake
    csr.run()
    assert csr.beam.step == 133 and abs(csr.beam.position - 13.3) < 1e-9
    assert log["wakes"] >= 50 and log["band"] >= 10 and log["wakes"] - log["band"] >= 5, log
    assert log["max_err"] < 5e-11, log        # measured 7.6e-12 over the whole lattice
    assert log["rebuilds"] >= 2 and len(log["shapes"]) == 2, log          # both deposit-grid branches were taken
    assert csr.beam.sigma_z < 0.2 * 200e-6                                  # the chicane compressed the bunch ~10x


def test_yaml_driven_run_matches_reference_schema(tmp_path, monkeypatch):
    """CSR2D(input_file=...) with the reference's YAML schema (examples/input/chicane_config.yaml, relative
    lattice path resolved against the working directory like the reference does), a few steps."""
    import torch
    from pydfcsr_b200 import CSR2D
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.chdir(os.path.join(root, "examples"))
    csr = CSR2D(input_file="input/chicane_config.yaml", verbose=False)
    assert csr.CSR_params.xbins == 10 and csr.CSR_params.zbins == 30 and csr.integration_params.xbins == 200
    assert csr.DF_tracker.xbins == 300 and csr.DF_tracker.upper_limit == 2000 and csr.lattice.total_steps == 134
    csr.run(stop_time=0.25)
    assert csr.beam.step == 3 and tuple(csr.dE_dct.shape) == (10, 30)
    assert bool(torch.isfinite(csr.dE_dct).all()) and float(csr.dE_dct.abs().max()) > 0
    with pytest.raises(AssertionError):
        CSR2D(input_file={"input_beam": {"style": "synthetic", "n_particle": 10}, "input_lattice": {}, "bogus": 1})


def test_prefetched_density_functions_are_bitwise_the_synchronous_ones():
    """DF_tracker.prefetch_DF enqueues deposit + density functions behind the statistics pass with the grid limits and
    max|px| read from the device (dfcsr_get_df_from_stats); get_DF adopts the result when its guess of the grid shape was
    right.  Fields, scalars and the record's axes must be the bits of the synchronous path (deposit.py:145-245), for both
    grid branches (deposit.py:157-167), and a wrong guess must fall back to the synchronous path."""
    import torch
    from pydfcsr_b200 import CSR2D, synth

    def make(prefetch):
        inp = {"input_beam": {"style": "synthetic", "n_particle": 200_000, "seed": 11},
               "input_lattice": {"lattice_config": synth.chicane_lattice_config()},
               "particle_deposition": dict(xbins=120, zbins=90, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                           velocity_threhold=1000, upper_limit=1000),
               "CSR_integration": dict(n_formation_length=1, zbins=40, xbins=40),
               "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, xbins=6, zbins=9, xlim=3, zlim=3,
                                       write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_test")}
        csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
        csr.DF_tracker.prefetch = prefetch
        return csr

    a, b = make(True), make(False)
    a.run(stop_time=2.6)
    b.run(stop_time=2.6)
    ta, tb = a.DF_tracker, b.DF_tracker
    assert ta.prefetch_hits > 10 and tb.prefetch_hits == 0
    ra, rb = ta._current, tb._current
    assert torch.equal(ra.fields, rb.fields) and torch.equal(ra.scalars, rb.scalars)
    assert (ra.x_axis.start, ra.x_axis.stop, ra.z_axis.start, ra.z_axis.stop) == (rb.x_axis.start, rb.x_axis.stop,
                                                                                 rb.z_axis.start, rb.z_axis.stop)
    assert torch.equal(a.dE_dct, b.dE_dct) and torch.equal(a.x_kick, b.x_kick)
    for u, v in zip(a.beam.coords, b.beam.coords):
        assert torch.equal(u, v)
    # the limits the device computed are the host's (same roundings)
    lim = ta._limits.cpu().numpy()
    assert (lim[0], lim[1], lim[2], lim[3]) == (ra.x_axis.start, ra.x_axis.stop, ra.z_axis.start, ra.z_axis.stop)
    # a wrong guess of the grid shape is discarded
    hits = ta.prefetch_hits
    ta._spec_shape = (64, 64, 5)
    a.beam.update_status()
    ta.prefetch_DF(a.beam)
    ta.get_DF(x=a.beam.x, z=a.beam.z, px=a.beam.px, t=a.beam.position, stats=a.beam.stats)
    tb.get_DF(x=b.beam.x, z=b.beam.z, px=b.beam.px, t=b.beam.position, stats=b.beam.stats)
    assert ta.prefetch_hits == hits and torch.equal(ta._current.fields, tb._current.fields)
