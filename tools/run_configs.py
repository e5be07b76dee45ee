"""Developer tool: every BASELINE.json configuration end to end on ONE B200 — CSR2D.run() to build the history,
then the hot path (deposit -> history -> mesh -> wake -> kick) timed with CUDA events, and the wake of a few
mesh points recomputed by the CPU oracle from the device's own history (K4 parity at full problem size).

    python tools/run_configs.py [names...]        names: example chicane_1e6 lcls_bc arc microbunched

configs[2..4] are quoted on 2-8 GPUs in BASELINE.json; the mesh of one rank's launch is what changes with the
rank count (reference split rule), so the single-GPU line below is the N = 1 point of those configurations.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dfcsr_oracle as O  # noqa: E402  (developer tool: the oracle is the checker here)
from pydfcsr_b200 import CSR2D, synth  # noqa: E402

DEPOSIT = dict(xbins=300, zbins=300, xlim=5, zlim=5, filter_order=1, filter_window=9, velocity_threhold=1000,
               upper_limit=2000)
ANGLE = 0.0483
ARC = [("D0", "drift", 0.1, 0.0, 0.0, 0.0, 1)]
for k in range(8):                                   # 8 dipoles of the chicane's strength, same bending direction
    ARC += [(f"B{k + 1}", "dipole", 0.5, ANGLE, 0.0, 0.0, 1), (f"DD{k + 1}", "drift", 0.25, 0.0, 0.0, 0.0, 1)]

CONFIGS = {
    # name: (beam kwargs, elements, step size, stop position, mesh (xbins, zbins), integration, deposit overrides)
    "example": (dict(n_particle=100_000, seed=0), None, 0.1, 0.6, (10, 30), dict(n_formation_length=1, xbins=200, zbins=200), {}),
    "chicane_1e6": (dict(n_particle=1_000_000, seed=0), None, 0.1, 0.6, (64, 64), dict(n_formation_length=1, xbins=200, zbins=200), {}),
    "lcls_bc": (dict(n_particle=10_000_000, seed=0, sigma_z=20.0e-6, chirp=-360.0), None, 0.1, 0.6, (128, 128),
                dict(n_formation_length=1, xbins=200, zbins=200), {}),
    "arc": (dict(n_particle=1_000_000, seed=0), ARC, 0.05, 5.5, (256, 128), dict(n_formation_length=4, xbins=200, zbins=200), {}),
    "microbunched": (dict(n_particle=50_000_000, seed=0, tilt=2.5, modulation=0.1, modulation_wavelength_sigma=0.05), None, 0.1, 0.4,
                     (64, 512), dict(n_formation_length=1, xbins=200, zbins=200), dict(xbins=64, zbins=512, filter_order=2)),
}


def build(name):
    beam, elements, step, stop, mesh, integ, dep = CONFIGS[name]
    elements = elements or synth.CHICANE_ELEMENTS
    elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in elements]
    lat = synth.chicane_lattice_config(elements=elements)
    lat["step_size"] = step
    inp = {"input_beam": dict(style="synthetic", **beam),
           "input_lattice": {"lattice_config": lat},
           "particle_deposition": dict(DEPOSIT, **dep),
           "CSR_integration": integ,
           "CSR_computation": dict(compute_CSR=1, apply_CSR=0, transverse_on=1, write_beam=None, write_wakes=False,
                                   workdir="/tmp/dfcsr_cfg", xbins=mesh[0], zbins=mesh[1], xlim=5, zlim=5)}
    return CSR2D(inp, parallel=False, verbose=False, precision=os.environ.get("DFCSR_PRECISION", "fp64")), stop


def oracle_points(csr, picks):
    trk = csr.DF_tracker
    data = {f: np.ascontiguousarray(getattr(trk, f"data_{f}_interp")) for f in O.FIELDS}
    stack = O.HistoryStack(data, float(trk.min_x), float(trk.min_y), float(trk.min_z), float(trk.delta_x),
                           float(trk.delta_y), float(trk.delta_z))
    lat = O.LatticeTables(coords=csr.lattice.coords, n_vec=csr.lattice.n_vec, tau_vec=csr.lattice.tau_vec,
                          rho=np.asarray(csr.lattice.rho, dtype=np.float64), distance=np.asarray(csr.lattice.distance, dtype=np.float64),
                          min_s=float(csr.lattice.min_x), delta_s=float(csr.lattice.delta_x))
    b = csr.beam
    ip = csr.integration_params
    sc = O.WakeScalars(t=b.position, sigma_x=b._sigma_x, sigma_z=b._sigma_z, slope0=float(b._slope[0]), mean_x=b._mean_x,
                       formation_window=ip.n_formation_length * csr.formation_length, csr_scaling=csr.CSR_scaling,
                       nx=ip.xbins, nz=ip.zbins)
    xm, zm = np.asarray(csr.CSR_xmesh), np.asarray(csr.CSR_zmesh)
    return np.array([O.wake_point(b.position + zm[k], xm[k], sc, lat, stack) for k in picks])


def main(names):
    print(f"{'config':14s} {'particles':>10s} {'mesh':>9s} {'history (T,X,Z)':>18s} {'step ms':>9s} {'K4 ms':>9s} "
          f"{'samples/s':>10s} {'in-grid':>8s} {'K4 parity (dE, kick)':>24s}")
    for name in names:
        torch.cuda.empty_cache()
        csr, stop = build(name)
        t0 = time.perf_counter()
        csr.run(stop_time=stop - 1e-9)
        setup = time.perf_counter() - t0
        trk, beam = csr.DF_tracker, csr.beam
        if csr.dE_dct is None or len(trk.time_interp) < 2:
            print(f"{name}: no wake computed up to s = {stop}")
            continue

        def step():
            trk.pop_right_interpolant()
            csr.hot_path_step(apply=False)

        for _ in range(2):
            step()
        reps = 5
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for _ in range(reps):
            step()
        ev[1].record()
        torch.cuda.synchronize()
        csr.get_CSR_mesh()
        csr.wake_counters = torch.zeros(3, dtype=torch.int64, device=beam.x.device)
        ev[2].record()
        for _ in range(reps):
            csr.calculate_2D_CSR()
        ev[3].record()
        torch.cuda.synchronize()
        step_ms, k4_ms = ev[0].elapsed_time(ev[1]) / reps, ev[2].elapsed_time(ev[3]) / reps
        n_mesh = csr.CSR_params.xbins * csr.CSR_params.zbins
        counters = [int(v) // reps for v in csr.wake_counters.cpu()]      # [in-grid samples, samples the reference evaluates]
        csr.wake_counters = None
        in_grid = f"{counters[0] / counters[1]:.3f}/{counters[2] / counters[1]:.3f}" if counters[1] else "-"
        rate = f"{counters[1] / (k4_ms * 1e-3):.3e}" if counters[1] else "-"
        de, kick = csr.dE_dct.cpu().numpy().ravel(), csr.x_kick.cpu().numpy().ravel()
        rng = np.random.default_rng(0)
        picks = np.concatenate([[int(np.argmax(np.abs(de)))], rng.choice(n_mesh, 5, replace=False)])
        ref = oracle_points(csr, picks)
        e_de = np.max(np.abs(de[picks] - ref[:, 0])) / np.max(np.abs(de))
        e_k = np.max(np.abs(kick[picks] - ref[:, 1])) / np.max(np.abs(kick))
        T = len(trk.time_interp)
        shape = (T, len(trk.x_grid_interp), len(trk.z_grid_interp))
        print(f"{name:14s} {beam.x.numel():10d} {csr.CSR_params.xbins:4d}x{csr.CSR_params.zbins:<4d} {str(shape):>18s} "
              f"{step_ms:9.3f} {k4_ms:9.3f} {rate:>10s} {in_grid:>8s} {e_de:11.2e} {e_k:11.2e}   (setup run {setup:.1f} s)",
              flush=True)
        del csr


if __name__ == "__main__":
    main(sys.argv[1:] or list(CONFIGS))
