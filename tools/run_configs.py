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

from tests.configs import CONFIGS, build, oracle_points  # noqa: E402,F401


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
        name = f"{name}[{getattr(csr, 'last_wake_mapping', 'point')[0]}]"          # [x] = x-group mapping, [p] = point kernel
        print(f"{name:14s} {beam.x.numel():10d} {csr.CSR_params.xbins:4d}x{csr.CSR_params.zbins:<4d} {str(shape):>18s} "
              f"{step_ms:9.3f} {k4_ms:9.3f} {rate:>10s} {in_grid:>8s} {e_de:11.2e} {e_k:11.2e}   (setup run {setup:.1f} s)",
              flush=True)
        del csr


if __name__ == "__main__":
    main(sys.argv[1:] or list(CONFIGS))
