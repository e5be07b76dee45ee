"""Run the bundled-style chicane example end to end on one GPU and print per-step timings.

    cd examples && python run_chicane.py [config.yaml] [stop_position_m]

The reference's notebook (example/example_chicane.ipynb) reports 133 steps in 604.8 s with 8.9-10.3 s per
CSR step on one CPU core for this configuration (1e6 particles, 10 x 30 mesh, 200 x 200 integration)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydfcsr_b200 import CSR2D  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "input/chicane_config.yaml"
stop = float(sys.argv[2]) if len(sys.argv) > 2 else None
os.chdir(os.path.dirname(os.path.abspath(__file__)))
t0 = time.perf_counter()
csr = CSR2D(input_file=cfg, verbose=False)
torch.cuda.synchronize()
t1 = time.perf_counter()
csr.run(stop_time=stop)
torch.cuda.synchronize()
t2 = time.perf_counter()
st = csr.statistics
n = csr.beam.step
print(f"set-up {t1 - t0:.2f} s; {n} steps to s = {csr.beam.position:.3f} m in {t2 - t1:.2f} s "
      f"({(t2 - t1) / max(n, 1) * 1e3:.2f} ms per step); history rebuilds {csr.DF_tracker.rebuilds}, "
      f"ring {tuple(csr.DF_tracker._ring.shape)}")
print(f"sigma_z {st['sigma_z'][0] * 1e6:.1f} -> {csr.beam.sigma_z * 1e6:.1f} um, sigma_x {st['sigma_x'][0] * 1e6:.1f} -> "
      f"{csr.beam.sigma_x * 1e6:.1f} um, max |slope| {np.max(np.abs(st['slope'][:n + 1, 0])):.2f}")
print(f"last wake: max |dE/dct| {float(csr.dE_dct.abs().max()):.4e} MeV/m, max |x kick| {float(csr.x_kick.abs().max()):.4e} MeV/m, "
      f"finite {bool(torch.isfinite(csr.dE_dct).all() and torch.isfinite(csr.x_kick).all())}")
