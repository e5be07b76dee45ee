"""Developer tool: time the wake kernel (K4) alone on the bench workload.  python tools/time_k4.py [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydfcsr_b200 import CSR2D  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = bench.WORKLOAD
inp = bench._input_dict(wl)
if os.environ.get('DFCSR_TILT'):
    inp['input_beam']['tilt'] = float(os.environ['DFCSR_TILT'])
csr = CSR2D(inp, parallel=False, verbose=False, precision=os.environ.get('DFCSR_PRECISION', 'fp64'))
csr.run(stop_time=wl["position"] - 0.05)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=csr.device)
for _ in range(3):
    csr.calculate_2D_CSR()
ref = csr.dE_dct.clone()
ts = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); csr.calculate_2D_CSR(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"K4 {os.environ.get('DFCSR_TAG', '')}: median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms  "
      f"checksum {float(ref.abs().sum()):.15e} bitwise_repeatable {bool(torch.equal(ref, csr.dE_dct))}")
