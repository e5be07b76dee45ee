"""Developer tool: time K4 kernel variants (DFCSR_WAKE_CFG, read per launch by the DEVELOPER build of the library,
tools/build_dev.py) on the bench workload in ONE process and compare each variant's wake grids with the first one.

    python tools/build_dev.py && DFCSR_LIB=pydfcsr_b200/libdfcsr_b200_dev.so python tools/k4_variants.py [reps] [cfg ...]
    DFCSR_TILT=2.5 adds an x-z tilt (chirp-band quadrature); cfg 0 = shipped kernels.
    Point mapping (DFCSR_WAKE_MAPPING=point or a chirp-band / sparse step): 20 / 40 = without bracket and queue interleave, 30 = register
    cache, 25 = two x' nodes per lane, 70 = patch kernel (71.. = patch of 0, 8, 16, ... nodes per warp), 80 = lean node records, 90 = prefetch.
    x-group mapping (the bench workload's default): 1 = unpipelined sweep, 2 = 192 x 3 CTAs, 3 / 4 = two steps per iteration, 5 = 320 x 2,
    6 = 384 x 2, 7 / 8 / 10 = two observation points per lane (256 x 1, 128 x 3, 192 x 2), 9 = plan without the spread criterion,
    11..19 = 1, 2, 3, 4, 7, 10, 13, 16, 19 CTAs per group
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydfcsr_b200 import CSR2D  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfgs = [int(a) for a in sys.argv[2:]] or [0, 20, 40, 30, 25, 70, 73, 80]
wl = bench.WORKLOAD
inp = bench._input_dict(wl)
tilt = os.environ.get("DFCSR_TILT")
if tilt:
    inp["input_beam"]["tilt"] = float(tilt)
csr = CSR2D(inp, parallel=False, verbose=False, precision=os.environ.get("DFCSR_PRECISION", "fp64"))
csr.run(stop_time=wl["position"] - 0.05)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=csr.device)
print(f"# K4 variants on the bench workload, tilt={tilt or 0}, precision={os.environ.get('DFCSR_PRECISION', 'fp64')}, "
      f"history {tuple(csr.DF_tracker.data_shape) if hasattr(csr.DF_tracker, 'data_shape') else ''}, "
      f"slope {float(csr.beam._slope[0]):.3f}")
ref = None
for cfg in cfgs:
    os.environ["DFCSR_WAKE_CFG"] = str(cfg)
    for _ in range(3):
        csr.calculate_2D_CSR()
    dE, kick = csr.dE_dct.clone(), csr.x_kick.clone()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        csr.calculate_2D_CSR()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    rep = bool(torch.equal(dE, csr.dE_dct) and torch.equal(kick, csr.x_kick))
    if ref is None:
        ref = (dE, kick)
    e1 = float((dE - ref[0]).abs().max() / ref[0].abs().max())
    e2 = float((kick - ref[1]).abs().max() / ref[1].abs().max())
    print(f"cfg {cfg:3d}: median {np.median(ts):7.3f} ms  min {np.min(ts):7.3f} ms  vs cfg {cfgs[0]}: dE {e1:.2e} kick {e2:.2e}  "
          f"bitwise_repeatable {rep}", flush=True)
