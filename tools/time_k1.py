"""CUDA-event timing of the CIC deposit alone (K1).  DFCSR_NCU=1: one launch per case, for an ncu capture.
usage: python tools/time_k1.py [n_particles]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydfcsr_b200 import ops  # noqa: E402

PEAK = 6550.0
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
ncu = os.environ.get("DFCSR_NCU") == "1"
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 6e-5
z = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 2e-4
px = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 4e-6
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for shape in ((100, 100), (300, 300), (64, 512)):
    args = (shape[0], -3e-4, 3e-4, shape[1], -1e-3, 1e-3)
    out = torch.empty((2,) + shape, dtype=torch.float64, device="cuda")
    for mode in (4, 3):
        if ncu:
            ops.deposit_cic(x, z, px, *args, mode=mode, out=out)
            continue
        for _ in range(3):
            ops.deposit_cic(x, z, px, *args, mode=mode, out=out)
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.deposit_cic(x, z, px, *args, mode=mode, out=out)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        gbs = 24 * n / (ms * 1e-3) / 1e9
        print(f"K1 cic n={n:.0e} grid={shape} mode={mode}: {ms * 1e3:8.1f} us  {gbs:7.1f} GB/s  {100 * gbs / PEAK:5.1f}% of HBM peak")
