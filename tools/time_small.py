"""Developer tool: CUDA-event timings of K1/K2/K3/K5/A14 at the BASELINE particle counts / grid sizes, with the
algorithmic bytes and the fraction of the measured HBM peak.  python tools/time_small.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydfcsr_b200 import ops  # noqa: E402
from pydfcsr_b200._lib import Axis  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:46s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {100 * gbs / PEAK:5.1f}% of measured HBM peak")


g = torch.Generator(device="cuda").manual_seed(1)
for n in (1_000_000, 10_000_000, 50_000_000):
    x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 6e-5
    z = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 2e-4
    px = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 4e-6
    pz = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 1e-3
    for shape in ((100, 100), (300, 300), (64, 512)):
        args = (shape[0], -3e-4, 3e-4, shape[1], -1e-3, 1e-3)
        out = torch.empty((2,) + shape, dtype=torch.float64, device="cuda")
        for mode in (0, 4, 5, 3, 2):     # 0 = the default (4 for n >= 65536); 4/5 = fixed point tile / direct; 3/2 = fp64 tile / direct
            if mode == 5 and n > 1_000_000:
                continue                  # 64-bit L2 reductions on hot cells: only of interest for small bunches
            ms = timeit(lambda: ops.deposit_cic(x, z, px, *args, mode=mode, out=out))
            report(f"K1 cic n={n:.0e} grid={shape} mode={mode} (absmax + deposit + finish)", ms, 24 * n)
        # the two-stage path the tracker uses (max|px| comes with the statistics): deposit_q + finish
        import ctypes as C
        q = torch.empty(2 * shape[0] * shape[1], dtype=torch.int64, device="cuda")
        amax = float(px.abs().max())
        ptr = (C.c_uint64 * 1)(q.data_ptr())
        ms = timeit(lambda: (ops.deposit_cic_q(x, z, px, n, *args, amax, q), ops.deposit_cic_finish(ptr, n, shape[0], shape[1], amax, out=out)))
        report(f"K1 cic n={n:.0e} grid={shape} two-stage q + finish", ms, 24 * n)
    ms = timeit(lambda: ops.deposit_ngp(x, z, 100, -3e-4, 3e-4, 100, -1e-3, 1e-3))
    report(f"K1 ngp n={n:.0e} grid=(100, 100)", ms, 16 * n)
    st = torch.zeros(16, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: ops.beam_stats(x, z, pz, px))
    report(f"A14 stats (2 passes incl. max|px| + D2H sync) n={n:.0e}", ms, (32 + 16) * n)
    de = torch.randn((64, 64), generator=g, device="cuda", dtype=torch.float64)
    ms = timeit(lambda: ops.apply_kick(x, z, px, pz, 0.1, 0.0, de, de, Axis.make(-2e-4, 2e-4, 64), Axis.make(-6e-4, 6e-4, 64), 0.1, 5e9, True))
    report(f"K5 kick n={n:.0e}", ms, 48 * n)
    del x, z, px, pz

for src, dst in (((100, 100), (500, 500)), ((300, 300), (2000, 503)), ((300, 300), (2000, 2000))):
    f = torch.randn((5,) + src, generator=g, device="cuda", dtype=torch.float64)
    out = torch.empty(dst + (6,), dtype=torch.float64, device="cuda")
    sup = torch.empty((dst[0], 2), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: ops.history_regrid(f, Axis.make(-3e-4, 3e-4, src[0]), Axis.make(-1e-3, 1e-3, src[1]),
                                           Axis.make(-3.2e-4, 3.1e-4, dst[0]), Axis.make(-1.1e-3, 1e-3, dst[1]), 0.5, out, sup))
    report(f"K3 regrid + row support {src}->{dst}", ms, dst[0] * dst[1] * 48 + 5 * src[0] * src[1] * 8)

for shape, win in (((100, 100), 5), ((300, 300), 9), ((64, 512), 9)):
    cnt = torch.rand(shape, generator=g, device="cuda", dtype=torch.float64) * 100
    vxs = torch.randn(shape, generator=g, device="cuda", dtype=torch.float64)
    ms = timeit(lambda: ops.make_df(cnt, vxs, Axis.make(-3e-4, 3e-4, shape[0]), Axis.make(-1e-3, 1e-3, shape[1]), win, 1, 1000))
    report(f"K2 make_df grid={shape} window={win}", ms, shape[0] * shape[1] * 8 * 7)
