"""Developer experiment: does the order in which observation points are handed to CTAs matter for K4?
Times dfcsr_wake_mesh on the bench workload with the point arrays permuted (results are un-permuted and compared).
python tools/k4_order_experiment.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydfcsr_b200 import CSR2D, ops  # noqa: E402

wl = bench.WORKLOAD
csr = CSR2D(bench._input_dict(wl), parallel=False, verbose=False)
csr.run(stop_time=wl["position"] - 0.05)
csr.get_CSR_mesh()
dev = csr.device
xm = torch.from_numpy(np.ascontiguousarray(csr.CSR_xmesh)).to(dev)
zm = torch.from_numpy(np.ascontiguousarray(csr.CSR_zmesh)).to(dev)
Mx, Mz = csr.CSR_params.xbins, csr.CSR_params.zbins
N = Mx * Mz
lat = csr.lattice.device_tables(dev)
wp = csr._wake_params()
hist = csr.DF_tracker.history
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
idx = np.arange(N).reshape(Mx, Mz)                 # mesh index = ix * Mz + iz


def pair_order(sorted_pts, sms=148, per_sm=2):
    """Blocks b and b + 148 (co-resident in the first wave) get consecutive entries of `sorted_pts`."""
    wave = sms * per_sm
    out = np.empty_like(sorted_pts)
    for b in range(len(sorted_pts)):
        w, s = divmod(b, wave)
        sm, half = s % sms, s // sms
        src = w * wave + sm * per_sm + half
        out[b] = sorted_pts[src] if src < len(sorted_pts) else -1
    if (out < 0).any() or len(set(out.tolist())) != len(out):      # ragged last wave: fall back to plain order there
        tail = (len(sorted_pts) // wave) * wave
        out[tail:] = sorted_pts[tail:]
    return out


orders = {
    "z fastest (shipped)": idx.ravel(),
    "x fastest": idx.T.ravel(),
    "co-resident pair = x neighbours, same z": pair_order(idx.T.ravel()),
    "co-resident pair = z neighbours, same x": pair_order(idx.ravel()),
    "random": np.random.default_rng(0).permutation(N),
}
ref = None
for name, perm in orders.items():
    p = torch.from_numpy(np.ascontiguousarray(perm)).to(dev)
    xs, zs = xm[p].contiguous(), zm[p].contiguous()
    for _ in range(3):
        de, kick = ops.wake_mesh(hist, lat, wp, xs, zs)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        de, kick = ops.wake_mesh(hist, lat, wp, xs, zs)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    full = torch.empty(N, dtype=torch.float64, device=dev)
    full[p] = de
    if ref is None:
        ref = full.clone()
    print(f"{name:42s} median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms  same bits as shipped order: {bool(torch.equal(full, ref))}")
