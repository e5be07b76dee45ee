"""Developer tool: where does a bench step go?  Host time per stage (no syncs added) and GPU time per stage
(CUDA events), for the configs[1] step of bench.py.  python tools/step_breakdown.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydfcsr_b200 import CSR2D  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
wl = bench.WORKLOAD
csr = CSR2D(bench._input_dict(wl), parallel=False, verbose=False)
csr.run(stop_time=wl["position"] - 0.05)
trk, beam = csr.DF_tracker, csr.beam
trk.pop_right_interpolant()
pristine = [c.clone() for c in beam.coords]
host = {}
gpu = {}


def stage(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    fn()
    b.record()
    host.setdefault(name, []).append(time.perf_counter() - t0)
    gpu.setdefault(name, []).append((a, b))


def one_step():
    stage("restore", lambda: (beam.coords[1].copy_(pristine[1]), beam.coords[5].copy_(pristine[5])))
    stage("stats#1 (enqueue) + prefetch", lambda: (beam.update_status(), trk.prefetch_DF(beam)))
    stage("get_DF (K1+K2)", lambda: trk.get_DF(x=beam.x, z=beam.z, px=beam.px, t=beam.position, stats=beam.stats))
    stage("append+regrid (K3)", lambda: (trk.append_DF(), trk.append_interpolant(csr.formation_length, 1), trk.build_interpolant()))
    stage("mesh (host+H2D)", csr.get_CSR_mesh)
    stage("wake (K4)", csr.calculate_2D_CSR)
    stage("kick (K5) + stats#2 (sync)", lambda: beam.apply_wakes(csr.dE_dct, csr.x_kick, csr.CSR_xrange_transformed, csr.CSR_zrange, 0.1, 1))
    stage("pop", trk.pop_right_interpolant)


for _ in range(5):
    one_step()
host.clear(); gpu.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    one_step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
print(f"wall per step {wall * 1e3:.3f} ms")
tot_h = tot_g = 0.0
for k in host:
    h = np.median(host[k]) * 1e6
    g = np.median([a.elapsed_time(b) for a, b in gpu[k]]) * 1e3
    tot_h += h; tot_g += g
    print(f"  {k:32s} host {h:8.1f} us   gpu-span {g:8.1f} us")
print(f"  {'sum':32s} host {tot_h:8.1f} us   gpu-span {tot_g:8.1f} us")
