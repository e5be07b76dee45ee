"""Developer tool: cProfile of the host side of the bench step (which Python functions the critical path spends its
time in between the statistics sync and the wake launch).  python tools/host_profile.py [steps]"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydfcsr_b200 import CSR2D  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
wl = bench.WORKLOAD
csr = CSR2D(bench._input_dict(wl), parallel=False, verbose=False)
csr.run(stop_time=wl["position"] - 0.05)
csr.DF_tracker.pop_right_interpolant()
for _ in range(5):
    bench._one_step(csr)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    bench._one_step(csr)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
