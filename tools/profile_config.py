"""Developer tool: bring one BASELINE.json configuration to its stop position and launch the wake kernel a few times --
the target of `ncu -k regex:wake_mesh_kernel -s <launches of the set-up run> -c 1`.  Prints the device counters of the
last launch (samples the reference evaluates / in-grid / actually gathered) and its CUDA-event time, so that the
algorithmic bytes (320 B or 160 B per gathered sample) can be set against the DRAM bytes ncu reports.

    python tools/profile_config.py arc [fp64|fp32]      # prints "setup_wake_launches N": use -s N with ncu
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
name = sys.argv[1] if len(sys.argv) > 1 else "arc"
os.environ["DFCSR_PRECISION"] = sys.argv[2] if len(sys.argv) > 2 else "fp64"
from tests.configs import build  # noqa: E402

csr, stop = build(name)
launches = {"n": 0}
orig = csr.calculate_2D_CSR


def counted():
    launches["n"] += 1
    orig()


csr.calculate_2D_CSR = counted
csr.run(stop_time=stop - 1e-9)
print("setup_wake_launches", launches["n"], flush=True)
csr.get_CSR_mesh()
cnt = torch.zeros(3, dtype=torch.int64, device=csr.device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=csr.device)
for k in range(3):
    cnt.zero_()
    csr.wake_counters = cnt
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
trk = csr.DF_tracker
n_in, n_all, n_gat = (int(v) for v in cnt.cpu())
vox = 48 if os.environ["DFCSR_PRECISION"] == "fp64" else 32
per = 320 if os.environ["DFCSR_PRECISION"] == "fp64" else 160
ring = trk._ring
print(f"config {name} precision {os.environ['DFCSR_PRECISION']}: history (T, X, Z) = ({len(trk.time_interp)}, {ring.shape[1]}, {ring.shape[2]}), "
      f"{len(trk.time_interp) * ring.shape[1] * ring.shape[2] * vox / 1e9:.2f} GB of voxels; mesh {tuple(csr.dE_dct.shape)}")
print(f"K4 {ms:.3f} ms; samples: reference {n_all}, in grid {n_in}, gathered {n_gat}; algorithmic bytes {per} B x gathered = "
      f"{per * n_gat / 1e9:.2f} GB -> {per * n_gat / (ms * 1e-3) / 1e9:.0f} GB/s; skipping "
      f"{'on' if n_gat < n_in else 'off or nothing to skip'}")
