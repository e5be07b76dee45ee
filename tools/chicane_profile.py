"""Developer tool: the WHOLE chicane (133 lattice steps, wake computed and applied at every step) with the bench
mesh, K4 timed per step with CUDA events.  Shows how the wake kernel behaves away from the bench position: the
chirp-band branch after compression, history rebuilds, growing grids.

    python tools/chicane_profile.py [n_particle] [mesh_x] [mesh_z]
    torchrun --nproc-per-node N tools/chicane_profile.py ...        (mesh sharded like pyDFCSR_mpi_run.py)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydfcsr_b200 import CSR2D, synth  # noqa: E402

n_particle = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
mx = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mz = int(sys.argv[3]) if len(sys.argv) > 3 else 64
parallel = "RANK" in os.environ
elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in synth.CHICANE_ELEMENTS]      # CSR at every step
inp = {"input_beam": {"style": "synthetic", "n_particle": n_particle, "seed": 0},
       "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
       "particle_deposition": dict(xbins=300, zbins=300, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                   velocity_threhold=1000, upper_limit=2000),
       "CSR_integration": dict(n_formation_length=1, xbins=200, zbins=200),
       "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, write_beam=None, write_wakes=False,
                               workdir="/tmp/dfcsr_profile", xbins=mx, zbins=mz, xlim=5, zlim=5)}
csr = CSR2D(inp, parallel=parallel, verbose=False)
if os.environ.get("DFCSR_SKIP"):                       # force the zero-density skipping policy (auto / on / off)
    csr.skip_mode = os.environ["DFCSR_SKIP"]
rank, world = (csr.rank, csr.world_size) if parallel else (0, 1)
counters = torch.zeros(3, dtype=torch.int64, device=csr.device)
csr.wake_counters = counters
log = []
inner = csr.calculate_2D_CSR_parallel if parallel else csr.calculate_2D_CSR


def timed():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = counters.clone()
    a.record()
    inner()
    b.record()
    trk = csr.DF_tracker
    log.append((csr.beam.position, float(csr.beam._slope[0]), len(trk.time_interp), len(trk.x_grid_interp),
                len(trk.z_grid_interp), a, b, before, counters.clone(), getattr(csr, "last_wake_mapping", "point")[0]))


if parallel:
    csr.calculate_2D_CSR_parallel = timed
else:
    csr.calculate_2D_CSR = timed
if parallel:          # NCCL communicator set-up (about a second) is not part of the lattice run
    warm = torch.zeros(8, device=csr.device)
    torch.distributed.all_reduce(warm)
    torch.distributed.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
csr.run()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
if rank == 0:
    k4 = np.array([rec[5].elapsed_time(rec[6]) for rec in log])
    full = np.array([int((rec[8] - rec[7])[1]) for rec in log], dtype=np.float64) * world   # whole mesh
    print(f"# full chicane, {n_particle} particles, {mx}x{mz} mesh, 200x200 integration, {world} GPU(s): "
          f"{len(log)} wake steps in {wall:.3f} s wall ({wall / len(log) * 1e3:.2f} ms per lattice step), "
          f"wake kernel + gather {k4.sum() * 1e-3:.3f} s, history rebuilds {csr.DF_tracker.rebuilds}")
    print(f"# samples the reference would evaluate: {full.sum():.3e}  =>  {full.sum() / wall:.3e} samples/s over the whole run")
    print(f"{'s [m]':>7s} {'slope':>8s} {'T':>3s} {'X':>5s} {'Z':>5s} {'wake ms':>8s} {'samples/s':>10s}  K4 mapping (x = x-groups, p = point)")
    for i in list(range(0, len(log), 6)) + [len(log) - 1]:
        pos, slope, T, X, Z = log[i][:5]
        print(f"{pos:7.2f} {slope:8.2f} {T:3d} {X:5d} {Z:5d} {k4[i]:8.3f} {full[i] / (k4[i] * 1e-3):10.3e}  {log[i][9]}")
    nx_steps = sum(1 for rec in log if rec[9] == "x")
    print(f"# steps on the x-group mapping: {nx_steps} of {len(log)}; wake ms on them {sum(k for k, rec in zip(k4, log) if rec[9] == 'x'):.1f}, "
          f"on the point kernel {sum(k for k, rec in zip(k4, log) if rec[9] == 'p'):.1f}")
if parallel:
    torch.distributed.destroy_process_group()
