"""Time the two K4 work splits (one CTA per point / one lane per point, x-groups) on the bench workload in ONE process,
compare their grids, check that the x-group result does not depend on how the groups are dealt out, and count samples.

    python tools/k4_mappings.py [reps]          DFCSR_TILT=0.3 adds an x-z tilt below the chirp-band switch;
                                                DFCSR_PRECISION=fp32 selects the optional fp32 history
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydfcsr_b200 import CSR2D, ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = bench.WORKLOAD
inp = bench._input_dict(wl)
tilt = os.environ.get("DFCSR_TILT")
if tilt:
    inp["input_beam"]["tilt"] = float(tilt)
if os.environ.get("DFCSR_SIGMA_Z"):               # a shorter bunch: the lanes of an x-group spread over more history cells
    inp["input_beam"]["sigma_z"] = float(os.environ["DFCSR_SIGMA_Z"])
mesh = os.environ.get("DFCSR_MESH")
if mesh:
    xb, zb = (int(v) for v in mesh.split("x"))
    inp["CSR_computation"]["xbins"], inp["CSR_computation"]["zbins"] = xb, zb
csr = CSR2D(inp, parallel=False, verbose=False, precision=os.environ.get("DFCSR_PRECISION", "fp64"))
csr.run(stop_time=wl["position"] - 0.05)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=csr.device)
print(f"# K4 mappings on the bench workload, sigma_z={os.environ.get('DFCSR_SIGMA_Z', 'default')}, tilt={tilt or 0}, precision={os.environ.get('DFCSR_PRECISION', 'fp64')}, "
      f"mesh {csr.CSR_params.xbins}x{csr.CSR_params.zbins}, slope {float(csr.beam._slope[0]):.3f}")
res = {}
for mapping in ("point", "auto"):
    csr.wake_mapping = mapping
    csr.wake_counters = torch.zeros(3, dtype=torch.int64, device=csr.device)
    for _ in range(3):
        csr.calculate_2D_CSR()
    csr.wake_counters.zero_()
    csr.calculate_2D_CSR()
    cnt = csr.wake_counters.cpu().numpy().copy()
    csr.wake_counters = None
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
    res[mapping] = (dE, kick)
    print(f"{mapping:6s} ({getattr(csr, 'last_wake_mapping', 'point'):6s}): median {np.median(ts):7.3f} ms  min {np.min(ts):7.3f} ms  "
          f"in-grid {cnt[0]:.4e} of {cnt[1]:.4e}  bitwise_repeatable {rep}", flush=True)
e1 = float((res["auto"][0] - res["point"][0]).abs().max() / res["point"][0].abs().max())
e2 = float((res["auto"][1] - res["point"][1]).abs().max() / res["point"][1].abs().max())
print(f"x-group vs point mapping: dE {e1:.2e} kick {e2:.2e} (of the mesh maximum)")

# split independence: the groups dealt out to 3 'ranks' (stride 3) and in two contiguous halves
wp = csr._wake_params()
plan = csr._xgroup_plan(wp)
if plan is not None:
    lat = csr.lattice.device_tables(csr.device)
    xa, za = csr._mesh_axes
    n = xa.n * za.n
    for label, parts in (("stride 3", [(r, None, 3) for r in range(3)]),
                         ("two blocks", [(0, plan.n_groups // 2, 1), (plan.n_groups // 2, plan.n_groups - plan.n_groups // 2, 1)])):
        out = torch.full((2, n), float("nan"), dtype=torch.float64, device=csr.device)
        for first, count, stride in parts:
            ops.wake_grid_xgroups(csr.DF_tracker.history, lat, wp, xa, za, *csr._mesh_slope, plan=plan, group_first=first,
                                  group_count=count, group_stride=stride, out=out)
        same = bool(torch.equal(out[0], res["auto"][0].reshape(-1)) and torch.equal(out[1], res["auto"][1].reshape(-1)))
        print(f"split '{label}': bitwise equal to the single launch: {same}")
    print(f"plan: groups {plan.n_groups}, unit_nodes {plan.unit_nodes}, max_units {plan.max_units}, "
          f"workspace {plan.n_groups * plan.workspace_bytes_per_group / 1e6:.1f} MB")

# parity against the CPU oracle on a spread of points (history exported from the device)
from oracle import dfcsr_oracle as O  # noqa: E402
trk = csr.DF_tracker
stacks = {name: getattr(trk, f"data_{name}_interp") for name in O.FIELDS}
hist = O.HistoryStack(stacks, trk.min_x, trk.min_y, trk.min_z, trk.delta_x, trk.delta_y, trk.delta_z)
olat = O.LatticeTables(csr.lattice.coords, csr.lattice.n_vec, csr.lattice.tau_vec, float(csr.lattice.min_x),
                       float(csr.lattice.delta_x), csr.lattice.rho, csr.lattice.distance)
sc = O.WakeScalars(t=wp.t, sigma_x=wp.sigma_x, sigma_z=wp.sigma_z, slope0=wp.slope0, mean_x=wp.mean_x,
                   formation_window=wp.formation_window, csr_scaling=wp.csr_scaling, nx=wp.nx, nz=wp.nz)
xm, zm = csr.CSR_xmesh, csr.CSR_zmesh
idx = np.linspace(0, len(xm) - 1, int(os.environ.get("DFCSR_ORACLE_POINTS", 24))).astype(int)
ode, okick = O.wake_mesh(xm[idx], zm[idx], sc, olat, hist)
for mapping in ("point", "auto"):
    g_de = res[mapping][0].reshape(-1).cpu().numpy()[idx]
    g_k = res[mapping][1].reshape(-1).cpu().numpy()[idx]
    print(f"{mapping:6s} vs oracle on {len(idx)} points: dE {np.abs(g_de - ode).max() / np.abs(ode).max():.2e} "
          f"kick {np.abs(g_k - okick).max() / np.abs(okick).max():.2e}")
