"""BASELINE.json's five configurations as CSR2D inputs, plus the oracle spot check used on them at full problem size.
Shared by tests/test_gpu_configs.py (the driver-run parity tests) and tools/run_configs.py (timings)."""
import os

import numpy as np

from oracle import dfcsr_oracle as O

DEPOSIT = dict(xbins=300, zbins=300, xlim=5, zlim=5, filter_order=1, filter_window=9, velocity_threhold=1000,
               upper_limit=2000)
ANGLE = 0.0483
ARC = [("D0", "drift", 0.1, 0.0, 0.0, 0.0, 1)]
for k in range(8):                                   # 8 dipoles of the chicane's strength, same bending direction
    ARC += [(f"B{k + 1}", "dipole", 0.5, ANGLE, 0.0, 0.0, 1), (f"DD{k + 1}", "drift", 0.25, 0.0, 0.0, 0.0, 1)]

CONFIGS = {
    # name: (beam kwargs, elements, step size, stop position, mesh (xbins, zbins), integration, deposit overrides)
    "example": (dict(n_particle=100_000, seed=0), None, 0.1, 0.6, (10, 30), dict(n_formation_length=1, xbins=200, zbins=200), {}),
    "chicane_1e6": (dict(n_particle=1_000_000, seed=0), None, 0.1, 0.6, (64, 64), dict(n_formation_length=1, xbins=200, zbins=200), {}),
    "lcls_bc": (dict(n_particle=10_000_000, seed=0, sigma_z=20.0e-6, chirp=-360.0), None, 0.1, 0.6, (128, 128),
                dict(n_formation_length=1, xbins=200, zbins=200), {}),
    "arc": (dict(n_particle=1_000_000, seed=0), ARC, 0.05, 5.5, (256, 128), dict(n_formation_length=4, xbins=200, zbins=200), {}),
    "microbunched": (dict(n_particle=50_000_000, seed=0, tilt=2.5, modulation=0.1, modulation_wavelength_sigma=0.05), None, 0.1, 0.4,
                     (64, 512), dict(n_formation_length=1, xbins=200, zbins=200), dict(xbins=64, zbins=512, filter_order=2)),
}


def build(name):
    beam, elements, step, stop, mesh, integ, dep = CONFIGS[name]
    from pydfcsr_b200 import synth
    elements = elements or synth.CHICANE_ELEMENTS
    elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in elements]
    lat = synth.chicane_lattice_config(elements=elements)
    lat["step_size"] = step
    inp = {"input_beam": dict(style="synthetic", **beam),
           "input_lattice": {"lattice_config": lat},
           "particle_deposition": dict(DEPOSIT, **dep),
           "CSR_integration": integ,
           "CSR_computation": dict(compute_CSR=1, apply_CSR=0, transverse_on=1, write_beam=None, write_wakes=False,
                                   workdir="/tmp/dfcsr_cfg", xbins=mesh[0], zbins=mesh[1], xlim=5, zlim=5)}
    from pydfcsr_b200 import CSR2D
    csr = CSR2D(inp, parallel=False, verbose=False, precision=os.environ.get("DFCSR_PRECISION", "fp64"))
    if os.environ.get("DFCSR_SKIP"):                 # developer runs: force the zero-density skipping policy
        csr.skip_mode = os.environ["DFCSR_SKIP"]
    return csr, stop


def oracle_points(csr, picks):
    trk = csr.DF_tracker
    data = {f: np.ascontiguousarray(getattr(trk, f"data_{f}_interp")) for f in O.FIELDS}
    stack = O.HistoryStack(data, float(trk.min_x), float(trk.min_y), float(trk.min_z), float(trk.delta_x),
                           float(trk.delta_y), float(trk.delta_z))
    lat = O.LatticeTables(coords=csr.lattice.coords, n_vec=csr.lattice.n_vec, tau_vec=csr.lattice.tau_vec,
                          rho=np.asarray(csr.lattice.rho, dtype=np.float64), distance=np.asarray(csr.lattice.distance, dtype=np.float64),
                          min_s=float(csr.lattice.min_x), delta_s=float(csr.lattice.delta_x))
    b = csr.beam
    ip = csr.integration_params
    sc = O.WakeScalars(t=b.position, sigma_x=b._sigma_x, sigma_z=b._sigma_z, slope0=float(b._slope[0]), mean_x=b._mean_x,
                       formation_window=ip.n_formation_length * csr.formation_length, csr_scaling=csr.CSR_scaling,
                       nx=ip.xbins, nz=ip.zbins)
    xm, zm = np.asarray(csr.CSR_xmesh), np.asarray(csr.CSR_zmesh)
    return np.array([O.wake_point(b.position + zm[k], xm[k], sc, lat, stack) for k in picks])




def spot_check(csr, n_random=5, seed=0):
    """(max rel dE error, max rel kick error, picks) of the device wake grids against the CPU oracle evaluated on the
    device's own history at the largest |dE| point plus n_random random mesh points; errors relative to the mesh maximum."""
    n_mesh = csr.CSR_params.xbins * csr.CSR_params.zbins
    de, kick = csr.dE_dct.cpu().numpy().ravel(), csr.x_kick.cpu().numpy().ravel()
    rng = np.random.default_rng(seed)
    picks = np.concatenate([[int(np.argmax(np.abs(de)))], rng.choice(n_mesh, n_random, replace=False)])
    ref = oracle_points(csr, picks)
    e_de = np.max(np.abs(de[picks] - ref[:, 0])) / np.max(np.abs(de))
    e_k = np.max(np.abs(kick[picks] - ref[:, 1])) / np.max(np.abs(kick))
    return float(e_de), float(e_k), picks
