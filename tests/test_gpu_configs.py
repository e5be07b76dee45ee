"""BASELINE.json configs[2], [3], [4] at FULL size on one B200: CSR2D.run() builds the history on the device, the wake
is evaluated on the full mesh, and six mesh points (the largest |dE| plus five random ones) are recomputed by the CPU
oracle from the device's own history (gate 1e-10 of the mesh maximum, north_star).  The multi-GPU runs of these
configurations cut the same mesh into rank blocks; that the result of a point does not depend on the cut is tested
bitwise in test_gpu_scale.py, test_gpu_kernels.py (fused exchange) and recorded on every multi-GPU bench line."""
import numpy as np
import pytest

from tests import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,history_min,n_particle", [
    ("lcls_bc", (7, 500, 500), 10_000_000),            # configs[2]: 1e7 particles, sigma_z 20 um, 128 x 128 mesh
    ("arc", (100, 1500, 1500), 1_000_000),             # configs[3]: 8-dipole arc, n_formation_length 4 -> T ~ 111, 256 x 128 mesh
    ("microbunched", (5, 500, 500), 50_000_000),       # configs[4]: 5e7 particles, tilt 2.5 + modulation, 64 x 512 mesh
])
def test_baseline_config_full_size(name, history_min, n_particle):
    import torch
    torch.cuda.empty_cache()
    csr, stop = configs.build(name)
    assert csr.beam.x.numel() == n_particle
    csr.wake_counters = torch.zeros(3, dtype=torch.int64, device=csr.device)
    csr.run(stop_time=stop - 1e-9)
    trk = csr.DF_tracker
    shape = (len(trk.time_interp), len(trk.x_grid_interp), len(trk.z_grid_interp))
    assert all(a >= b for a, b in zip(shape, history_min)), shape
    mesh = configs.CONFIGS[name][4]
    assert tuple(csr.dE_dct.shape) == mesh
    assert bool(torch.isfinite(csr.dE_dct).all()) and bool(torch.isfinite(csr.x_kick).all())
    e_de, e_kick, picks = configs.spot_check(csr)
    print(f"{name}: history {shape}, mesh {mesh}, parity dE {e_de:.2e} kick {e_kick:.2e} at points {picks.tolist()}")
    assert e_de < 1e-10 and e_kick < 1e-10
    # every sample the reference evaluates was accounted for by the last launch's counters
    n_in, n_all, n_gat = (int(v) for v in csr.wake_counters.cpu())
    assert n_all > 0 and 0 < n_gat <= n_in < n_all
    del csr
    torch.cuda.empty_cache()
