"""Worker for test_two_rank_nccl_pipeline: CSR2D(parallel=True) on 2 GPUs vs the serial launch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydfcsr_b200 import CSR2D, synth  # noqa: E402

elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in synth.CHICANE_ELEMENTS]
inp = {"input_beam": {"style": "synthetic", "n_particle": 100_000, "seed": 1},
       "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
       "particle_deposition": dict(xbins=64, zbins=96, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                   velocity_threhold=1000, upper_limit=2000),
       "CSR_integration": dict(n_formation_length=1, zbins=40, xbins=40),
       "CSR_computation": dict(compute_CSR=1, apply_CSR=0, transverse_on=1, xbins=5, zbins=7, xlim=3, zlim=3,
                               write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_nccl")}
csr = CSR2D(inp, parallel=True, verbose=False)
csr.run(stop_time=0.25)
par = (csr.dE_dct.clone(), csr.x_kick.clone())
csr.calculate_2D_CSR()                      # serial launch on this rank, same state
# The default deposit is 64-bit fixed point (integer adds commute), so the replicated histories are bit-identical on
# every rank; the sharded result is compared with this rank's serial launch with a 1e-12 gate (it is bitwise in
# practice: same kernel, same inputs).  What MUST be bitwise identical is the gathered grid every rank kicks with.
for a, b in ((par[0], csr.dE_dct), (par[1], csr.x_kick)):
    err = float((a - b).abs().max() / b.abs().max())
    assert err < 1e-12, f"sharded != serial: {err:.3e}"
# fused K4 + exchange over peer memory (default when the ranks can map each other's grids) against the NCCL all-gather:
# the same kernel computes the same block either way, so the two grids must agree to the last bit
path = "fused" if csr._peer_grid is not None else "nccl"
if csr._peer_grid is not None:
    for _ in range(3):                       # both alternating grids, several rounds
        csr.calculate_2D_CSR_parallel()
        assert torch.equal(csr.dE_dct, par[0]) and torch.equal(csr.x_kick, par[1]), "fused exchange not reproducible"
    keep, csr._peer_grid = csr._peer_grid, None
    csr.calculate_2D_CSR_parallel()
    assert torch.equal(csr.dE_dct, par[0]) and torch.equal(csr.x_kick, par[1]), "fused exchange != NCCL all-gather"
    csr._peer_grid = keep
gathered = [torch.empty_like(par[0]) for _ in range(csr.world_size)]
torch.distributed.all_gather(gathered, par[0])
assert all(torch.equal(g, par[0]) for g in gathered), "ranks disagree"
os.write(1, f"nccl ok {csr.rank} exchange={path}\n".encode())        # one write per rank: print() pieces interleave across ranks
torch.distributed.destroy_process_group()
