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
assert torch.equal(par[0], csr.dE_dct) and torch.equal(par[1], csr.x_kick), "sharded != serial"
gathered = [torch.empty_like(par[0]) for _ in range(csr.world_size)]
torch.distributed.all_gather(gathered, par[0])
assert all(torch.equal(g, par[0]) for g in gathered), "ranks disagree"
print("nccl ok", csr.rank, flush=True)
torch.distributed.destroy_process_group()
