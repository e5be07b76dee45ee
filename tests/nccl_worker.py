"""Worker for test_two_rank_nccl_pipeline: CSR2D(parallel=True) on 2 GPUs vs the serial launch, with the particles
replicated on every rank (the reference's layout) and sharded over the ranks (this implementation's default)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydfcsr_b200 import CSR2D, synth  # noqa: E402

elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in synth.CHICANE_ELEMENTS]


def make(parallel, shard, apply_csr, xbins=5, zbins=7):
    inp = {"input_beam": {"style": "synthetic", "n_particle": 100_000, "seed": 1},
           "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
           "particle_deposition": dict(xbins=64, zbins=96, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                       velocity_threhold=1000, upper_limit=2000),
           "CSR_integration": dict(n_formation_length=1, zbins=40, xbins=40),
           "CSR_computation": dict(compute_CSR=1, apply_CSR=apply_csr, transverse_on=1, xbins=xbins, zbins=zbins, xlim=3, zlim=3,
                                   write_beam=None, write_wakes=False, workdir="/tmp/dfcsr_nccl")}
    return CSR2D(inp, parallel=parallel, verbose=False, shard_particles=shard)


csr = make(True, False, 0)                  # particles replicated on every rank, like every MPI rank of the reference
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

# ---- particles sharded over the ranks: every bit of a run with the kick applied must be the single-GPU run's ----------
one = make(False, False, 1)                  # this rank alone, all particles
one.run(stop_time=0.45)
shd = make(True, True, 1)
assert shd.beam.shards is not None and shd.beam.x.numel() < one.beam.x.numel()
shd.run(stop_time=0.45)
mode = shd.beam.shards.mode
assert torch.equal(shd.dE_dct, one.dE_dct) and torch.equal(shd.x_kick, one.x_kick), "sharded particles: wake grids differ"
lo, hi = shd.beam.shards.lo, shd.beam.shards.hi
for k in range(6):
    assert torch.equal(shd.beam.coords[k], one.beam.coords[k][lo:hi]), f"sharded particles: coordinate {k} differs"
st_a, st_b = shd.beam.stats, one.beam.stats
assert all(float(a) == float(b) for a, b in zip(st_a[:14], st_b[:14])), "sharded particles: statistics differ"
tw_a, tw_b = shd.beam.twiss, one.beam.twiss
assert all(tw_a[k] == tw_b[k] for k in tw_b), "sharded particles: Twiss parameters differ"
full = shd.beam.to_host()
assert full.shape == (6, one.beam.x.numel()) and (full[0] == one.beam.x.cpu().numpy()).all()
if mode == "peers":                           # same run through the NCCL fallback of the shard exchange
    os.environ["DFCSR_FUSED_GATHER"] = "0"
    alt = make(True, True, 1)
    assert alt.beam.shards.mode == "nccl"
    alt.run(stop_time=0.45)
    assert torch.equal(alt.dE_dct, one.dE_dct) and torch.equal(alt.beam.coords[1], shd.beam.coords[1]), "NCCL shard exchange differs"
    os.environ["DFCSR_FUSED_GATHER"] = "1"
# ---- a mesh row of 32 points: K4 runs in its x-group mapping (one lane per point); the groups are dealt out round-robin,
# stored into all ranks' grids over peer memory (or gathered by NCCL), and the result must be bitwise the serial launch's
xg = make(True, False, 0, xbins=32, zbins=5)
xg.skip_mode = "off"
xg.run(stop_time=0.25)
assert xg.last_wake_mapping == "xgroup"
par = (xg.dE_dct.clone(), xg.x_kick.clone())
xg.calculate_2D_CSR()
assert torch.equal(par[0], xg.dE_dct) and torch.equal(par[1], xg.x_kick), "x-groups: sharded != serial"
assert float(par[0].abs().max()) > 0
if xg._peer_grid is not None:
    keep, xg._peer_grid = xg._peer_grid, None
    xg.calculate_2D_CSR_parallel()
    assert torch.equal(xg.dE_dct, par[0]) and torch.equal(xg.x_kick, par[1]), "x-groups: NCCL gather != fused exchange"
    xg._peer_grid = keep
os.write(1, f"nccl ok {csr.rank} exchange={path} shards={mode}\n".encode())   # one write per rank: print() pieces interleave
torch.distributed.destroy_process_group()
