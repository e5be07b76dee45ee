"""GPU tests at BASELINE.json's larger configurations, through size-independent properties
(conservation, linearity, partition invariance, idempotence) plus oracle spot checks on sub-samples
the CPU can finish in seconds."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import dfcsr_oracle as O
from tests import scenario

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _up(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n,shape", [(10_000_000, (100, 100)), (50_000_000, (64, 512))])
def test_deposit_conservation_and_linearity_at_scale(n, shape):
    """configs[2]/[4] particle counts.  A grid that covers every particle conserves the weights:
    sum(count) = Np, sum(vxsum) = sum(px); NGP counts sum to Np bit-exactly; deposit(A u B) =
    deposit(A) + deposit(B); both code paths agree."""
    import torch
    from pydfcsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 6.0e-5
    z = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 2.0e-4
    px = torch.randn(n, generator=g, device="cuda", dtype=torch.float64) * 4.0e-6
    nx, nz = shape
    lim_x, lim_z = float(x.abs().max()) * 1.08, float(z.abs().max()) * 1.08   # > one cell of margin
    args = (nx, -lim_x, lim_x, nz, -lim_z, lim_z)
    c2, v2 = (t.clone() for t in ops.deposit_cic(x, z, px, *args, mode=2))
    assert abs(float(c2.sum()) - n) <= 1e-9 * n
    assert abs(float(v2.sum()) - float(px.sum())) <= 1e-9 * float(px.abs().sum())
    assert float(c2.min()) >= 0.0
    for mode in (1, 3):          # shared-memory tile paths (the 64x512 grid only partly fits the tile)
        c1, v1 = ops.deposit_cic(x, z, px, *args, mode=mode)
        assert float((c1 - c2).abs().max()) <= 1e-10 * float(c2.max())
        assert float((v1 - v2).abs().max()) <= 1e-10 * float(v2.abs().max())
    h = n // 2
    ca, va = (t.clone() for t in ops.deposit_cic(x[:h], z[:h], px[:h], *args))
    cb, vb = ops.deposit_cic(x[h:], z[h:], px[h:], *args)
    assert float((ca + cb - c2).abs().max()) <= 1e-10 * float(c2.max())
    assert float((va + vb - v2).abs().max()) <= 1e-10 * float(v2.abs().max())
    ngp = ops.deposit_ngp(x, z, *args)
    assert int(ngp.sum()) == n and int(ngp.min()) >= 0
    # oracle spot check on the first 200k particles
    m = 200_000
    ref = O.cic_deposit_2d(x[:m].cpu().numpy(), z[:m].cpu().numpy(), np.ones(m), *args)
    got, _ = ops.deposit_cic(x[:m].contiguous(), z[:m].contiguous(), px[:m].contiguous(), *args)
    assert _rel(got.cpu().numpy(), ref) < 1e-12


def test_make_df_fine_longitudinal_grid():
    """configs[4]: micro-bunched, tilted beam on a 64 x 512 deposit grid, window 9 / order 2."""
    from pydfcsr_b200 import DF_tracker, synth
    b = synth.gaussian_bunch(1_000_000, seed=3, tilt=2.5, modulation=0.1, modulation_wavelength_sigma=0.05)
    cfg = dict(xbins=64, zbins=512, xlim=5, zlim=5, filter_order=2, filter_window=9, velocity_threhold=1000, upper_limit=2000)
    df = O.make_density_functions(b[0], b[4], b[1], 0.0, O.DepositConfig(**cfg))
    assert df.density.shape == (64, 512)          # the tilt makes the reference honour the YAML grid
    trk = DF_tracker(cfg, device="cuda:0")
    trk.get_DF(_up(b[0]), _up(b[4]), _up(b[1]), 0.0)
    for k in O.FIELDS:
        assert _rel(getattr(trk, k), getattr(df, k)) < 1e-10, k


def test_regrid_2000x2000_slice():
    """Largest history slice the reference allows (upper_limit: 2000): bit-exact against the
    scipy-equivalent oracle, and idempotent (re-gridding onto the source nodes returns the source)."""
    import torch
    from pydfcsr_b200 import ops
    from pydfcsr_b200._lib import Axis
    rng = np.random.default_rng(5)
    src = rng.normal(size=(5, 300, 300))
    sx, sz = np.linspace(-3e-4, 3e-4, 300), np.linspace(-1e-3, 1e-3, 300)
    dx, dz = np.linspace(-3.3e-4, 3.1e-4, 2000), np.linspace(-0.9e-3, 1.2e-3, 2000)
    out = torch.empty((2000, 2000, 6), dtype=torch.float64, device="cuda")
    ops.history_regrid(_up(src), Axis.make(sx[0], sx[-1], 300), Axis.make(sz[0], sz[-1], 300),
                       Axis.make(dx[0], dx[-1], 2000), Axis.make(dz[0], dz[-1], 2000), -0.5, out)
    got = out.cpu().numpy()
    for k in (0, 4):
        assert np.array_equal(got[..., k], O.regrid_bilinear(src[k], sx, sz, dx, dz, -0.5 if k == 4 else 0.0))
    same = torch.empty((300, 300, 6), dtype=torch.float64, device="cuda")
    ops.history_regrid(_up(src), Axis.make(sx[0], sx[-1], 300), Axis.make(sz[0], sz[-1], 300),
                       Axis.make(sx[0], sx[-1], 300), Axis.make(sz[0], sz[-1], 300), 0.0, same)
    assert np.array_equal(np.moveaxis(same.cpu().numpy()[..., :5], -1, 0), src)


@pytest.mark.parametrize("tilt,mesh", [(0.0, (128, 128)), (2.5, (64, 512))])
def test_wake_partition_invariance_large_mesh(tilt, mesh):
    """configs[2]/[4] meshes: the result of a mesh point does not depend on how the mesh is cut into
    rank blocks (bitwise), 8-way split reassembles to the single launch, and a sub-sample matches the
    oracle to 1e-10."""
    import torch
    from pydfcsr_b200 import ops
    from pydfcsr_b200.distributed import split_counts
    sc = scenario.chicane_entry(tilt=tilt)
    st, lat = sc["stack"], sc["lattice"]
    hist = ops.DeviceHistory.from_stacks([st.data[k] for k in O.FIELDS], st.min_x, st.min_y, st.min_z,
                                         st.delta_x, st.delta_y, st.delta_z, "cuda:0", cap=16, head=13)   # wraps
    dlat = ops.DeviceLattice.upload(lat.coords, lat.n_vec, lat.tau_vec, lat.rho, lat.distance, lat.min_s, lat.delta_s, "cuda:0")
    nx = nz = 64
    wp = ops.wake_params(nx=nx, nz=nz, **sc["wake_scalars"])
    x, z = sc["coords"][0], sc["coords"][4]
    s = sc["scalars"]
    xm, zm, _, _ = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], 3, 3, *mesh)
    dxm, dzm = _up(xm), _up(zm)
    de, kick = (t.clone() for t in ops.wake_mesh(hist, dlat, wp, dxm, dzm))
    assert bool(torch.isfinite(de).all()) and bool(torch.isfinite(kick).all())
    n = xm.size
    count, displ = split_counts(n, 8)
    parts = [ops.wake_mesh(hist, dlat, wp, dxm, dzm, first=d, count=c)[0].clone() for d, c in zip(displ, count)]
    assert torch.equal(torch.cat(parts), de)
    idx = np.linspace(0, n - 1, 24).astype(int)
    osc = O.WakeScalars(nx=nx, nz=nz, **sc["wake_scalars"])
    ref_de, ref_kick = O.wake_mesh(xm[idx], zm[idx], osc, lat, st)
    assert np.max(np.abs(de.cpu().numpy()[idx] - ref_de)) <= 1e-10 * np.max(np.abs(ref_de))
    assert np.max(np.abs(kick.cpu().numpy()[idx] - ref_kick)) <= 1e-10 * np.max(np.abs(ref_kick))


def test_deep_history_window():
    """configs[3]: long retarded-time window.  60 pushes through a 40-slice window on the device ring
    (pop-left + wrap-around + growth) against the oracle's deque."""
    import torch
    from pydfcsr_b200 import DF_tracker, synth
    cfg = dict(scenario.DEPOSIT_CFG)
    trk = DF_tracker(cfg, device="cuda:0")
    hist = O.HistoryOracle(O.DepositConfig(**cfg))
    b = synth.gaussian_bunch(40_000, seed=8)
    x, px, z = b[0], b[1], b[4]
    dx, dpx, dz = _up(x), _up(px), _up(z)
    for k in range(60):
        t = 0.05 * k
        zk = z * (1.0 + 0.002 * k)
        trk.get_DF(dx, _up(zk), dpx, t)
        trk.append_DF()
        fl = float("inf") if k == 0 else 1.975
        got = trk.append_interpolant(fl, 1)
        hist.append(O.make_density_functions(x, zk, px, t, hist.cfg))
        assert got == hist.push(fl, 1)
    assert len(trk.time_interp) == len(hist.time_interp) == 40
    assert trk._ring.shape[0] >= 40 and trk._head != 0
    trk.build_interpolant()
    ref = hist.stack()
    assert abs(trk.delta_x - ref.delta_x) < 1e-15 and trk.min_x == ref.min_x
    for name in ("density", "vx_x"):
        got = getattr(trk, f"data_{name}_interp")
        assert _rel(got[::13], ref.data[name][::13]) < 1e-10


def test_two_rank_nccl_pipeline():
    """N > 1 path on real GPUs (skipped on a single-GPU box): torchrun with 2 ranks, NCCL all-gather,
    rank results identical to a single-GPU launch."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    script = os.path.join(ROOT, "tests", "nccl_worker.py")
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "nccl ok 0" in out.stdout and "nccl ok 1" in out.stdout
    print(out.stdout[-300:])          # which exchange ran (fused peer-memory stores or the NCCL fallback)
