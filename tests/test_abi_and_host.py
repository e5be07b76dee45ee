"""CPU tests: the C-ABI library loads and exports every symbol include/dfcsr_b200.h declares, the
ctypes structs match the header layout, and the host-side logic (lattice, YAML, block split,
all-gather on gloo with world_size 2) behaves like the reference's."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dfcsr_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfcsr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pydfcsr_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 14
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib.dfcsr_abi_version() == _lib.ABI_VERSION
    assert _lib.lib.dfcsr_beam_stats_workspace() > 0
    assert _lib.lib.dfcsr_make_df_workspace(100, 100) > 2 * 100 * 100 * 8      # two scratch planes + tile partials


def test_struct_layouts_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    from pydfcsr_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dfcsr_b200.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(dfcsr_axis), sizeof(dfcsr_history), sizeof(dfcsr_lattice),'
                   ' sizeof(dfcsr_wake_params), offsetof(dfcsr_history, min_t), offsetof(dfcsr_lattice, d_rho),'
                   ' offsetof(dfcsr_wake_params, nx));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    exp = [C.sizeof(_lib.Axis), C.sizeof(_lib.History), C.sizeof(_lib.Lattice), C.sizeof(_lib.WakeParams),
           _lib.History.min_t.offset, _lib.Lattice.d_rho.offset, _lib.WakeParams.nx.offset]
    assert got == exp


def test_errors_are_reported_not_raised_across_the_abi():
    from pydfcsr_b200 import _lib
    rc = _lib.lib.dfcsr_deposit_cic(None, None, None, 10, 4, 0.0, 1.0, 4, 0.0, 1.0, None, None, 0, None)
    assert rc == -1 and b"null pointer" in _lib.lib.dfcsr_last_error()
    with pytest.raises(_lib.DfcsrError):
        _lib.check(rc, "dfcsr_deposit_cic")


def test_no_cpu_fallback():
    import torch
    from pydfcsr_b200 import _lib, ops
    with pytest.raises(_lib.DfcsrError):
        ops.beam_stats(torch.zeros(8, dtype=torch.float64), torch.zeros(8, dtype=torch.float64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pydfcsr_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
            assert "import_module(\"oracle" not in text and "/root/reference" not in text, fn


def test_lattice_and_yaml(tmp_path):
    import yaml
    from oracle import dfcsr_oracle as O
    from pydfcsr_b200 import synth
    from pydfcsr_b200.lattice import Lattice
    from pydfcsr_b200.yaml_parser import parse_yaml
    path = tmp_path / "lat.yaml"
    path.write_text(yaml.safe_dump(dict(synth.chicane_lattice_config()), sort_keys=False))
    cfg = parse_yaml(str(path))
    assert list(cfg.keys())[0] == "step_size" and list(cfg.keys())[1:] == [e[0] for e in synth.CHICANE_ELEMENTS]
    lat = Lattice({"lattice_input_file": str(path)})
    ref = O.reference_orbit([(e[1], e[2], e[3]) for e in synth.CHICANE_ELEMENTS])
    assert np.array_equal(lat.coords, ref.coords) and np.array_equal(lat.n_vec, ref.n_vec)
    assert np.array_equal(lat.rho, ref.rho) and np.array_equal(lat.distance, ref.distance)
    assert lat.total_steps == 134 and lat.steps_per_element.sum() == 133          # 13.3124 m / 0.1 m
    assert abs(lat.coords[-1, 1]) < 1e-9 and abs(lat.tau_vec[-1, 1]) < 1e-12      # chicane closes
    with pytest.raises(FileNotFoundError):
        parse_yaml(str(tmp_path / "missing.yaml"))


def test_savgol_operators_match_oracle():
    from oracle import dfcsr_oracle as O
    from pydfcsr_b200 import ops
    for w, o in ((5, 0), (5, 2), (9, 1), (9, 2), (1, 0)):
        a, b = ops.savgol_operators(w, o), O.savgol_operators(w, o)
        for p, q in zip(a, b):
            assert np.allclose(p, q, rtol=0, atol=1e-15)
        assert abs(a[0].sum() - 1) < 1e-14


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pydfcsr_b200 import distributed as D
rank, world = D.init_process_group("gloo")
for n in (7, 10, 4096):
    count, displ = D.split_counts(n, world)
    full = np.arange(2 * n, dtype=np.float64).reshape(2, n)
    send = torch.zeros((2, max(count)), dtype=torch.float64)
    send[:, :count[rank]] = torch.from_numpy(full[:, displ[rank]:displ[rank] + count[rank]])
    got = D.all_gather_blocks(send, count, n).numpy()
    assert np.array_equal(got, full), (rank, n)
# x-groups dealt out round-robin (group = iz * ceil(nx/32) + ix // 32, rank = group % world): every rank fills its own
# points of a full grid; the exchange copies every point from its owner (-0.0 stays -0.0, NaN placeholders never leak)
for nx, nz, gw in ((32, 5, 32), (48, 7, 64), (70, 3, 32), (128, 3, 64)):
    n = nx * nz
    owner = D.xgroup_owner(nx, nz, world, torch.device("cpu"), gw)
    ix, iz = np.divmod(np.arange(n), nz)
    assert np.array_equal(owner.numpy(), (iz * ((nx + gw - 1) // gw) + ix // gw) % world)
    full = -np.arange(2 * n, dtype=np.float64).reshape(2, n)          # element 0 is -0.0
    mine = torch.full((2, n), float("nan"), dtype=torch.float64)
    sel = (owner == rank)
    mine[:, sel] = torch.from_numpy(full)[:, sel]
    got = D.all_gather_select(mine, owner).numpy()
    assert np.array_equal(got, full) and np.signbit(got[0, 0]), (rank, nx, nz)
# the fused K4 exchange needs NCCL ranks with CUDA peer mappings: on gloo/CPU the collective decision is "NCCL path"
assert D.make_peer_wake_grid(4096, torch.device("cpu")) is None
# particle shards (host logic; the kernels are covered on the GPU): whole chunks per rank, exact cover, gather restores
# the bunch, and the table exchange of the NCCL/gloo fallback completes every rank's table
for n in (10_003, 1_000_000, 1500):
    sh = D.ParticleShards(n, torch.device("cpu"), max_cells=64, use_peers=False)
    assert sh.mode == "nccl" and sh.chunk == -(-n // 1024)
    ranges = [None] * world
    dist.all_gather_object(ranges, (sh.lo, sh.hi, sh.first_block, sh.n_blocks))
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert sum(r[3] for r in ranges) == 1024 and all(r[0] == min(r[2] * sh.chunk, n) for r in ranges)
    full = torch.arange(n, dtype=torch.float64)
    assert torch.equal(sh.gather(sh.local_slice(full).clone()), full)
    table, ptrs = sh.stats_table(0)
    assert ptrs is None and float(table.abs().sum()) == 0.0
    table.view(1024, 8)[sh.first_block:sh.first_block + sh.n_blocks] = float(rank + 1)
    sh.exchange(table)
    want = torch.cat([torch.full((r[3], 8), float(k + 1), dtype=torch.float64) for k, r in enumerate(ranges)])
    assert torch.equal(table.view(1024, 8), want)
    q, qptr = sh.q_buffer(64)
    q.fill_(rank + 1)
    sh.reduce_q(q)
    assert int(q[0]) == sum(range(1, world + 1)) and len(qptr) == 1
dist.barrier()
open(os.path.join(sys.argv[2], f"ok{rank}"), "w").write("ok")     # one file per rank: stdout of two ranks interleaves
'''


def test_all_gather_blocks_world2_gloo(tmp_path):
    import socket
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as sock:                 # a free rendezvous port (fixed ports collide between runs)
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT, str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_parameter_blocks_mirror_reference_defaults():
    """params.py:15-43: same keys/defaults; unknown keys are a TypeError like an unexpected keyword there."""
    from pydfcsr_b200.params import CSR_params, Integration_params
    ip = Integration_params()
    assert (ip.n_formation_length, ip.zbins, ip.xbins) == (4, 200, 200)
    cp = CSR_params({"xbins": 10, "write_beam": [16, 17]})
    assert (cp.xbins, cp.zbins, cp.xlim, cp.zlim, cp.apply_CSR, cp.compute_CSR, cp.transverse_on) == (10, 30, 5, 5, 1, 1, 1)
    assert cp.write_wakes is True and cp.write_beam == [16, 17] and os.path.isabs(cp.workdir) and cp.write_name == ""
    with pytest.raises(TypeError):
        CSR_params({"not_a_key": 1})


def test_xgroup_plan_is_host_arithmetic():
    """dfcsr_wake_xgroup_plan touches no device memory: it can be pinned on the CPU.  The plan (which K4 mapping runs, unit
    size, scratch per group) must be a function of the history geometry, the beam scalars and the WHOLE mesh -- the rule
    every rank of a parallel run relies on to take the same decision (CSR.py:420-451 splits the mesh, not the physics)."""
    import ctypes as C
    from pydfcsr_b200 import _lib
    from pydfcsr_b200._lib import Axis

    def plan(xbins=64, zbins=64, slope0=-0.4, sigma_z=2.0e-4, skip="off", support=1, nx=200, nz=200, X=500, Z=500):
        sx = 1.0e-4
        hist = _lib.History(0x1000, X * Z * 6, 16, 0, 7, X, Z, 0, 0.0, -5 * sx, -5 * sigma_z, 0.1, 10 * sx / (X - 1),
                            10 * sigma_z / (Z - 1), 0x2000 if support else None)
        wp = _lib.WakeParams(0.6, sx, sigma_z, slope0, 0.0, 0.5, 1e-5, nx, nz, _lib.SKIP_MODES[skip], 0)
        out = _lib.XGroupPlan()
        xa, za = Axis.make(-3 * sx, 3 * sx, xbins), Axis.make(-3 * sigma_z, 3 * sigma_z, zbins)
        assert _lib.lib.dfcsr_wake_xgroup_plan(C.byref(hist), C.byref(wp), xa, za, C.byref(out)) == 0
        return out

    p = plan()
    assert (p.n_groups, p.group_points, p.unit_nodes, p.max_units) == (128, 32, 1, 800)
    assert p.workspace_bytes_per_group == 800 * 64 * 8 + 256
    assert plan(xbins=64, zbins=512).n_groups == 1024 and plan(xbins=64, zbins=512).unit_nodes == 2      # larger units on larger meshes
    assert plan(xbins=10, zbins=30).n_groups == 0                 # 10 of 32 lanes would carry a point
    assert plan(xbins=23).n_groups == 64 and plan(xbins=22).n_groups == 0
    assert plan(slope0=1.5).n_groups == 0 and plan(slope0=-1.0).n_groups == 128      # chirp band (CSR.py:480)
    assert plan(skip="on").n_groups == 0 and plan(skip="on", support=0).n_groups == 128   # skipping needs the row-support table
    assert plan(skip="auto").n_groups == 128                      # a bunch that fills its grid: AUTO does not skip
    assert plan(skip="auto", X=2000).n_groups == 128 and plan(skip="auto", sigma_z=2.0e-4, Z=500, X=500).n_groups == 128
    assert plan(sigma_z=2.0e-5).n_groups == 0                     # compressed bunch: a group spreads over ~80 history cells
    assert plan(nz=400).n_groups == 0                             # node table would not leave room for two CTAs per SM
