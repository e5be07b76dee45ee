#!/usr/bin/env python
"""bench.py — CSR-wake hot path on BASELINE.json configs[1]: the bundled 4-dipole chicane at 1e6
synthetic Gaussian macro-particles, 64x64 observation mesh, 200x200 integration nodes, fp64.

A "step" = one pass of the hot path over one particle batch at a fixed lattice position (0.6 m, end
of the first dipole, 7 history slices): beam statistics -> CIC deposit (K1) -> smoothing/gradients
(K2) -> re-grid into the history ring (K3) -> wake on the mesh (K4, sharded over ranks + all-gather)
-> kick (K5) -> statistics.  Tracking is not part of the path.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # CPU oracle port, all host cores
    torchrun --nproc-per-node N bench.py --gpus N ...            # N > 1: one rank per GPU

metric = obs-points x integrand-samples per second (every sample the reference would evaluate counts:
4*nx*nz per point without chirp band, CSR.py:577-585); ms_per_step = seconds per lattice step * 1e3.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
METRIC = "csr_wake_obs_points_x_integrand_samples_per_s"
UNIT = "point-samples/s"

PORT_NOTE = ("oracle port of get_CSR_wake (numpy temporaries + numba gathers, the reference's structure); per point it is about "
             "1.2x FASTER than the unmodified reference (39.7 vs 48.2 ms at 200x200, measured in the build container against "
             "/root/reference, which does not exist on the GPU box), so GPU/CPU ratios against it are conservative")

WORKLOAD = dict(
    name="chicane_1e6_mesh64x64_int200x200_fp64",
    n_particle=1_000_000, seed=0, position=0.6,
    deposition=dict(xbins=300, zbins=300, xlim=5, zlim=5, filter_order=1, filter_window=9,
                    velocity_threhold=1000, upper_limit=2000),         # example/input/chicane_config.yaml:10-18
    integration=dict(n_formation_length=1, zbins=200, xbins=200),     # chicane_config.yaml:21-24
    mesh=dict(xbins=64, zbins=64, xlim=3, zlim=3),
)


def _input_dict(wl, apply_csr=0, world=1):
    """Weak scaling: every GPU keeps configs[1]'s 64 x 64 = 4096 observation points; with N ranks the mesh is
    64 x (64 N) (finer in z), cut into N contiguous blocks by the reference's MPI split rule."""
    from pydfcsr_b200 import synth
    mesh = dict(wl["mesh"])
    mesh["zbins"] = mesh["zbins"] * world
    elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _nsep) in synth.CHICANE_ELEMENTS]   # CSR every step
    return {
        "input_beam": {"style": "synthetic", "n_particle": wl["n_particle"], "seed": wl["seed"]},
        "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
        "particle_deposition": dict(wl["deposition"]),
        "CSR_integration": dict(wl["integration"]),
        "CSR_computation": dict(compute_CSR=1, apply_CSR=apply_csr, transverse_on=1, write_beam=None,
                                write_wakes=False, workdir="/tmp/dfcsr_bench", **mesh),
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU side: the oracle port (numpy + numba, the reference's own implementation style) on host cores
# -------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_block(args):
    first, count = args
    from oracle import dfcsr_oracle as O
    return O.wake_mesh(_CPU["xm"], _CPU["zm"], _CPU["sc"], _CPU["lat"], _CPU["hist"], first=first, count=count)


def cpu_wake_sample(xm, zm, sc, lat, hist, indices, cores, repeats=1):
    """Wake at mesh points `indices` with `cores` fork workers, each taking a contiguous block of the
    sample (the reference's MPI split rule, CSR.py:121-125).  numba is compiled in the parent first,
    so JIT time is excluded.  Returns (dE, kick, best seconds)."""
    import multiprocessing as mp
    from oracle import dfcsr_oracle as O
    xs, zs = np.ascontiguousarray(xm[indices]), np.ascontiguousarray(zm[indices])
    _CPU.update(xm=xs, zm=zs, sc=sc, lat=lat, hist=hist)
    O.wake_mesh(xs, zs, sc, lat, hist, first=0, count=1)                 # warm the JIT
    count, displ = O.split_counts(len(indices), cores)
    blocks = [(d, c) for d, c in zip(displ, count) if c > 0]
    best = float("inf")
    with mp.get_context("fork").Pool(len(blocks)) as pool:
        pool.map(_cpu_block, [(0, 1)] * len(blocks))                      # page in the workers
        for _ in range(repeats):
            t0 = time.perf_counter()
            parts = pool.map(_cpu_block, blocks)
            best = min(best, time.perf_counter() - t0)
    de = np.concatenate([p[0] for p in parts])
    kick = np.concatenate([p[1] for p in parts])
    return de, kick, best


def oracle_state(wl):
    """Build the bench state (history at 0.6 m, mesh, scalars) with the CPU oracle only."""
    from oracle import dfcsr_oracle as O
    from pydfcsr_b200 import synth, tracking          # pure-host modules: the CUDA library is not loaded by these
    cfg = O.DepositConfig(**wl["deposition"])
    hist = O.HistoryOracle(cfg)
    coords = tuple(synth.gaussian_bunch(wl["n_particle"], seed=wl["seed"]))
    R = 0.5002 / 0.0483
    pos = 0.0

    def log(c, pos, fl):
        hist.append(O.make_density_functions(c[0], c[4], c[1], pos, cfg))
        hist.push(fl, wl["integration"]["n_formation_length"])

    log(coords, 0, float("inf"))
    coords = tracking.track_exact(coords, tracking.Drift(0.1), 5.0e9); pos += 0.1
    fl = 0.1
    log(coords, pos, fl)
    for k in range(5):
        last = (k == 4)
        # CSR2D.run: formation length from sigma_z at element entry (CSR.py:256)
        if k == 0:
            fl = (24 * R ** 2 * 5 * float(np.std(coords[4]))) ** (1 / 3)
        el = tracking.SBend(L=0.1, G=0.0483 / 0.5002, E1=0.0, E2=0.0,
                            FRINGE_AT="entrance_end" if k == 0 else "no_end")
        coords = tracking.track_exact(coords, el, 5.0e9); pos += 0.1
        log(coords, pos, fl)
        del last
    x, z = coords[0], coords[4]
    s = O.beam_scalars(x, z)
    m = wl["mesh"]
    xm, zm, xr, zr = O.observation_mesh(x, z, s["slope"], s["sigma_z"], s["mean_z"], m["xlim"], m["zlim"], m["xbins"], m["zbins"])
    sc = O.WakeScalars(t=pos, sigma_x=float(s["sigma_x"]), sigma_z=float(s["sigma_z"]), slope0=float(s["slope"][0]),
                       mean_x=float(s["mean_x"]), formation_window=wl["integration"]["n_formation_length"] * fl,
                       csr_scaling=8.98755e3 * 1.0e-9, nx=wl["integration"]["xbins"], nz=wl["integration"]["zbins"])
    lat = O.reference_orbit([(e[1], e[2], e[3]) for e in synth.CHICANE_ELEMENTS])
    return xm, zm, sc, lat, hist.stack()


def l1_probe():
    """Measured L1 -> register load bandwidth of this GPU (tools/l1_probe.cu, built by pydfcsr_b200/build.py):
    the unit that bounds the wake kernel.  None when the probe binary is absent."""
    exe = os.path.join(ROOT, "pydfcsr_b200", "l1_probe")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception:
        return None


def samples_per_point(sc):
    return (4 if abs(sc.slope0) <= 1 else 5) * sc.nx * sc.nz


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the wake (oracle port: numpy
    temporaries + numba gathers, exactly the reference's structure) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = dict(WORKLOAD)
    if args.gpus > 1:                       # same weak-scaling mesh and workload name as the GPU arm at this N
        wl["mesh"] = dict(wl["mesh"], zbins=wl["mesh"]["zbins"] * args.gpus)
        wl["name"] = wl["name"].replace("mesh64x64", f"mesh{wl['mesh']['xbins']}x{wl['mesh']['zbins']}")
    cores = os.cpu_count() or 1
    xm, zm, sc, lat, hist = oracle_state(wl)
    n = len(xm)
    per_core = max(1, args.ref_points_per_core)
    idx = np.linspace(0, n - 1, min(n, cores * per_core)).astype(np.int64)      # deterministic sub-mesh
    spp = samples_per_point(sc)
    times = []
    for k in range(args.warmup + args.steps):
        _, _, sec = cpu_wake_sample(xm, zm, sc, lat, hist, idx, cores)
        if k >= args.warmup:
            times.append(sec)
    total = float(np.sum(times))
    value = len(idx) * spp * len(times) / total
    sample = (f"{len(idx)} of {n} mesh points (evenly spaced sub-mesh) x {spp} samples per step, wake stage only "
              f"(get_CSR_wake), {cores} fork workers with the reference's block split; mpi4py absent")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "mesh": [wl["mesh"]["xbins"], wl["mesh"]["zbins"]], "integration": [sc.nx, sc.nz],
                       "n_particle": wl["n_particle"], "history": list(hist.shape)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "kind_note": PORT_NOTE, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pydfcsr_b200 import CSR2D, _lib, tracking
    from oracle import dfcsr_oracle as O          # checker / cpu_baseline leg only

    wl = WORKLOAD
    world = int(os.environ.get("WORLD_SIZE", "1"))
    parallel = world > 1
    if args.gpus != world:
        print(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun for N>1", file=sys.stderr)
    probe = l1_probe() if int(os.environ.get("RANK", "0")) == 0 else None
    csr = CSR2D(_input_dict(wl, world=world), parallel=parallel, verbose=False, precision=args.precision)
    rank, dev = csr.rank, csr.device
    csr.run(stop_time=wl["position"] - 0.05)                  # builds the 7-slice history on the device
    assert abs(csr.beam.position - wl["position"]) < 1e-9, csr.beam.position
    trk = csr.DF_tracker
    trk.pop_right_interpolant()                               # the timed step re-deposits the 0.6 m slice
    beam = csr.beam
    sharded = beam.shards is not None
    n_particle_total = beam.n_total
    pristine = [c.clone() for c in beam.coords]
    host = [c.cpu().pin_memory() for c in pristine]
    n_pts = csr.CSR_params.xbins * csr.CSR_params.zbins
    wp = csr._wake_params()
    spp = (4 if abs(wp.slope0) <= 1 else 5) * wp.nx * wp.nz
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    counters = torch.zeros(3, dtype=torch.int64, device=dev)
    csr.wake_counters = counters
    out_host = [torch.empty_like(host[1]).pin_memory(), torch.empty_like(host[5]).pin_memory(),
                torch.empty((2, csr.CSR_params.xbins, csr.CSR_params.zbins), dtype=torch.float64).pin_memory()]
    k4_events = []
    k4_e2e_events = []          # K4 inside the pipelined end-to-end loop (copies in flight on the second stream)

    def hot_path(timed_k4, sink=None):
        beam.update_status()
        b = beam
        trk.prefetch_DF(b)            # as CSR2D.run does: deposit + density functions enqueued behind the statistics pass
        trk.get_DF(x=b.x, z=b.z, px=b.px, t=b.position, stats=b.stats)
        trk.append_DF()
        trk.append_interpolant(formation_length=csr.formation_length,
                               n_formation_length=csr.integration_params.n_formation_length)
        trk.build_interpolant()
        csr.get_CSR_mesh()
        if timed_k4:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if parallel:
            csr.calculate_2D_CSR_parallel()
        else:
            csr.calculate_2D_CSR()
        if timed_k4:
            ev[1].record()
            (k4_events if sink is None else sink).append(ev)
        b.apply_wakes(csr.dE_dct, csr.x_kick, csr._mesh_axes[0], csr._mesh_axes[1], 0.1, 1)
        trk.pop_right_interpolant()

    def step_resident(timed_k4=False):
        flush.zero_()                                         # L2 flush between steps
        beam.coords[1].copy_(pristine[1]); beam.coords[5].copy_(pristine[5])   # kick is in place: restore px, pz
        hot_path(timed_k4)

    def step_e2e():
        flush.zero_()
        for k in (0, 1, 4, 5):                                # x, px, z, pz from pinned host memory
            beam.coords[k].copy_(host[k], non_blocking=True)
        hot_path(False)
        out_host[0].copy_(beam.coords[1], non_blocking=True)
        out_host[1].copy_(beam.coords[5], non_blocking=True)
        if rank == 0:                                         # every rank holds the same gathered grids: one download
            out_host[2][0].copy_(csr.dE_dct, non_blocking=True)
            out_host[2][1].copy_(csr.x_kick, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # End-to-end loop with the copies taken off the critical path: step k+1's particle batch is uploaded, and
    # step k's results are downloaded, on a copy stream while the compute stream works (double-buffered device
    # inputs).  Every step still moves its own inputs H2D from pinned memory and its own results D2H.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    dbuf = [[torch.empty_like(c) for c in pristine] for _ in range(2)]
    for b in range(2):
        for k in (2, 3):
            dbuf[b][k].copy_(pristine[k])

    def e2e_pipelined(steps, diag=""):
        up = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        saved = beam.coords

        def upload(b):
            with torch.cuda.stream(copy_stream):
                if diag not in ("nocopy", "noup"):
                    for k in (0, 1, 4, 5):
                        dbuf[b][k].copy_(host[k], non_blocking=True)
                up[b].record(copy_stream)

        upload(0)
        for i in range(steps):
            b = i & 1
            flush.zero_()
            if i + 1 < steps:
                upload(1 - b)          # enqueued BEFORE this step's (host-synchronising) hot path; on the copy
                                       # stream it is ordered after the download of step i-1 from that buffer
            main_stream.wait_event(up[b])
            beam.coords = dbuf[b]
            hot_path(True, k4_e2e_events)
            res = (csr.dE_dct, csr.x_kick)
            done[b].record(main_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[b])
                if diag in ("nocopy", "nodown"):
                    continue
                out_host[0].copy_(dbuf[b][1], non_blocking=True)
                out_host[1].copy_(dbuf[b][5], non_blocking=True)
                if rank == 0:
                    out_host[2][0].copy_(res[0], non_blocking=True)
                    out_host[2][1].copy_(res[1], non_blocking=True)
        copy_stream.synchronize()
        beam.coords = saved

    def barrier():
        if parallel:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, **kw):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(**kw)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if parallel:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    sampler = ClockSampler(dev.index or 0).start() if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step_resident()
    # device-side sample accounting for the roofline figure: ONE untimed step with the counters attached (every step
    # works on the same restored batch); the timed steps run without them, like a production run
    counters.zero_()
    step_resident()
    torch.cuda.synchronize()
    n_in_local = float(int(counters[0]))
    n_gat_local = float(int(counters[2]))                 # in-grid samples whose voxels were actually gathered
    csr.wake_counters = None
    launches0 = _lib.lib.dfcsr_launch_count()
    ms_total = timed(step_resident, args.steps, timed_k4=True)
    launches = _lib.lib.dfcsr_launch_count() - launches0
    k4_ms = float(np.mean([a.elapsed_time(b) for a, b in k4_events]))
    for _ in range(2):
        step_e2e()
    ms_e2e_serial = timed(step_e2e, args.steps)
    e2e_pipelined(3)
    barrier()
    k4_e2e_events.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_pipelined(args.steps)
    e1.record()
    barrier()
    ms_pipe = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if parallel:
        dist.all_reduce(ms_pipe, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_pipe[0])
    if os.environ.get("DFCSR_BENCH_E2E_DIAG") and rank == 0:      # developer diagnostic: which copies cost what
        for mode in ("nocopy", "noup", "nodown", ""):
            e2e_pipelined(3, mode)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_pipelined(args.steps, mode)
            torch.cuda.synchronize()
            print(f"[bench diag] pipelined loop, {mode or 'all copies'}: {(time.perf_counter() - t0) / args.steps * 1e3:.3f} ms per step", file=sys.stderr)
    clocks = sampler.stop() if sampler else None

    ms_step = ms_total / args.steps
    value = n_pts * spp / (ms_step * 1e-3)
    e2e_value = n_pts * spp / (ms_e2e / args.steps * 1e-3)

    # ---- parity on every line (all ranks take part: the parallel launch is collective) -----------------------------------
    hot_state = _final_state(csr, trk, O, pristine, parallel)
    strong = None if args.no_strong else strong_scaling_record(world, args.steps, flush)
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # 5 fields x 8 corners x 8 B per in-grid sample; 4 B in the optional fp32-storage mode
    k4_bytes = (320.0 if args.precision == "fp64" else 160.0) * n_gat_local
    traffic = None                      # DRAM bytes per K4 launch from the committed ncu --set full capture
    traffic_src = None
    ncu_pipes = {}
    mapping = getattr(csr, "last_wake_mapping", "point")
    k4_name = ("wake_xgroup_kernel (K4, one lane per observation point of a mesh row)" if mapping == "xgroup"
               else "wake_mesh_kernel_p (K4, one CTA per observation point)")
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "k4_ncu_traffic.json")))
        tr = tr.get(mapping, tr)        # one record per K4 mapping ("xgroup" / "point")
        if args.precision == "fp64":
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            traffic_src = tr.get("source")
            ncu_pipes = {k: tr[k] for k in ("fp64_pipe_pct", "l1_data_pipe_pct", "issue_active_pct", "kernel") if k in tr}
    except Exception:
        pass
    achieved = k4_bytes / (k4_ms * 1e-3) / 1e9
    # The unit that bounds K4 is the L1 -> register load path (the 96 MB stack is L2-resident, DRAM traffic ~0.2 % of
    # peak): the peak is measured on this GPU by tools/l1_probe.cu; without the probe binary, nominal 128 B/clk/SM.
    sm_mhz = float((clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
    if probe:
        l1_peak = max(probe["l1_broadcast_gbs"], probe["l1_contiguous_gbs"])
        l1_src = "tools/l1_probe.cu on this GPU at process start: 16-byte loads from an L1-resident window"
    else:
        l1_peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        l1_src = "nominal 148 SMs x 128 B/clk x max SM clock (probe binary absent)"
    mesh_now = [csr.CSR_params.xbins, csr.CSR_params.zbins]
    wl_name = wl["name"] if world == 1 else wl["name"].replace("mesh64x64", f"mesh{mesh_now[0]}x{mesh_now[1]}")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "s_per_lattice_step": ms_step * 1e-3,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "mesh": mesh_now, "k4_mapping": mapping,
                   "integration": [wp.nx, wp.nz], "n_particle": wl["n_particle"],
                   "history": list(hot_state["hist"].shape),
                   "samples_per_point": spp, "position_m": wl["position"],
                   "l2": "256 MiB device memset between steps (inside the timed region)",
                   "points_per_gpu": n_pts // world,
                   "parallelism": ((f"obs-mesh x-groups (32 points of a mesh row) dealt out round-robin x{world} (4096 points per GPU), "
                                    if mapping == "xgroup" else f"obs-mesh block split x{world} (4096 points per GPU), ") +
                                   ("exchange fused into K4 (NVLink peer-memory stores + one barrier)"
                                    if getattr(csr, "_peer_grid", None) is not None else "NCCL all-gather") +
                                   (f"; particles sharded x{world} (statistics tables + integer deposit grids combined over "
                                    f"{beam.shards.mode}, bit-identical to one GPU)" if sharded else "; particles replicated"))
                   if parallel else "single GPU"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "k4_ms_per_launch": float(np.mean([a.elapsed_time(b) for a, b in k4_e2e_events])) if k4_e2e_events else None,
                "ms_per_step_unpipelined": ms_e2e_serial / args.steps,
                "how": "per step: x, px, z, pz H2D from pinned host memory, hot path, px, pz and both wake grids D2H; "
                       "copies double-buffered on a second stream (the unpipelined figure serialises them)" +
                       ("; the particles are sharded over the ranks, so every rank moves 1/N of the batch, and the wake "
                        "grids (identical on all ranks) are downloaded by rank 0 only" if sharded else ""),
                "h2d_bytes_per_step": int(n_particle_total * 8 * 4),
                "d2h_bytes_per_step": int(n_particle_total * 8 * 2 + out_host[2].numel() * 8),
                "h2d_bytes_per_step_per_rank": int(sum(host[k].numel() * 8 for k in (0, 1, 4, 5))),
                "d2h_bytes_per_step_rank0": int(out_host[0].numel() * 8 * 2 + out_host[2].numel() * 8)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "l1", "kernel": k4_name, "achieved": achieved, "peak": l1_peak,
                     "unit": "GB/s", "frac": achieved / l1_peak, "traffic": traffic,
                     "peak_source": l1_src, "traffic_source": traffic_src,
                     "hbm_actual": {"frac": (traffic / (k4_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None, "peak": hbm_peak,
                                    "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                                    "how": "DRAM bytes per launch (ncu) / K4 launch time / measured HBM copy peak: the history "
                                           "stack is L2-resident, HBM does not bind this kernel at this configuration"},
                     "k4_ms_per_launch": k4_ms, "k4_share_of_step": k4_ms / ms_step,
                     "in_grid_samples_per_launch": n_in_local, "in_grid_fraction": n_in_local / (n_pts * spp / world),
                     "gathered_samples_per_launch": n_gat_local,
                     "ncu": ncu_pipes,
                     "note": "achieved = algorithmic gather bytes (320 B per in-grid integrand sample that is actually gathered; "
                             "samples whose eight voxels carry no density are skipped via the row-support table and count 0 B) / "
                             "K4 launch time; this traffic never reaches HBM (the history stack is L2-resident), so the roofline "
                             "is the L1 -> register load path.  The point kernel moves every one of these bytes through that "
                             "path; the x-group kernel serves a sample's four transverse-blended corners (160 B) from a per-warp "
                             "shared-memory window and is bound by the fp64 pipe instead (roofline.ncu, DESIGN.md section 4)"},
        "parity": hot_state["parity"],
    }
    if strong is not None:
        line["strong"] = strong
    if probe:
        line["roofline"]["probe"] = probe
    if args.precision != "fp64":
        line["dtype"] = "f64 math, f32 history storage (optional mode, wakes within 1e-4)"
        line["config"]["workload"] += "_" + args.precision + "_history"
    cores = os.cpu_count() or 1
    n_all = len(hot_state["xm"])
    if world == 1 and not args.no_cpu_baseline:
        n_chk = min(n_all, cores * args.cpu_points_per_core)
    else:
        n_chk = min(n_all, 64)                      # N > 1: the parity check only (>= 64 evenly spaced points of the gathered grid)
    idx = np.linspace(0, n_all - 1, n_chk).astype(np.int64)
    de, kick, sec = cpu_wake_sample(hot_state["xm"], hot_state["zm"], hot_state["sc"], hot_state["lat"],
                                    hot_state["hist"], idx, min(cores, n_chk))
    g_de, g_kick = hot_state["gpu_de"][idx], hot_state["gpu_kick"][idx]
    line["parity"].update({"points_checked": int(n_chk), "checker": "oracle port (CPU), same history exported from the device",
                           "max_rel_dE": float(np.max(np.abs(g_de - de)) / np.max(np.abs(de))),
                           "max_rel_kick": float(np.max(np.abs(g_kick - kick)) / np.max(np.abs(kick))), "gate": 1e-10})
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = {"value": len(idx) * spp / sec, "unit": UNIT, "cores": cores, "kind": "port", "kind_note": PORT_NOTE,
                                "sample": f"{len(idx)} of {n_all} mesh points (evenly spaced) x {spp} samples, wake stage "
                                          f"only, {cores} fork workers, {sec:.2f} s",
                                "parity_max_rel_dE": line["parity"]["max_rel_dE"],
                                "parity_max_rel_kick": line["parity"]["max_rel_kick"]}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


STRONG = dict(
    name="lcls_bc_1e7_mesh128x128_int200x200_fp64",          # BASELINE.json configs[2]: "1e7 particles, 128x128 mesh, sharded over 8xB200"
    beam=dict(n_particle=10_000_000, seed=0, sigma_z=20.0e-6, chirp=-360.0), position=0.6,
    mesh=dict(xbins=128, zbins=128, xlim=5, zlim=5),
)


def _one_step(csr, events=None):
    """One pass of the hot path at the current lattice position (same stages as the headline step)."""
    import torch
    b, trk = csr.beam, csr.DF_tracker
    b.update_status()
    trk.prefetch_DF(b)
    trk.get_DF(x=b.x, z=b.z, px=b.px, t=b.position, stats=b.stats)
    trk.append_DF()
    trk.append_interpolant(formation_length=csr.formation_length, n_formation_length=csr.integration_params.n_formation_length)
    trk.build_interpolant()
    csr.get_CSR_mesh()
    if events is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if csr.parallel:
        csr.calculate_2D_CSR_parallel()
    else:
        csr.calculate_2D_CSR()
    if events is not None:
        ev[1].record()
        events.append(ev)
    b.apply_wakes(csr.dE_dct, csr.x_kick, csr._mesh_axes[0], csr._mesh_axes[1], 0.1, 1)
    trk.pop_right_interpolant()


def strong_scaling_record(world, steps, flush):
    """STRONG scaling on BASELINE.json configs[2]: the FIXED 128 x 128 mesh and 1e7 particles, full step (statistics, K1-K5),
    on the N ranks of this job (mesh block split + particle shards) and, in the same processes, on every rank alone with all
    particles and the whole mesh (t_1).  Collective; returns the record on every rank."""
    import torch
    import torch.distributed as dist
    from pydfcsr_b200 import CSR2D, synth
    wl = dict(WORKLOAD, n_particle=STRONG["beam"]["n_particle"], mesh=STRONG["mesh"])
    inp = _input_dict(wl, world=1)
    inp["input_beam"] = dict(style="synthetic", **STRONG["beam"])
    steps = max(3, min(steps, 20))

    def measure(parallel):
        csr = CSR2D(inp, parallel=parallel, verbose=False)
        csr.run(stop_time=STRONG["position"] - 0.05)
        csr.DF_tracker.pop_right_interpolant()
        pristine = [c.clone() for c in csr.beam.coords]
        events = []

        def step(ev=None):
            flush.zero_()
            csr.beam.coords[1].copy_(pristine[1]); csr.beam.coords[5].copy_(pristine[5])
            _one_step(csr, ev)

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(events)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps, float(np.mean([a.elapsed_time(b) for a, b in events]))],
                          dtype=torch.float64, device=csr.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        grids = torch.stack([csr.dE_dct, csr.x_kick]).clone()
        shards = csr.beam.shards.mode if csr.beam.shards is not None else None
        mapping = getattr(csr, "last_wake_mapping", "point")
        del csr
        torch.cuda.empty_cache()
        return float(ms[0]), float(ms[1]), grids, (shards, mapping)

    t1, k41, g1, (_, mapping1) = measure(False)
    rec = {"workload": STRONG["name"], "mesh": [STRONG["mesh"]["xbins"], STRONG["mesh"]["zbins"]],
           "n_particle": STRONG["beam"]["n_particle"], "steps": steps, "t_1_ms": t1, "k4_1_ms": k41,
           "non_k4_1_ms": t1 - k41, "k4_mapping": mapping1,
           "how": "full step (statistics, K1, K2, K3, K4 + exchange, K5, statistics) with the particles resident, L2 flushed "
                  "between steps, CUDA events, max over ranks; t_1 = every rank alone with all particles and the whole mesh, "
                  "measured in the same processes"}
    if world > 1:
        tn, k4n, gn, (shards, _) = measure(True)
        rec.update({"n_gpus": world, "t_N_ms": tn, "k4_N_ms": k4n, "non_k4_N_ms": tn - k4n, "speedup": t1 / tn,
                    "efficiency": t1 / (world * tn), "particles": f"sharded ({shards})" if shards else "replicated",
                    "grids_bitwise_equal_to_single_gpu": bool(torch.equal(gn, g1))})
    return rec


def _final_state(csr, trk, O, pristine, parallel):
    """Host copies of exactly what a GPU wake launch of the timed step consumes (history incl. the 0.6 m slice), plus the
    wake grids of that launch.  Collective: with N > 1 every rank runs the sharded launch (K4 + exchange) and then the whole
    mesh by itself; rank 0 reports whether the gathered grid is bitwise the serial one and bitwise equal on all ranks."""
    import torch
    import torch.distributed as dist
    b = csr.beam
    b.coords[1].copy_(pristine[1]); b.coords[5].copy_(pristine[5])
    # redo deposit + push so that the ring holds the 0.6 m slice, run the wake, export, then restore
    b.update_status()
    trk.get_DF(x=b.x, z=b.z, px=b.px, t=b.position, stats=b.stats)
    trk.append_DF()
    trk.append_interpolant(formation_length=csr.formation_length, n_formation_length=csr.integration_params.n_formation_length)
    trk.build_interpolant()
    csr.get_CSR_mesh()
    parity = {}
    if parallel:
        csr.calculate_2D_CSR_parallel()
        par = torch.stack([csr.dE_dct.clone(), csr.x_kick.clone()])
        csr.calculate_2D_CSR()
        ser = torch.stack([csr.dE_dct, csr.x_kick])
        flags = torch.tensor([int(torch.equal(par, ser))], dtype=torch.int32, device=par.device)
        gathered = [torch.empty_like(par) for _ in range(csr.world_size)]
        dist.all_gather(gathered, par)
        flags2 = torch.tensor([int(all(torch.equal(g, par) for g in gathered))], dtype=torch.int32, device=par.device)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        dist.all_reduce(flags2, op=dist.ReduceOp.MIN)
        parity["gathered_equals_serial_launch_bitwise"] = bool(int(flags[0]))
        parity["ranks_bitwise_equal"] = bool(int(flags2[0]))
        parity["grid_checked"] = "the gathered (sharded + exchanged) grid"
        g = par
    else:
        csr.calculate_2D_CSR()
        g = torch.stack([csr.dE_dct, csr.x_kick])
        parity["grid_checked"] = "single-GPU launch"
    torch.cuda.synchronize()
    out = dict(parity=parity)
    if csr.rank == 0:
        stacks = {name: getattr(trk, f"data_{name}_interp") for name in O.FIELDS}
        hist = O.HistoryStack(stacks, trk.min_x, trk.min_y, trk.min_z, trk.delta_x, trk.delta_y, trk.delta_z)
        lat = O.LatticeTables(csr.lattice.coords, csr.lattice.n_vec, csr.lattice.tau_vec, float(csr.lattice.min_x),
                              float(csr.lattice.delta_x), csr.lattice.rho, csr.lattice.distance)
        wp = csr._wake_params()
        sc = O.WakeScalars(t=wp.t, sigma_x=wp.sigma_x, sigma_z=wp.sigma_z, slope0=wp.slope0, mean_x=wp.mean_x,
                           formation_window=wp.formation_window, csr_scaling=wp.csr_scaling, nx=wp.nx, nz=wp.nz)
        out.update(xm=csr.CSR_xmesh, zm=csr.CSR_zmesh, sc=sc, lat=lat, hist=hist,
                   gpu_de=g[0].cpu().numpy().ravel(), gpu_kick=g[1].cpu().numpy().ravel())
    trk.pop_right_interpolant()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (configs[2], 1e7 particles)")
    ap.add_argument("--precision", choices=["fp64", "fp32"], default="fp64",
                    help="history storage: fp64 = parity mode (default, the benchmarked configuration); fp32 = optional mode")
    ap.add_argument("--cpu-points-per-core", type=int, default=256, help="mesh points per host core in the cpu_baseline leg")
    ap.add_argument("--ref-points-per-core", type=int, default=256, help="mesh points per host core per step (--impl reference)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record; everything the library prints on the way (the reference's
    # "start reinterpolation" messages, deposit.py:340) goes to stderr
    global _JSON_OUT
    # ... including what native libraries write to file descriptor 1 (NCCL prints its version there): fd 1 is pointed at
    # stderr for the whole process and the JSON line goes to a duplicate of the original stdout
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        reference_arm(args)
    else:
        args.steps = 100 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        gpu_arm(args)


if __name__ == "__main__":
    main()
