"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference through oracle/refstub.py) on seeded synthetic inputs.

    python -m oracle.make_golden            # only works where /root/reference exists

The fixtures travel to the GPU box (the reference cannot); tests/test_golden.py checks the oracle
restatement AND the CUDA path against them.  Inputs are never stored: they are regenerated from the
seeds recorded here by the same code (pydfcsr_b200.synth / tracking, tests/scenario.py).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
HIST_STRIDE = 211            # history stacks are pinned on every 211th voxel (flattened) + their sums

from oracle import refstub  # noqa: E402
from pydfcsr_b200 import synth, tracking  # noqa: E402
from tests import scenario  # noqa: E402


def _lattice_yaml(path):
    import yaml
    with open(path, "w") as fh:
        yaml.safe_dump(dict(synth.chicane_lattice_config()), fh, sort_keys=False)
    return path


def golden_cic():
    refstub.load_reference()
    from pyDFCSR_2D.deposit import histogram_cic_2d
    b = synth.gaussian_bunch(200_000, seed=5)
    x, px, z = b[0], b[1], b[4]
    xs, xe = np.mean(x) - 2.5 * np.std(x), np.mean(x) + 3.0 * np.std(x)
    zs, ze = np.mean(z) - 5 * np.std(z), np.mean(z) + 5 * np.std(z)
    cnt = histogram_cic_2d(x, z, np.ones_like(x), 37, xs, xe, 53, zs, ze)
    vxs = histogram_cic_2d(x, z, px, 37, xs, xe, 53, zs, ze)
    np.savez(os.path.join(OUT, "cic.npz"), n_particle=200_000, seed=5, shape=(37, 53), bounds=(xs, xe, zs, ze),
             count=cnt, vxsum=vxs)


def golden_df():
    refstub.load_reference()
    from pyDFCSR_2D.deposit import DF_tracker
    for name, tilt, order, window in (("untilted", 0.0, 1, 9), ("tilted_o1", 2.5, 1, 9), ("tilted_o2", 2.5, 2, 9)):
        b = synth.gaussian_bunch(200_000, seed=11, tilt=tilt)
        cfg = dict(scenario.DEPOSIT_CFG, filter_order=order, filter_window=window)
        tr = DF_tracker(cfg)
        tr.get_DF(b[0], b[4], b[1], 0.25)
        np.savez(os.path.join(OUT, f"df_{name}.npz"), n_particle=200_000, seed=11, tilt=tilt, order=order, window=window,
                 x_grids=tr.x_grids, z_grids=tr.z_grids, density=tr.density, vx=tr.vx, density_x=tr.density_x,
                 density_z=tr.density_z, vx_x=tr.vx_x, mean_vx_x=np.mean(tr.vx_x))


def golden_wake():
    """Reference DF_tracker + CSR2D driven through the steps of tests/scenario.chicane_entry."""
    lat_yaml = _lattice_yaml("/tmp/dfcsr_golden_lattice.yaml")
    for name, tilt in (("notilt", 0.0), ("tilt_pos", 2.5), ("tilt_neg", -2.5)):
        sc = scenario.chicane_entry(tilt=tilt)
        csr = refstub.make_reference_csr(lat_yaml, scenario.DEPOSIT_CFG, dict(n_formation_length=1, zbins=50, xbins=50),
                                         dict(xbins=5, zbins=7, xlim=3, zlim=3, workdir="/tmp"))
        for st in sc["steps"]:
            x, px, y, py, z, pz = st["coords"]
            csr.DF_tracker.get_DF(x=x, z=z, px=px, t=st["pos"])
            csr.DF_tracker.append_DF()
            csr.DF_tracker.append_interpolant(formation_length=st["formation_length"], n_formation_length=1)
        tr = csr.DF_tracker
        tr.build_interpolant()
        x, px, y, py, z, pz = sc["coords"]
        beam = refstub.FakeBeam(x, px, z, pz, sc["pos"])
        csr.beam = beam
        csr.CSR_scaling = 8.98755e3 * beam.charge
        csr.formation_length = sc["steps"][-1]["formation_length"]
        csr.get_CSR_mesh()
        csr.calculate_2D_CSR()
        out = dict(tilt=tilt, n_particle=100_000, seed=1, hist_stride=HIST_STRIDE, hist_shape=tr.data_density_interp.shape,
                   meta=np.array([tr.min_x, tr.min_y, tr.min_z, tr.delta_x, tr.delta_y, tr.delta_z], dtype=np.float64),
                   xmesh=csr.CSR_xmesh, zmesh=csr.CSR_zmesh, xrange=csr.CSR_xrange_transformed, zrange=csr.CSR_zrange,
                   dE_dct=csr.dE_dct, x_kick=csr.x_kick,
                   scalars=np.array([beam.position, beam._sigma_x, beam._sigma_z, beam._slope[0], beam._slope[1],
                                     beam._mean_x, beam._mean_z, csr.formation_length, csr.CSR_scaling]))
        for fld, attr in (("density", "data_density_interp"), ("density_x", "data_density_x_interp"),
                          ("density_z", "data_density_z_interp"), ("vx", "data_vx_interp"), ("vx_x", "data_vx_x_interp")):
            a = getattr(tr, attr)
            out[f"hist_{fld}"] = a.ravel()[::HIST_STRIDE].copy()
            out[f"hist_{fld}_sum"] = np.array([a.sum(), np.abs(a).sum()])
        # integrand arrays of one point (debug=True), 24 x 24 nodes
        csr.integration_params.xbins = csr.integration_params.zbins = 24
        s_dbg = beam.position + 0.3 * beam._sigma_z
        x_dbg = 0.4 * beam._sigma_x
        dbg = csr.get_CSR_wake(s_dbg, x_dbg, debug=True)
        arrays = [a for a in dbg if isinstance(a, np.ndarray) and a.ndim == 2]
        out["dbg_point"] = np.array([s_dbg, x_dbg])
        out["dbg_integrand_z"] = np.concatenate([a.ravel() for a in arrays[0::2]])
        out["dbg_integrand_x"] = np.concatenate([a.ravel() for a in arrays[1::2]])
        # kick application (beams.py:108-131) through scipy, every 50th particle stored
        from scipy.interpolate import RegularGridInterpolator
        xt = beam.x_transform
        pts = np.array([xt, z]).T
        dpz = RegularGridInterpolator((csr.CSR_xrange_transformed, csr.CSR_zrange), 0.1 * csr.dE_dct * 1e6 / 5e9,
                                      fill_value=0.0, bounds_error=False)(pts)
        dpx = RegularGridInterpolator((csr.CSR_xrange_transformed, csr.CSR_zrange), 0.1 * csr.x_kick * 1e6 / 5e9,
                                      fill_value=0.0, bounds_error=False)(pts)
        out["kick_stride"] = 50
        out["pz_new"] = (pz + dpz)[::50]
        out["px_new"] = (px + dpx)[::50]
        np.savez(os.path.join(OUT, f"wake_{name}.npz"), **out)
        print(name, "max|dE|", np.max(np.abs(csr.dE_dct)), "slope", beam._slope[0])


def sgolay2d_input():
    """Seeded input of the sgolay2d fixture: the CIC density of a tilted bunch on a 48 x 70 grid (serial oracle
    deposit, so the array is reproducible to the bit)."""
    from oracle import dfcsr_oracle as O
    b = synth.gaussian_bunch(60_000, seed=21, tilt=1.0)
    x, z = b[0], b[4]
    return O.cic_deposit_2d(x, z, np.ones_like(x), 48, np.mean(x) - 4 * np.std(x), np.mean(x) + 4 * np.std(x),
                            70, np.mean(z) - 4 * np.std(z), np.mean(z) + 4 * np.std(z))


def golden_sgolay2d():
    refstub.load_reference()
    from pyDFCSR_2D.SGolay_filter import sgolay2d
    z = sgolay2d_input()
    out = {}
    for window, order in ((7, 2), (5, 3)):
        out[f"smooth_{window}_{order}"] = sgolay2d(z, window, order)
        out[f"col_{window}_{order}"], out[f"row_{window}_{order}"] = sgolay2d(z, window, order, derivative="both")
    np.savez(os.path.join(OUT, "sgolay2d.npz"), cases=np.array([(7, 2), (5, 3)]), checksum=float(z.sum()), **out)


def golden_mpi_split():
    """test/test_mpi.py:14-17 evaluated for a few (work_size, ranks) pairs."""
    rows = []
    for n, p in ((10, 4), (127, 8), (300, 7), (4096, 8), (5, 8)):
        ave, res = divmod(n, p)
        count = [ave + 1 if r < res else ave for r in range(p)]
        displ = [sum(count[:r]) for r in range(p)]
        rows.append((n, p, count, displ))
    np.savez(os.path.join(OUT, "mpi_split.npz"), cases=np.array([(n, p) for n, p, _, _ in rows]),
             **{f"count_{n}_{p}": np.array(c) for n, p, c, _ in rows}, **{f"displ_{n}_{p}": np.array(d) for n, p, _, d in rows})


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    golden_cic()
    golden_df()
    golden_wake()
    golden_mpi_split()
    golden_sgolay2d()
    print("golden vectors written to", OUT)
