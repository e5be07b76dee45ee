"""Output layer (pydfcsr_b200/outputs.py): the driver writes the reference's file layouts (CSR.py:799-835 wakes,
:860-879 statistics) through the h5py call surface; offline the container is an .npz with the same tree."""
import os

import numpy as np
import pytest

from pydfcsr_b200 import outputs


def test_store_mimics_the_h5py_calls_the_reference_uses(tmp_path):
    fn = outputs.store_path(str(tmp_path / "run-wakes"))
    for step in (1, 2):                      # 'a' mode: one group per CSR step appended to the same file (CSR.py:816-820)
        with outputs.open_store(fn, "a") as hf:
            g = hf.create_group("step_" + str(step))
            g.attrs["step"] = step
            g.attrs["position"] = 0.1 * step
            g.attrs["element"] = "B1"
            g1 = g.create_group("longitudinal")
            g1.attrs["unit"] = "MeV/m"
            g1.create_dataset("dE_dct", data=np.full((3, 4), float(step)))
    tree = outputs.load_store(fn)
    assert sorted(tree) == ["step_1", "step_2"]
    assert tree["step_2"]["@attrs"]["step"] == 2 and tree["step_2"]["@attrs"]["element"] == "B1"
    assert tree["step_1"]["longitudinal"]["@attrs"]["unit"] == "MeV/m"
    assert np.array_equal(tree["step_2"]["longitudinal"]["dE_dct"], np.full((3, 4), 2.0))
    with outputs.open_store(fn, "a") as hf:
        with pytest.raises(ValueError):
            hf.create_group("step_1")        # h5py raises on an existing name, and so does the stand-in
    stats = {"twiss": {"alpha_x": np.arange(3.0), "beta_x": np.ones(3)}, "sigma_x": np.zeros(3), "slope": np.zeros((3, 2))}
    fn2 = outputs.store_path(str(tmp_path / "run-statistics"))
    with outputs.open_store(fn2, "w") as hf:
        hf.create_dataset(name="step_positions", data=np.arange(3.0), shape=(3,))
        outputs.dict2hdf5(hf, stats)
    t2 = outputs.load_store(fn2)
    assert sorted(t2) == ["sigma_x", "slope", "step_positions", "twiss"] and sorted(t2["twiss"]) == ["alpha_x", "beta_x"]


@pytest.mark.gpu
def test_driver_writes_the_reference_layouts(tmp_path):
    """A short CSR2D run with write_wakes / write_beam on: file names, groups, datasets and attributes of CSR.py:799-879."""
    import torch  # noqa: F401
    from pydfcsr_b200 import CSR2D, synth
    elements = [(n, k, L, a, e1, e2, 1) for (n, k, L, a, e1, e2, _s) in synth.CHICANE_ELEMENTS]
    inp = {"input_beam": {"style": "synthetic", "n_particle": 50_000, "seed": 4},
           "input_lattice": {"lattice_config": synth.chicane_lattice_config(elements=elements)},
           "particle_deposition": dict(xbins=64, zbins=64, xlim=5, zlim=5, filter_order=1, filter_window=9,
                                       velocity_threhold=1000, upper_limit=1000),
           "CSR_integration": dict(n_formation_length=1, zbins=30, xbins=30),
           "CSR_computation": dict(compute_CSR=1, apply_CSR=1, transverse_on=1, xbins=4, zbins=5, xlim=3, zlim=3,
                                   write_beam=[2], write_wakes=True, write_name="t", workdir=str(tmp_path))}
    csr = CSR2D(inp, parallel=False, device="cuda:0", verbose=False)
    csr.run(stop_time=0.25)
    csr.write_statistics()
    files = sorted(os.listdir(tmp_path))
    wakes = [f for f in files if "-wakes" in f]
    assert len(wakes) == 1 and any("-particles-2" in f for f in files) and any("-statistics" in f for f in files)
    tree = outputs.load_store(os.path.join(tmp_path, wakes[0]))
    assert sorted(tree) == ["step_1", "step_2", "step_3"]
    g = tree["step_3"]
    assert set(g["@attrs"]) == {"step", "position", "mean_gamma", "beam_energy", "element", "charge"}
    assert abs(g["@attrs"]["position"] - 0.3) < 1e-12 and g["@attrs"]["element"] == "B1"
    assert sorted(k for k in g["longitudinal"] if k != "@attrs") == ["dE_dct", "x_grids", "z_grids"]
    assert sorted(k for k in g["transverse"] if k != "@attrs") == ["x_grids", "xkicks", "z_grids"]
    assert g["longitudinal"]["dE_dct"].shape == (4, 5) and g["transverse"]["@attrs"]["unit"] == "MeV/m"
    assert np.array_equal(g["longitudinal"]["dE_dct"], csr.dE_dct.cpu().numpy())
    st = outputs.load_store(os.path.join(tmp_path, [f for f in files if "-statistics" in f][0]))
    assert {"step_positions", "coords", "n_vec", "tau_vec", "twiss", "slope", "sigma_x", "sigma_z", "mean_energy"} <= set(st)
    assert sorted(st["twiss"])[:3] == ["alpha_x", "alpha_y", "beta_x"] and st["sigma_x"][1] > 0
    pt = outputs.load_store(os.path.join(tmp_path, [f for f in files if "-particles-2" in f][0]))
    assert pt["x"].shape == (50_000,) and abs(pt["@attrs"]["position"] - 0.2) < 1e-12
