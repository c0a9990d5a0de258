"""GPU tests against the committed golden fixtures (tests/golden/*.npz), which were produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  These need neither the oracle nor the
reference tree at run time: the CUDA path is compared with stored reference outputs directly."""
import os

import numpy as np
import pytest

from test_gpu_parity import assert_close, RTOL

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAUSS_FIXTURES = ["gauss_1d_rdf_mcgdp", "gauss_1d_periodic", "gauss_1d_inner_mcgdp", "gauss_2d_mixed", "gauss_2d_mcgdp",
                  "gauss_3d_inner_mcgdp"]
BIAS_FIXTURES = ["bias_c5_tight_limiter", "bias_1d_local_tempering", "bias_2d_local_tempering", "bias_1d_targeting",
                 "bias_3d_density", "bias_2d_walls_threshold"]


@pytest.fixture(scope="module")
def edm():
    import edm_b200
    if edm_b200.device_count() == 0:
        pytest.fail("no CUDA device visible: the GPU tests cannot fall back to the CPU")
    return edm_b200


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", GAUSS_FIXTURES)
def test_gauss_fixture(edm, name):
    z = load(name)
    dim = int(z["dim"])
    g = edm.GaussGrid(dim, z["min"], z["max"], z["spacing"], z["periodic"], 1, z["sigma"])
    if bool(z["has_boundary"]):
        g.set_boundary(z["bmin"], z["bmax"], z["bper"])
    info = g.info()
    assert np.array_equal(info["n"], z["n"]) and np.array_equal(info["dx"], z["dx"])
    assert np.array_equal(info["max"], z["gmax"]) and np.array_equal(info["minisize"], z["minisize"])
    ba = g.add_values(z["centres"], z["heights"])
    assert np.array_equal(ba == 0.0, z["bias_added"] == 0.0)
    assert_close(ba, z["bias_added"], "bias_added")
    v, d = g.get_arrays()
    assert np.array_equal(v == 0.0, z["grid"] == 0.0)
    assert_close(v, z["grid"], "grid")
    assert_close(d, z["deriv"], "grid derivative")
    g.set_arrays(z["grid"], z["deriv"])
    val, der = g.eval(z["x"])
    assert np.array_equal(val == 0.0, z["value"] == 0.0)
    assert_close(val, z["value"], "value")
    assert_close(der, z["der"], "derivative")
    assert_close(g.get_value(z["x"]), z["get_value"], "get_value")


@pytest.mark.parametrize("name", BIAS_FIXTURES)
def test_bias_fixture(edm, name, tmp_path):
    z = load(name)
    f = tmp_path / "case.edm"
    f.write_text(str(z["edm_text"]) + "\nhills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (tmp_path, tmp_path))
    target, expected = None, 0.0
    if "target_values" in z.files:   # the reference's own read-back of the target grid file
        D = len(z["target_min"])
        target = edm.Grid(D, z["target_min"], z["target_max"], z["target_spacing"], z["target_periodic"], 0, 0)
        assert np.array_equal(target.info()["n"], z["target_n"]) and np.array_equal(target.info()["dx"], z["target_dx"])
        target.set_arrays(z["target_values"])
        expected = float(z["expected_target"])
    b = edm.bias_from_edm(str(f), float(z["T"]), float(z["kB"]), z["sublo"], z["subhi"], z["sublo"], z["subhi"],
                          z["periodic"], z["skin"], target=target, expected_target=expected)
    for k, (x, u) in enumerate(zip(z["x"], z["u"])):
        x = np.ascontiguousarray(x)
        force = np.zeros_like(x)
        e = b.update_forces(x, force)
        if k:
            assert abs(e - z["energy"][k]) <= RTOL * abs(z["energy"][k])
            assert_close(force, z["forces"][k], "forces step %d" % k)
        b.add_hills(x, u)
    log = b.log()
    assert np.array_equal(log["steps"], z["log_steps"])
    assert np.array_equal(log["type"], z["log_type"])
    assert np.array_equal(log["hills_added"], z["log_hills_added"])
    assert np.allclose(log["height"], z["log_height"], rtol=0, atol=0.6e-8)      # HILLS keeps 8 decimals
    assert np.allclose(log["bias_added"], z["log_bias_added"], rtol=0, atol=0.6e-8)
    left, right, buf = b.backlog()
    assert (left, right) == (int(z["backlog_left"]), int(z["backlog_right"]))
    assert_close(buf, z["backlog"], "backlog")
    v, d = b.bias_grid.get_arrays()
    assert_close(v, z["grid"], "bias grid")
    assert_close(d, z["deriv"], "bias grid derivative")
    assert np.array_equal(b.hist_grid.get_arrays()[0], z["hist"])
    st = b.state()
    assert abs(st["cum_bias"] - float(z["cum_bias"])) <= RTOL * abs(float(z["cum_bias"]))


def test_plumed_grid_known_answer(edm):
    """edm_test.cpp:117-125 (value 1.260095 at {0.75, 0, 1.00} of tests/3.grid) and interpolation on it."""
    z = load("plumed_grids")
    n, per = z["n3"], z["periodic3"]
    bins = np.where(per == 1, n, n - 1)
    mx0 = np.where(per == 1, z["max3"], z["max3"] - z["dx3"])
    g = edm.Grid(3, z["min3"], mx0, None, per, b_deriv=int(z["b_derivatives3"]), b_interp=0, header_bins=bins)
    assert np.array_equal(g.info()["n"], n)
    g.set_arrays(z["grid3"], z["deriv3"])
    assert abs(g.get_value(z["known_x"])[0] - 1.260095) < 1e-5
    g.set_interpolation(1)
    val, der = g.eval(z["x3"])
    assert_close(val, z["value3"], "3.grid interpolation", rtol=1e-9)
    assert_close(der, z["der3"], "3.grid derivative", rtol=1e-9)
