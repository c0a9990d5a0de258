"""BASELINE.json configs 3 and 4 at FULL size on the GPU (2-D 4096^2 grid with 10^7 coordinates and local
well-tempering; 3-D 512^3 grid): size-independent properties, plus parity with the oracle on a sample where
the oracle's memory allows (2-D: ~0.8 GB of host memory for the reference's over-allocated arrays).

Properties: the sum of add_value's returns equals the integral of the grid (a checksum of checksums);
each hill integrates to its height up to the mass outside the support (exp(-8) per dimension pair);
update_forces over all atoms equals the batched grid evaluation of the same points; periodic images
evaluate alike; forces are minus the finite-difference gradient of the energy."""
import os
import tempfile

import numpy as np
import pytest

from test_gpu_parity import RTOL, assert_close

pytestmark = pytest.mark.gpu

C3_TEXT = ("tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 250\n"
           "dimension 2\nbox_low 0 0\nbox_high 64 64\nbias_spacing 0.015625 0.015625\nbias_sigma 0.0625 0.0625\n")
C4_TEXT = ("tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 250\ndimension 3\nbox_low 0 0 0\n"
           "box_high 64 64 64\nbias_spacing 0.125 0.125 0.125\nbias_sigma 0.25 0.25 0.25\n")


@pytest.fixture(scope="module")
def edm():
    import edm_b200
    if edm_b200.device_count() == 0:
        pytest.fail("no CUDA device visible: the GPU tests cannot fall back to the CPU")
    return edm_b200


def make_bias(edm, text, D):
    d = tempfile.mkdtemp()
    f = os.path.join(d, "full.edm")
    open(f, "w").write(text + "hills_filename %s/H\nhistogram_filename %s/G\n" % (d, d))
    lo, hi = [0.0] * D, [64.0] * D
    return edm.bias_from_edm(f, 300.0, 0.0019872, lo, hi, lo, hi, [1] * D, [0.0] * D), f


def test_c3_full_size_2d(edm, port):
    bd, f = make_bias(edm, C3_TEXT, 2)
    g = bd.bias_grid
    assert tuple(g.info()["n"]) == (4096, 4096)
    rng = np.random.default_rng(1234 + 2)
    n = 10_000_000
    # three steps of fix edm at full size: forces, selection, local-tempering round
    logs_before = 0
    for step in range(3):
        x = np.ascontiguousarray(rng.uniform(0, 64, size=(n, 2)))
        fd = np.zeros((n, 2))
        e = bd.step_coords(x, fd, runiform=None, seed=7, step=step)
        if step:
            assert e > 0 and np.abs(fd).max() > 0
    log = bd.log()
    st = bd.state()
    assert st["steps"] == 3 and 600 < len(log) < 900          # ~250 hills per step
    info = bd.round_info()
    assert info["parallel"] == 3, info
    # checksum of checksums: integral of the grid == sum of add_value's returns == cum_bias
    v, dv = g.get_arrays()
    vol = 0.015625 ** 2
    assert abs(v.sum() * vol - log["bias_added"].sum()) <= 1e-10 * log["bias_added"].sum()
    assert abs(st["cum_bias"] - log["bias_added"].sum()) <= 1e-12 * st["cum_bias"]
    # each hill integrates to its height minus the mass outside the support (exp(-8) of it in 2-D)
    rel = log["bias_added"] / log["height"]
    assert np.all(rel < 1.0) and np.all(rel > 1.0 - 2 * np.exp(-8.0))
    # local well-tempering: heights never exceed the untempered height and shrink where hills pile up
    assert log["height"].max() <= 0.02 / 250 * (1 + 1e-12) and log["height"].min() < 0.02 / 250
    # update_forces over all atoms == batched grid evaluation of the same points (another kernel)
    fd = np.zeros((n, 2))
    e = bd.update_forces(x, fd)
    val, der = g.eval(x)
    assert abs(e - val.sum()) <= 1e-10 * abs(val.sum())
    assert np.array_equal(fd, -der)
    # periodic images
    sub = x[:100000]
    v2, d2 = g.eval(sub + np.array([64.0, -128.0]))
    assert np.abs(v2 - val[:100000]).max() <= 1e-9 * np.abs(val).max()
    # oracle parity on a sample: same hills (the device's own log) deposited by the reference's algorithm
    go = port.GaussGrid("port", 2, [0.0, 0.0], [64.0, 64.0], [0.015625, 0.015625], [1, 1], 1, [0.0625, 0.0625])
    bo = go.add_values(np.ascontiguousarray(log["pos"][:, :2]), log["height"])
    assert_close(log["bias_added"], bo, "bias_added at full size")
    hit = np.abs(val) > 0
    pts = np.concatenate([x[hit][:20000], x[:5000]])
    vo, do = go.eval(pts)
    vd, dd = g.eval(pts)
    assert_close(vd, vo, "interpolated bias at full size")
    assert_close(dd, do, "interpolated derivative at full size")


def test_c4_full_size_3d(edm):
    bd, f = make_bias(edm, C4_TEXT, 3)
    g = bd.bias_grid
    assert tuple(g.info()["n"]) == (512, 512, 512)
    rng = np.random.default_rng(1234 + 3)
    n = 1_250_000                                   # one GPU's share of the 10^7 atoms
    for step in range(3):
        x = np.ascontiguousarray(rng.uniform(0, 64, size=(n, 3)))
        fd = np.zeros((n, 3))
        e = bd.step_coords(x, fd, runiform=None, seed=11, step=step)
    log = bd.log()
    st = bd.state()
    assert st["steps"] == 3 and 600 < len(log) < 900 and bd.round_info()["parallel"] == 3
    rel = log["bias_added"] / log["height"]
    assert np.all(rel < 1.0) and np.all(rel > 1.0 - 3 * np.exp(-8.0) * 8)     # truncated 3-D Gaussian
    assert abs(st["cum_bias"] - log["bias_added"].sum()) <= 1e-12 * st["cum_bias"]
    # evaluate near the deposited hills so that the bias is non-zero
    centres = log["pos"][rng.integers(0, len(log), 200000)]
    pts = np.ascontiguousarray(centres + rng.normal(0, 0.2, size=centres.shape))
    fd = np.zeros_like(pts)
    e = bd.update_forces(pts, fd)
    val, der = g.eval(pts)
    assert e > 0 and abs(e - val.sum()) <= 1e-10 * val.sum()
    assert np.array_equal(fd, -der)
    # forces are minus the gradient: central differences of the interpolated energy inside one cell
    h = 1e-4
    inner = pts[:20000].copy()
    frac = (inner / 0.125) % 1.0
    keep = np.all((frac > 0.1) & (frac < 0.9), axis=1)
    inner = inner[keep]
    _, d0 = g.eval(inner)
    for d in range(3):
        step_v = np.zeros(3)
        step_v[d] = h
        vp, _ = g.eval(inner + step_v)
        vm, _ = g.eval(inner - step_v)
        fdiff = (vp - vm) / (2 * h)
        assert np.abs(fdiff - d0[:, d]).max() <= 2e-5 * np.abs(d0).max()
    # periodic images
    v2, _ = g.eval(pts[:50000] + np.array([64.0, 0.0, -64.0]))
    assert np.abs(v2 - val[:50000]).max() <= 1e-9 * np.abs(val).max()
