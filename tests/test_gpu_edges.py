"""Edge cases of the hill round and the batched calls on the GPU, against the oracle: empty and
fully masked inputs, candidates outside a walled boundary, rounds larger than the parallel plan,
a backlog that fills up (the reference aborts there), rounds that deposit nothing."""
import numpy as np
import pytest

from test_gpu_parity import BIAS_CASES, RTOL, assert_close, compare_bias, write_edm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def edm():
    import edm_b200
    if edm_b200.device_count() == 0:
        pytest.fail("no CUDA device visible: the GPU tests cannot fall back to the CPU")
    return edm_b200


def make_both(edm, port, tmp_path, name, text, T, kB, lo, hi, periodic):
    f = write_edm(tmp_path, name + ".edm", text)
    D = len(lo)
    bo = port.Bias("port", f)
    bo.setup(T, kB)
    bo.subdivide(lo, hi, lo, hi, periodic, [0.0] * D)
    bd = edm.bias_from_edm(f, T, kB, lo, hi, lo, hi, periodic, [0.0] * D)
    return bd, bo


def test_empty_and_fully_masked_rounds(edm, port, tmp_path):
    """n = 0 and "no atom in the group" still run pre_add_hill / post_add_hill: steps advance, nothing is deposited."""
    cfg = BIAS_CASES["c2_rdf_threshold_tempering"]
    bd, bo = make_both(edm, port, tmp_path, "empty", cfg["text"], cfg["T"], cfg["kB"], [1.68], [5.0], [0])
    rng = np.random.default_rng(3)
    x = np.ascontiguousarray(rng.uniform(0.5, 5.5, size=(5000, 3)))
    u = rng.uniform(0, 1, 5000)
    mask = np.ones(5000, np.int32)            # bit 0 only: group bit 2 selects nobody
    for step in range(4):
        fo, fd = np.zeros((5000, 3)), np.zeros((5000, 3))
        if step == 1:                          # empty call
            e0, f0 = np.zeros((0, 3)), np.zeros((0, 3))
            assert bd.update_forces(e0, f0) == 0.0
            bo.add_hills(e0, np.zeros(0), -1)
            bd.add_hills(e0, np.zeros(0))
        elif step == 2:                        # everybody masked out
            bo.set_mask(mask)
            assert bo.update_forces(x, fo, 2) == bd.update_forces(x, fd, mask, 2) == 0.0
            assert not fd.any()
            bo.add_hills(x, u, 2)
            bd.add_hills(x, u, mask, 2)
        else:
            eo = bo.update_forces(x, fo, -1)
            ed = bd.step_coords(x, fd, u)
            if step:
                assert abs(ed - eo) <= RTOL * abs(eo)
                assert_close(fd, fo, "forces")
            bo.add_hills(x, u, -1)
    compare_bias(bd, bo)
    assert bd.state()["steps"] == 4


def test_every_candidate_outside_the_walls(edm, port, tmp_path):
    """Centres outside a non-periodic boundary are accepted, logged and deposit nothing (T10)."""
    text = ("tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 50\ndimension 2\nbox_low 0 0\n"
            "box_high 4 4\nbias_spacing 0.0625 0.0625\nbias_sigma 0.125 0.125")
    bd, bo = make_both(edm, port, tmp_path, "outside", text, 1.0, 1.0, [0.0, 0.0], [4.0, 4.0], [0, 0])
    rng = np.random.default_rng(5)
    for step in range(3):
        x = np.ascontiguousarray(rng.uniform(4.5, 6.0, size=(2000, 3)))
        if step == 2:
            x[::7, :2] = rng.uniform(0.0, 4.0, size=(x[::7].shape[0], 2))   # now a few land inside
        u = rng.uniform(0, 1, 2000)
        fo, fd = np.zeros((2000, 3)), np.zeros((2000, 3))
        bo.update_forces(x, fo, -1)
        bd.update_forces(x, fd)
        assert_close(fd, fo, "forces")
        bo.add_hills(x, u, -1)
        bd.add_hills(x, u)
    log = compare_bias(bd, bo)
    assert (log["bias_added"] == 0).sum() > 0 and (log["bias_added"] > 0).sum() > 0


def test_round_larger_than_the_parallel_plan(edm, port, tmp_path):
    """hill_density < 0: every candidate deposits.  6 000 hills in one round exceed the plan's capacity
    (in-order kernel, bitonic ordering of the accepted list) and 1 500 fit it."""
    text = ("tempering 0\nhill_prefactor 0.5\nbias_per_step 1000\ndimension 1\nbox_low 0\nbox_high 10\n"
            "bias_spacing 0.01\nbias_sigma 0.05")
    bd, bo = make_both(edm, port, tmp_path, "big", text, 1.0, 1.0, [0.0], [10.0], [1])
    rng = np.random.default_rng(9)
    for n in (6000, 1500):
        x = np.ascontiguousarray(rng.uniform(-1, 11, size=(n, 1)))
        u = rng.uniform(0, 1, n)
        bo.add_hills(x, u, -1)
        bd.add_hills(x, u)
    log = compare_bias(bd, bo)
    assert len(log) == 7500
    info = bd.round_info()
    assert info["in_order"] == 1 and info["parallel"] == 1, info


def test_backlog_overflow_is_reported(edm, port, tmp_path):
    """More buffered hills than BIAS_BUFFER_SIZE: the reference aborts (lib/edm_bias.cpp:503-507); the C ABI
    returns EDM_ERR_BACKLOG_FULL and the C++ mirror turns that into edm_error."""
    text = ("tempering 0\nhill_prefactor 1.0\nbias_per_step 0.0001\ndimension 1\nbox_low 0\nbox_high 10\n"
            "bias_spacing 0.01\nbias_sigma 0.05")
    f = write_edm(tmp_path, "full.edm", text)
    bd = edm.bias_from_edm(f, 1.0, 1.0, [0.0], [10.0], [0.0], [10.0], [1], [0.0])
    rng = np.random.default_rng(1)
    # the first hill fits under bias_per_step, the second overshoots and is pushed, then one push per hill:
    # 2 050 candidates are one more than the deque takes
    x = np.ascontiguousarray(rng.uniform(0, 10, size=(2050, 1)))
    with pytest.raises(edm.EdmError) as err:
        bd.add_hills(x, rng.uniform(0, 1, 2050))
    assert "overflow buffer is full" in str(err.value)


def test_backlog_fills_to_the_brim_without_overflow(edm, port, tmp_path):
    """Exactly as many pushes as the deque takes (the parallel tail push must stop at the same slot)."""
    text = ("tempering 0\nhill_prefactor 1.0\nbias_per_step 0.0001\ndimension 1\nbox_low 0\nbox_high 10\n"
            "bias_spacing 0.01\nbias_sigma 0.05")
    bd, bo = make_both(edm, port, tmp_path, "brim", text, 1.0, 1.0, [0.0], [10.0], [1])
    rng = np.random.default_rng(2)
    for n in (2049, 400, 300):     # 2 048 pushes fill the deque exactly; later rounds are skipped while it drains
        x = np.ascontiguousarray(rng.uniform(0, 10, size=(n, 1)))
        u = rng.uniform(0, 1, n)
        bo.add_hills(x, u, -1)
        bd.add_hills(x, u)
    compare_bias(bd, bo)
    assert bd.backlog()[1] == 2048


def test_deposit_empty_and_single(edm, port):
    gd = edm.GaussGrid(2, [0.0, 0.0], [4.0, 4.0], [0.0625, 0.0625], [1, 1], 1, [0.125, 0.125])
    go = port.GaussGrid("port", 2, [0.0, 0.0], [4.0, 4.0], [0.0625, 0.0625], [1, 1], 1, [0.125, 0.125])
    assert len(gd.add_values(np.zeros((0, 2)), np.zeros(0))) == 0
    c = np.array([[3.99, 0.01]])
    bd = gd.add_values(c, np.array([0.5]))
    bo = go.add_values(c, np.array([0.5]))
    assert_close(bd, bo, "bias_added")
    assert_close(gd.get_arrays()[0], go.get_arrays()[0], "grid")


def test_poor_estimate_accepts_many_times_hill_density(edm, port, tmp_path):
    """est_hill_count far below the real number of proposals (fix edm_pair's first round starts from atom->nmax,
    lammps/fix_edm_pair.cpp:105): thirty times hill_density hills are accepted and all of them reach the limiter."""
    text = ("tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 250\ndimension 1\nbox_low 1.68\n"
            "box_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025")
    bd, bo = make_both(edm, port, tmp_path, "poor", text, 300.0, 0.0019872, [1.68], [5.0], [0])
    rng = np.random.default_rng(12)
    r = rng.uniform(0.5, 5.5, 60000)
    u = rng.uniform(0, 1, 60000)
    for b in (bo, bd):
        b.pre_add_hill(2000)            # acceptance 250/2000 per proposal -> ~7 500 hills
        b.add_hill_many(r[:25000], u[:25000])
        b.add_hill_many(r[25000:], u[25000:])
        b.post_add_hill()
    log = compare_bias(bd, bo)
    assert 7000 < len(log) < 8000


def test_grid_add_minmax_hist(edm, port):
    """Grid::add (restart via initial_bias_filename, lib/grid.h:275-290, lib/edm_bias.cpp:166-167), max_value /
    min_value (:292-309) and the histogram bump DimmedGrid::add_value (:370-385) as direct calls."""
    rng = np.random.default_rng(41)
    src = edm.GaussGrid(2, [0.0, 0.0], [4.0, 4.0], [0.0625, 0.0625], [1, 1], 1, [0.125, 0.125])
    src.add_values(rng.uniform(0, 4, size=(300, 2)), rng.uniform(0.5, 1.5, 300))
    sv, sd = src.get_arrays()
    dst = edm.GaussGrid(2, [0.0, 0.0], [4.0, 4.0], [0.0625, 0.0625], [1, 1], 1, [0.125, 0.125])
    dst.add_values(rng.uniform(0, 4, size=(50, 2)), rng.uniform(0.5, 1.5, 50))
    dv, dd = dst.get_arrays()
    dst.add(src, 2.0, 0.5)
    v, d = dst.get_arrays()
    assert_close(v, dv + 2.0 * sv + 0.5, "Grid::add values")
    nz = (np.abs(sv) >= 1e-7)[:, None]      # interp drops the tabulated derivative under a near-zero value (T6)
    assert_close(d, dd + 2.0 * sd * nz, "Grid::add derivatives")
    mn, mx = dst.minmax()
    assert mn == v.min() and mx == v.max()
    ho = port.Grid("port", 2, [0.0, -1.0], [4.0, 3.0], [0.25, 0.5], [1, 0], 0, 0)
    hd = edm.Grid(2, [0.0, -1.0], [4.0, 3.0], [0.25, 0.5], [1, 0], 0, 0)
    pts = rng.uniform(-2, 6, size=(5000, 2))
    w = rng.integers(-1, 2, 5000).astype(float)
    ho.hist_add(pts, w)
    hd.hist_add(pts, w)
    assert np.array_equal(hd.get_arrays()[0], ho.get_arrays()[0])


def test_device_pointer_entry_points_match_the_host_ones(edm, port, tmp_path):
    """The _dev variants (device pointers, caller's stream, no synchronisation) against their host-buffer
    twins: grid_eval_dev, gauss_deposit_dev, add_hills_dev, pair_step_cells_dev; plus edm_host_pin and the
    backlog / cum_bias setters used for restart."""
    import ctypes as C
    import torch
    L = edm.lib()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(51)
    # deposit + eval
    ga = edm.GaussGrid(2, [0.0, 0.0], [4.0, 4.0], [0.0625, 0.0625], [1, 1], 1, [0.125, 0.125])
    gb = edm.GaussGrid(2, [0.0, 0.0], [4.0, 4.0], [0.0625, 0.0625], [1, 1], 1, [0.125, 0.125])
    c, h = rng.uniform(0, 4, size=(400, 2)), rng.uniform(0.5, 1.5, 400)
    ba_host = ga.add_values(c, h)
    ct, ht, bt = torch.from_numpy(c).cuda(), torch.from_numpy(h).cuda(), torch.zeros(400, dtype=torch.float64, device="cuda")
    edm.check(L.edm_gauss_deposit_dev(gb.h, 400, ct.data_ptr(), ht.data_ptr(), bt.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(bt.cpu().numpy(), ba_host)
    assert_close(gb.get_arrays()[0], ga.get_arrays()[0], "deposit_dev grid")   # REDs: order of overlapping adds differs
    pts = rng.uniform(-1, 5, size=(3000, 2))
    vh, dh = ga.eval(pts)
    pt = torch.from_numpy(pts).cuda()
    vt, dt = torch.zeros(3000, dtype=torch.float64, device="cuda"), torch.zeros((3000, 2), dtype=torch.float64, device="cuda")
    edm.check(L.edm_grid_eval_dev(ga.h, 3000, pt.data_ptr(), 2, vt.data_ptr(), dt.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(vt.cpu().numpy(), vh) and np.array_equal(dt.cpu().numpy(), dh)
    # pair step on device buffers, host arrays page-locked through the library
    cfg = BIAS_CASES["c2_rdf_threshold_tempering"]
    b1, _ = make_both(edm, port, tmp_path, "dev1", cfg["text"], cfg["T"], cfg["kB"], [1.68], [5.0], [0])
    b2, _ = make_both(edm, port, tmp_path, "dev2", cfg["text"], cfg["T"], cfg["kB"], [1.68], [5.0], [0])
    n, box = 5000, np.array([34.0, 34.0, 34.0])
    for step in range(3):
        x = np.ascontiguousarray(rng.uniform(0, 34.0, size=(n, 3)))
        f1 = np.zeros((n, 3))
        edm.check(L.edm_host_pin(x.ctypes.data, x.nbytes))
        r1 = b1.pair_step_cells(x, f1, box, 5.0, do_hills=True, est=300000, seed=3, step=step)
        edm.check(L.edm_host_unpin(x.ctypes.data))
        xt, ft = torch.from_numpy(x).cuda(), torch.zeros((n, 3), dtype=torch.float64, device="cuda")
        r2 = edm.PairResult()
        edm.check(L.edm_pair_step_cells_dev(b2.h, n, xt.data_ptr(), ft.data_ptr(), None, 0, 0,
                                            box.ctypes.data_as(C.POINTER(C.c_double)), 5.0, 1, 300000, 3, step, C.byref(r2), st))
        torch.cuda.synchronize()
        assert r2.n_pairs == r1["n_pairs"] and r2.n_calls == r1["n_calls"]
        # forces and energy are fixed-point sums: bit-identical whatever order the candidate chunks were handed out in
        assert r2.energy == r1["energy"]
        assert np.array_equal(ft.cpu().numpy(), f1)
    l1, l2 = b1.log(), b2.log()
    assert len(l1) > 0 and np.array_equal(l1["pos"], l2["pos"]) and np.array_equal(l1["height"], l2["height"])
    # restart helpers: cum_bias and the backlog travel
    b1.set_cum_bias(0.125)
    assert b1.state()["cum_bias"] == 0.125
    buf = np.zeros(8192)
    buf[2:8] = [2.0, 1e-5, 3.0, 2e-5, 4.0, 3e-5]     # slots 1..3 of a 1-D backlog
    b1.set_backlog(1, 4, buf)
    le, ri, got = b1.backlog()
    assert (le, ri) == (1, 4) and np.array_equal(got[:8], buf[:8])


def test_pair_step_on_a_periodic_cv_grid_takes_the_general_interpolation(edm, port, tmp_path):
    """The pair kernels' lean 1-D interpolation applies to the grid fix edm_pair sets up (non-periodic grid and
    boundary); any other grid goes through the general routine.  A periodic CV axis exercises that branch in the
    cell path and in the list path."""
    text = ("tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 200\ndimension 1\nbox_low 0\nbox_high 6\n"
            "bias_spacing 0.001\nbias_sigma 0.05")
    bd, bo = make_both(edm, port, tmp_path, "pcv", text, 300.0, 0.0019872, [0.0], [6.0], [1])
    rng = np.random.default_rng(61)
    n, L, rc = 4000, 32.0, 5.0
    for step in range(3):
        x = np.ascontiguousarray(rng.uniform(0, L, size=(n, 3)))
        pi, pj, sh = port.build_half_list(x, [L, L, L], rc)
        u = port.pair_uniforms(5, step, pi, pj, n)
        fo, fd = np.zeros((n, 3)), np.zeros((n, 3))
        eo, _ = bo.pair_step(pi, pj, x, fo, shift=sh, do_hills=True, est=2 * pi.size, uniforms=u)
        res = bd.pair_step_cells(x, fd, [L, L, L], rc, do_hills=True, est=2 * pi.size, seed=5, step=step)
        assert res["n_pairs"] == pi.size
        if step:
            assert abs(res["energy"] - eo) <= RTOL * abs(eo)
            assert_close(fd, fo, "forces on a periodic CV grid, step %d" % step)
    compare_bias(bd, bo)
    # the list path on the same bias: plain pairs only (a list carries no image shifts)
    keep = np.all(sh == 0.0, axis=1)
    pi, pj = pi[keep], pj[keep]
    first = np.zeros(n + 1, np.int64)
    np.add.at(first, pi + 1, 1)
    first = np.cumsum(first)
    fo, fd = np.zeros((n, 3)), np.zeros((n, 3))
    eo, _ = bo.pair_step(pi, pj, x, fo, do_hills=False)
    res = bd.pair_step_list(x, fd, n, np.arange(n, dtype=np.int32), first, pj, do_hills=False)
    assert abs(res["energy"] - eo) <= RTOL * abs(eo)
    assert_close(fd, fo, "list forces on a periodic CV grid")


@pytest.mark.parametrize("dim", [2, 3])
def test_interpolation_with_extreme_and_near_zero_corner_values(edm, port, dim):
    """K1's blend takes ONE reciprocal per four corners (product / product).  Corner values under the reference's
    1e-7 zero threshold (T6, lib/grid.h:113) must drop out exactly as they do there, and values whose product leaves
    the normal range (1e75^4) must take the per-corner division: the result stays the oracle's either way."""
    rng = np.random.default_rng(100 + dim)
    lo, hi, dx = [0.0] * dim, [4.0] * dim, [0.25] * dim
    per = [1] + [0] * (dim - 1)
    gd = edm.Grid(dim, lo, hi, dx, per, 1, 1)
    go = port.Grid("port", dim, lo, hi, dx, per, 1, 1)
    v0, d0 = go.get_arrays()
    n = v0.size
    mags = np.array([1e-9, 5e-8, 1.0000001e-7, 1e-6, 1e-3, 1.0, 1e3, 1e60, 1e75, 1e80])
    v = mags[rng.integers(0, mags.size, n)] * rng.choice([-1.0, 1.0], n) * rng.uniform(1.0, 2.0, n)
    d = (rng.uniform(-1, 1, size=d0.shape) * np.abs(v).reshape(-1, 1)).reshape(d0.shape)
    go.set_arrays(v, d)
    gd.set_arrays(v, d)
    x = rng.uniform(0.0, 3.999, size=(40000, dim))
    val_d, der_d = gd.eval(x)
    val_o, der_o = go.eval(x)
    assert np.array_equal(val_d == 0.0, val_o == 0.0)
    # cells mix magnitudes 80 orders apart, so the yardstick is per point: the largest corner of the point's own cell
    nn = gd.info()["n"]
    idx = np.floor(x / 0.25).astype(np.int64)
    stride = np.concatenate([[1], np.cumprod(nn[:-1])]).astype(np.int64)
    corner_mag = np.zeros(x.shape[0])
    mag = np.abs(v) + np.abs(d).max(axis=1) * 0.25
    for c in range(1 << dim):
        lin = np.zeros(x.shape[0], np.int64)
        for k in range(dim):
            i = idx[:, k] + ((c >> k) & 1)
            if per[k]:
                i = i % nn[k]
            lin += i * stride[k]
        corner_mag = np.maximum(corner_mag, mag[lin])
    assert np.all(np.abs(val_d - val_o) <= 1e-10 * corner_mag)
    assert np.all(np.abs(der_d - der_o).max(axis=1) <= 1e-10 * corner_mag / 0.25 * 8)


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_points_on_grid_lines_land_in_the_reference_cell(edm, port, dim):
    """The cell index is floor((x - min) / dx) with a true division in the reference (lib/grid.h:315-325, T5).  The
    device multiplies by 1/dx and only divides when the product is within rounding of an integer: points exactly on
    grid lines, one ulp either side, and just inside the upper edge are where that matters."""
    rng = np.random.default_rng(7 + dim)
    lo, hi = [0.3] * dim, [3.5] * dim
    dx = [0.1, 0.05, 0.2][:dim]
    per = [0] * dim
    gd = edm.Grid(dim, lo, hi, dx, per, 1, 1)
    go = port.Grid("port", dim, lo, hi, dx, per, 1, 1)
    v0, d0 = go.get_arrays()
    v = rng.uniform(0.5, 2.0, v0.size)
    d = rng.uniform(-1, 1, size=d0.shape)
    go.set_arrays(v, d)
    gd.set_arrays(v, d)
    k = rng.integers(0, 30, size=(6000, dim))
    x = np.array(lo) + k * np.array(dx)
    x[2000:4000] = np.nextafter(x[2000:4000], 10.0)
    x[4000:] = np.nextafter(x[4000:], -10.0)
    x = np.vstack([x, np.nextafter(np.array(hi), -10.0)[None, :], np.array(lo)[None, :]])
    val_d, der_d = gd.eval(x)
    val_o, der_o = go.eval(x)
    assert np.array_equal(val_d == 0.0, val_o == 0.0)
    assert_close(val_d, val_o, "value on grid lines")
    assert_close(der_d, der_o, "derivative on grid lines")
