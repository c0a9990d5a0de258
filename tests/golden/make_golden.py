"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libedm_ref.so, built by
`make -C oracle ref` from /root/reference/lib).  Run here, in the container that has /root/reference;
the fixtures travel with the repo so the oracle and the CUDA path can be checked where the
reference is absent.

    python tests/golden/make_golden.py

Every array is an input or an output of the reference on that input; nothing is hand-typed.
The three PLUMED grid fixtures of the reference's own tests (tests/1.grid, 2.grid, 3.grid) are read
THROUGH the reference's reader and stored as arrays (geometry + values), so that the known answer
edm_test.cpp:117-125 can be re-checked without the text files.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402

REF_TESTS = "/root/reference/tests"


def grid_case(name, dim, mn, mx, sp, per, sg, bnd, nh, seed, pad=0.15):
    rng = np.random.default_rng(seed)
    g = po.GaussGrid("ref", dim, mn, mx, sp, per, 1, sg)
    if bnd is not None:
        g.set_boundary(*bnd)
    lo = np.array(bnd[0] if bnd else mn, float)
    hi = np.array(bnd[1] if bnd else mx, float)
    span = hi - lo
    c = rng.uniform(lo - pad * span, hi + pad * span, size=(nh, dim))
    c[0], c[1] = lo, hi
    h = rng.uniform(0.5, 1.5, nh)
    h[2] = -0.25
    ba = g.add_values(c, h)
    v, d = g.get_arrays()
    x = rng.uniform(lo - 0.3 * span, hi + 0.3 * span, size=(4000, dim))
    val, der = g.eval(x)
    gv = g.get_value(x)
    info = g.info()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), dim=dim, min=mn, max=mx, spacing=sp, periodic=per, sigma=sg,
                        has_boundary=bnd is not None, bmin=lo, bmax=hi, bper=np.array(bnd[2] if bnd else per),
                        centres=c, heights=h, bias_added=ba, grid=v, deriv=d, x=x, value=val, der=der, get_value=gv,
                        n=info["n"], dx=info["dx"], gmax=info["max"], minisize=info["minisize"])
    print(name, "hills", nh, "points", v.size)


def bias_case(name, text, T, kB, sub, periodic, skin, n, lo, hi, steps, seed, target=None):
    """target = (min, max, spacing, periodic, values): written as a PLUMED grid by the reference's own writer and named
    by target_filename, so the reference reads it back itself (8 decimals); the fixture keeps the values as read."""
    tmp = tempfile.mkdtemp()
    f = os.path.join(tmp, name + ".edm")
    extra = {}
    full_text = text
    if target is not None:
        tmn, tmx, tsp, tper, tvals = target
        D = len(tmn)
        tg = po.Grid("ref", D, tmn, tmx, tsp, tper, 0, 0)
        tg.set_arrays(tvals)
        tfile = os.path.join(tmp, "target.grid")
        tg.write(tfile)
        back = po.Grid("ref", dim=D, filename=tfile, b_interp=0)
        extra = dict(target_min=tmn, target_max=tmx, target_spacing=tsp, target_periodic=tper,
                     target_values=back.get_arrays()[0], target_n=back.info()["n"], target_dx=back.info()["dx"])
        full_text = text + "\ntarget_filename " + tfile
    open(f, "w").write(full_text + "\nhills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (tmp, tmp))
    b = po.Bias("ref", f)
    b.setup(T, kB)
    b.subdivide(sub[0], sub[1], sub[0], sub[1], periodic, skin)
    D = b.dim
    rng = np.random.default_rng(seed)
    xs, us, es, fs = [], [], [], []
    for _ in range(steps):
        x = np.ascontiguousarray(rng.uniform(lo, hi, size=(n, D)))
        u = rng.uniform(0, 1, n)
        force = np.zeros((n, D))
        es.append(b.update_forces(x, force))
        fs.append(force)
        b.add_hills(x, u)
        xs.append(x)
        us.append(u)
    log = b.log()
    left, right, buf = b.backlog()
    v, d = b.gauss.get_arrays()
    hist = b.hist.get_arrays()[0]
    p = b.params()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), edm_text=text, T=T, kB=kB, sublo=sub[0], subhi=sub[1],
                        periodic=periodic, skin=skin, x=np.array(xs), u=np.array(us), energy=np.array(es),
                        forces=np.array(fs), log_steps=log["steps"], log_type=log["type"],
                        log_hills_added=log["hills_added"], log_pos=log["pos"], log_height=log["height"],
                        log_bias_added=log["bias_added"], log_cum=log["cum_over_vol"], backlog_left=left,
                        backlog_right=right, backlog=buf, grid=v, deriv=d, hist=hist, cum_bias=p["cum_bias"],
                        total_volume=p["total_volume"], steps=p["steps"], expected_target=p["expected_target"], **extra)
    print(name, "events", len(log), "backlog", left, right, "expected_target", p["expected_target"])


def plumed_grids():
    out = {}
    for dim in (1, 2, 3):
        g = po.Grid("ref", dim=dim, filename=os.path.join(REF_TESTS, "%d.grid" % dim), b_interp=1)
        info = g.info()
        v, d = g.get_arrays()
        out["n%d" % dim] = info["n"]
        out["dx%d" % dim] = info["dx"]
        out["min%d" % dim] = info["min"]
        out["max%d" % dim] = info["max"]
        out["periodic%d" % dim] = info["periodic"]
        out["b_derivatives%d" % dim] = info["b_derivatives"]
        out["grid%d" % dim] = v
        out["deriv%d" % dim] = d
    g3 = po.Grid("ref", dim=3, filename=os.path.join(REF_TESTS, "3.grid"), b_interp=0)
    out["known_x"] = np.array([0.75, 0.0, 1.00])          # edm_test.cpp:122
    out["known_value_nointerp"] = g3.get_value(out["known_x"])
    g3.set_interpolation(1)
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.2, 2.7, size=(2000, 3))
    out["x3"] = x
    out["value3"], out["der3"] = g3.eval(x)
    np.savez_compressed(os.path.join(HERE, "plumed_grids.npz"), **out)
    print("plumed grids", out["known_value_nointerp"])


if __name__ == "__main__":
    grid_case("gauss_1d_rdf_mcgdp", 1, [1.68], [5.0], [0.00025], [0], [0.025], None, 120, 1)
    grid_case("gauss_1d_periodic", 1, [0.0], [10.0], [10.0 / 1024], [1], [0.025], None, 200, 2)
    grid_case("gauss_1d_inner_mcgdp", 1, [-100.0], [100.0], [1.0], [1], [10.0], ([-50.0], [50.0], [0]), 40, 3)
    grid_case("gauss_2d_mixed", 2, [0.0, 0.0], [10.0, 5.0], [0.1, 0.13], [1, 0], [0.3, 0.25],
              ([0.0, 0.0], [10.0, 10.0], [1, 0]), 150, 4)
    grid_case("gauss_2d_mcgdp", 2, [0.0, 1.0], [4.0, 3.0], [0.05, 0.04], [0, 0], [0.2, 0.15], None, 100, 5)
    grid_case("gauss_3d_inner_mcgdp", 3, [-10.0] * 3, [10.0] * 3, [0.9, 1.1, 1.4], [1, 1, 1], [3.0, 3.0, 3.0],
              ([-5.0] * 3, [5.0] * 3, [0, 0, 0]), 30, 6)
    bias_case("bias_c5_tight_limiter",
              "tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.0002\n"
              "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025",
              300.0, 0.0019872, ([1.68], [5.0]), [0], [0.0], 8000, 0.5, 5.5, 10, 7)
    bias_case("bias_1d_local_tempering",
              "tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.5\nbias_per_step 1000\n"
              "hill_density 60\ndimension 1\nbox_low 0\nbox_high 10\nbias_spacing 0.01\nbias_sigma 0.1",
              300.0, 0.0019872, ([0.0], [10.0]), [1], [0.0], 2000, 0.0, 10.0, 4, 8)
    bias_case("bias_2d_local_tempering",
              "tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\n"
              "hill_density 40\ndimension 2\nbox_low 0 0\nbox_high 8 8\nbias_spacing 0.125 0.125\nbias_sigma 0.25 0.25",
              300.0, 0.0019872, ([0.0, 0.0], [8.0, 8.0]), [1, 1], [0.0, 0.0], 1500, 0.0, 8.0, 3, 9)
    bias_case("bias_1d_targeting",
              "tempering 1\nglobal_tempering 0.00002\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\n"
              "hill_density 100\ndimension 1\nbox_low 0\nbox_high 10\nbias_spacing 0.01\nbias_sigma 0.05",
              300.0, 0.0019872, ([0.0], [10.0]), [1], [0.0], 4000, -1.0, 11.0, 4, 10,
              target=([0.0], [10.0], [0.1], [1], 1.5 + np.sin(np.linspace(0, 9, 100)) + 0.05 * np.cos(np.arange(100.0))))
    bias_case("bias_3d_density",
              "tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\ndimension 3\nbox_low 0 0 0\nbox_high 8 8 8\n"
              "bias_spacing 0.25 0.25 0.25\nbias_sigma 0.5 0.5 0.5\nhill_density 60",
              1.0, 1.0, ([0.0] * 3, [8.0] * 3), [1, 1, 1], [0.0] * 3, 3000, -1.0, 9.0, 3, 11)
    bias_case("bias_2d_walls_threshold",
              "tempering 1\nglobal_tempering 0.00001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.006\n"
              "hill_density 80\ndimension 2\nbox_low 0 0\nbox_high 4 4\nbias_spacing 0.0625 0.0625\n"
              "bias_sigma 0.125 0.125",
              300.0, 0.0019872, ([0.0, 0.0], [4.0, 4.0]), [0, 0], [0.0, 0.0], 2500, -0.5, 4.5, 5, 12)
    plumed_grids()
