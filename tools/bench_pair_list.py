"""Throughput of the neighbour-list entry point (edm_pair_step_list, what fix edm_pair calls) on a
caller-built half list, host buffers in and out.  Usage: python tools/bench_pair_list.py [natoms]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import edm_b200 as edm  # noqa: E402
import pyoracle  # noqa: E402  (only its numpy neighbour-list builder: the list is the caller's input here)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rc, dens = 5.0, 0.1
L = (n / dens) ** (1.0 / 3.0)
text = ("tempering 1\nglobal_tempering 2.0\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.1\nhill_density 250\n"
        "dimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n")
d = tempfile.mkdtemp()
f = os.path.join(d, "c2.edm")
open(f, "w").write(text + "hills_filename %s/H\nhistogram_filename %s/G\n" % (d, d))
b = edm.bias_from_edm(f, 300.0, 0.0019872, [1.68], [5.0], [1.68], [5.0], [0], [0.0])
rng = np.random.default_rng(0)
b.bias_grid.add_values(rng.uniform(1.68, 5.0, 5000), np.full(5000, 8e-5))
x = np.ascontiguousarray(rng.uniform(0, L, size=(n, 3)))
t0 = time.time()
pi, pj, sh = pyoracle.build_half_list(x, [L, L, L], rc)
print("list built: %d pairs in %.1f s" % (pi.size, time.time() - t0))
# ghosts are not modelled here: keep the pairs that need no image shift (the list form takes raw coordinates)
keep = ~np.any(sh != 0, axis=1) if sh is not None and np.ndim(sh) == 2 else np.ones(pi.size, bool)
pi, pj = pi[keep], pj[keep]
order = np.argsort(pi, kind="stable")
pi, pj = pi[order], pj[order]
ilist = np.unique(pi).astype(np.int32)
counts = np.bincount(pi, minlength=n)[ilist]
first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
jlist = pj.astype(np.int32)
fbuf = np.zeros((n, 3))
t0 = time.perf_counter()
b.pair_list_set(ilist, first, jlist)
print("edm_pair_list_set (upload + row map): %.2f ms" % (1e3 * (time.perf_counter() - t0)))
for do_hills in (False, True):
    ts = []
    for step in range(6):
        t0 = time.perf_counter()
        r = b.pair_step_listed(x, fbuf, n, do_hills=do_hills, est=2 * jlist.size, seed=1, step=step)
        ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    print("list on device, do_hills=%d: %d pairs, %.3f ms per call -> %.3e evals/s (48 MB up, 24 MB down, pageable)" %
          (do_hills, r["n_pairs"], 1e3 * t, r["n_pairs"] / t))
L_ = edm.lib()
edm.check(L_.edm_host_pin(x.ctypes.data, x.nbytes))
edm.check(L_.edm_host_pin(fbuf.ctypes.data, fbuf.nbytes))
ts = []
for step in range(6):
    t0 = time.perf_counter()
    r = b.pair_step_listed(x, fbuf, n, do_hills=True, est=2 * jlist.size, seed=1, step=step)
    ts.append(time.perf_counter() - t0)
t = min(ts[1:])
print("list on device, x and f pinned with edm_host_pin: %.3f ms per call -> %.3e evals/s" % (1e3 * t, r["n_pairs"] / t))
edm.check(L_.edm_host_unpin(x.ctypes.data))
edm.check(L_.edm_host_unpin(fbuf.ctypes.data))
for do_hills in (False, True):
    ts = []
    for step in range(6):
        t0 = time.perf_counter()
        r = b.pair_step_list(x, fbuf, n, ilist, first, jlist, do_hills=do_hills, est=2 * jlist.size, seed=1, step=step)
        ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    print("do_hills=%d: %d pairs, %.3f ms per call -> %.3e evals/s (host buffers; list upload %d MB)" %
          (do_hills, r["n_pairs"], 1e3 * t, r["n_pairs"] / t, (jlist.nbytes + first.nbytes + ilist.nbytes) >> 20))
