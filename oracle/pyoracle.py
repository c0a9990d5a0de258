"""ctypes front-end to the CPU oracles.  TEST INFRASTRUCTURE ONLY.

Two interchangeable back-ends behind one Python surface:

* ``load("port")``  -> oracle/libedm_oracle.so, the C restatement (oracle/edm_oracle.c)
* ``load("ref")``   -> oracle/_ref/libedm_ref.so, the UNMODIFIED reference lib/ compiled by
                       oracle/Makefile plus the C shim oracle/ref_shim.cpp

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (electronic-dance-music_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libedm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libedm_ref.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _d(a, n=None):
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))
    return a


def _i(a):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.int32)))


HILL_EVENT = np.dtype(
    [("steps", "<i8"), ("type", "<i4"), ("hills_added", "<i4"), ("pos", "<f8", (3,)), ("height", "<f8"),
     ("bias_added", "<f8"), ("cum_over_vol", "<f8")],
    align=True,
)


def build(which=("port", "ref")):
    """Compile the oracles (checker build; called by __graft_entry__.build())."""
    for w in which:
        subprocess.check_call(["make", "-s", "-C", HERE, w])


def available(kind):
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


class _Lib:
    def __init__(self, kind):
        self.kind = kind
        self.prefix = "orc_" if kind == "port" else "ref_"
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        self.lib = C.CDLL(path)
        vp, d, i, l = C.c_void_p, C.c_double, C.c_int, C.c_long
        sig = {
            "grid_create": (vp, [i, c_dp, c_dp, c_dp, c_ip, i, i]),
            "grid_destroy": (None, [vp]),
            "grid_info": (None, [vp, c_ip, c_dp, c_dp, c_dp, c_ip]),
            "grid_size": (C.c_size_t, [vp]),
            "grid_get_arrays": (None, [vp, c_dp, c_dp]),
            "grid_set_arrays": (None, [vp, c_dp, c_dp]),
            "grid_set_interpolation": (None, [vp, i]),
            "grid_eval": (None, [vp, l, c_dp, c_dp, c_dp]),
            "grid_get_value": (None, [vp, l, c_dp, c_dp]),
            "grid_hist_add": (None, [vp, l, c_dp, c_dp]),
            "grid_expected_bias": (d, [vp]),
            "gauss_create": (vp, [i, c_dp, c_dp, c_dp, c_ip, i, c_dp]),
            "gauss_destroy": (None, [vp]),
            "gauss_set_boundary": (None, [vp, c_dp, c_dp, c_ip]),
            "gauss_info": (None, [vp, c_ip, c_dp, c_dp, c_dp, c_ip]),
            "gauss_size": (C.c_size_t, [vp]),
            "gauss_get_arrays": (None, [vp, c_dp, c_dp]),
            "gauss_set_arrays": (None, [vp, c_dp, c_dp]),
            "gauss_tables": (None, [vp, i, c_dp, c_dp]),
            "gauss_add_value": (d, [vp, c_dp, d]),
            "gauss_add_values": (None, [vp, l, c_dp, c_dp, c_dp]),
            "gauss_eval": (None, [vp, l, c_dp, c_dp, c_dp]),
            "gauss_get_value": (None, [vp, l, c_dp, c_dp]),
            "gauss_remap": (None, [vp, c_dp]),
            "gauss_set_interpolation": (None, [vp, i]),
            "bias_destroy": (None, [vp]),
            "bias_setup": (None, [vp, d, d]),
            "bias_subdivide": (None, [vp, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp]),
            "bias_gauss": (vp, [vp]),
            "bias_hist": (vp, [vp]),
            "bias_params": (None, [vp, c_dp]),
            "bias_set_cum_bias": (None, [vp, d]),
            "bias_backlog": (None, [vp, C.POINTER(l), C.POINTER(l), c_dp]),
            "bias_set_mask": (None, [vp, c_ip]),
            "bias_update_forces": (d, [vp, l, c_dp, l, c_dp, l, i]),
            "bias_add_hills": (None, [vp, l, c_dp, l, c_dp, i]),
            "bias_pre_add_hill": (None, [vp, i]),
            "bias_add_hill_many": (None, [vp, l, c_dp, c_dp]),
            "bias_post_add_hill": (None, [vp]),
            "pair_step": (d, [vp, l, c_ip, c_ip, c_dp, c_dp, c_dp, i, i, c_dp, c_dp]),
            "pair_step_ghost": (d, [vp, l, c_ip, c_ip, c_dp, c_dp, c_dp, l, i, i, c_dp, c_dp, C.POINTER(l)]),
            "time_pair_eval": (d, [vp, l, c_dp, i]),
            "time_add_values": (d, [vp, l, c_dp, c_dp]),
            "time_fix_pair": (d, [vp, l, c_ip, c_ip, C.POINTER(C.c_byte), c_dp, c_dp, c_dp, i, i, C.c_ulonglong,
                                  C.c_ulonglong, c_dp, C.POINTER(l)]),
        }
        if kind == "port":
            sig.update({
                "bias_create": (vp, [i, i, d, d, d, d, d, c_dp, c_dp, c_dp, c_dp]),
                "bias_set_target": (None, [vp, vp, d]),
                "bias_log_size": (l, [vp]),
                "bias_log_copy": (None, [vp, vp]),
                "bias_log_clear": (None, [vp]),
                "bias_log_enable": (None, [vp, i]),
                "build_half_list": (l, [l, c_dp, c_dp, d, l, c_ip, c_ip, c_dp]),
                "half_list_build": (vp, [l, c_dp, c_dp, d]),
                "half_list_size": (l, [vp]),
                "half_list_copy": (None, [vp, c_ip, c_ip, C.POINTER(C.c_byte)]),
                "half_list_free": (None, [vp]),
                "uniform": (d, [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong]),
                "uniform_fill": (None, [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, l, c_dp]),
                "uniform_pair": (d, [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, i]),
                "uniform_pair_fill": (None, [C.c_ulonglong, C.c_ulonglong, l, C.POINTER(C.c_ulonglong), c_dp]),
            })
        else:
            sig.update({
                "bias_create": (vp, [C.c_char_p]),
                "grid_read": (vp, [i, C.c_char_p, i]),
                "grid_write": (None, [vp, C.c_char_p]),
                "gauss_write": (None, [vp, C.c_char_p]),
                "bias_arrays": (None, [vp, c_dp, c_dp, c_dp, c_dp]),
                "bias_flush_log": (None, [vp]),
                "bias_write_bias": (None, [vp, C.c_char_p]),
                "bias_write_histogram": (None, [vp]),
                "bias_clear_histogram": (None, [vp]),
            })
        for name, (res, args) in sig.items():
            fn = getattr(self.lib, self.prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)


_LIBS = {}


def load(kind="port"):
    if kind not in _LIBS:
        _LIBS[kind] = _Lib(kind)
    return _LIBS[kind]


class _GridBase:
    """Shared surface of Grid (lib/grid.h:185) and GaussGrid (lib/gaussian_grid.h:59)."""

    _pfx = "grid"

    def _f(self, name):
        return getattr(self.L, self._pfx + "_" + name)

    @property
    def size(self):
        return int(self._f("size")(self.h))

    def get_arrays(self):
        v = np.zeros(self.size)
        d = np.zeros(self.size * self.dim)
        self._f("get_arrays")(self.h, _dp(v), _dp(d))
        return v, d.reshape(self.size, self.dim)

    def set_arrays(self, v, d=None):
        v = _d(v)
        d = _d(d).ravel() if d is not None else np.zeros(self.size * self.dim)
        assert v.size == self.size and d.size == self.size * self.dim
        self._f("set_arrays")(self.h, _dp(v), _dp(d))

    def set_interpolation(self, b):
        self._f("set_interpolation")(self.h, int(b))

    def eval(self, x):
        x = _d(x).reshape(-1, self.dim)
        n = x.shape[0]
        val = np.zeros(n)
        der = np.zeros((n, self.dim))
        self._f("eval")(self.h, n, _dp(x), _dp(val), _dp(der))
        return val, der

    def get_value(self, x):
        x = _d(x).reshape(-1, self.dim)
        n = x.shape[0]
        val = np.zeros(n)
        self._f("get_value")(self.h, n, _dp(x), _dp(val))
        return val


class Grid(_GridBase):
    def __init__(self, kind, dim=None, mn=None, mx=None, spacing=None, periodic=None, b_deriv=0, b_interp=0,
                 handle=None, filename=None):
        self.L = load(kind)
        self.owned = handle is None
        if handle is not None:
            self.h = handle
            self.dim = dim
        elif filename is not None:
            self.dim = dim
            self.h = self.L.grid_read(dim, filename.encode(), int(b_interp))
        else:
            self.dim = dim
            self.h = self.L.grid_create(dim, _dp(_d(mn)), _dp(_d(mx)), _dp(_d(spacing)), _ip(_i(periodic)),
                                        int(b_deriv), int(b_interp))

    def info(self):
        n = np.zeros(3, np.int32)
        dx, mn, mx = np.zeros(3), np.zeros(3), np.zeros(3)
        flags = np.zeros(5, np.int32)
        self.L.grid_info(self.h, _ip(n), _dp(dx), _dp(mn), _dp(mx), _ip(flags))
        D = self.dim
        return dict(n=n[:D].copy(), dx=dx[:D].copy(), min=mn[:D].copy(), max=mx[:D].copy(),
                    b_derivatives=int(flags[0]), b_interpolate=int(flags[1]), periodic=flags[2:2 + D].copy())

    def hist_add(self, x, v):
        x = _d(x).reshape(-1, self.dim)
        v = _d(v)
        self.L.grid_hist_add(self.h, x.shape[0], _dp(x), _dp(v))

    def expected_bias(self):
        return float(self.L.grid_expected_bias(self.h))

    def write(self, fn):
        self.L.grid_write(self.h, fn.encode())

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            self.L.grid_destroy(self.h)
            self.h = None


class GaussGrid(_GridBase):
    _pfx = "gauss"

    def __init__(self, kind, dim=None, mn=None, mx=None, spacing=None, periodic=None, interp=1, sigma=None,
                 handle=None):
        self.L = load(kind)
        self.dim = dim
        self.owned = handle is None
        if handle is not None:
            self.h = handle
        else:
            self.h = self.L.gauss_create(dim, _dp(_d(mn)), _dp(_d(mx)), _dp(_d(spacing)), _ip(_i(periodic)),
                                         int(interp), _dp(_d(sigma)))

    def set_boundary(self, mn, mx, periodic):
        self.L.gauss_set_boundary(self.h, _dp(_d(mn)), _dp(_d(mx)), _ip(_i(periodic)))

    def info(self):
        n = np.zeros(3, np.int32)
        mini = np.zeros(3, np.int32)
        dx, mn, mx = np.zeros(3), np.zeros(3), np.zeros(3)
        self.L.gauss_info(self.h, _ip(n), _dp(dx), _dp(mn), _dp(mx), _ip(mini))
        D = self.dim
        return dict(n=n[:D].copy(), dx=dx[:D].copy(), min=mn[:D].copy(), max=mx[:D].copy(),
                    minisize=mini[:D].copy())

    def tables(self, dimi):
        a, b = np.zeros(65536), np.zeros(65536)
        self.L.gauss_tables(self.h, dimi, _dp(a), _dp(b))
        return a, b

    def add_value(self, x, h):
        return float(self.L.gauss_add_value(self.h, _dp(_d(x)), float(h)))

    def add_values(self, x, h):
        x = _d(x).reshape(-1, self.dim)
        h = _d(h)
        ba = np.zeros(x.shape[0])
        self.L.gauss_add_values(self.h, x.shape[0], _dp(x), _dp(h), _dp(ba))
        return ba

    def remap(self, x):
        x = _d(x).copy()
        self.L.gauss_remap(self.h, _dp(x))
        return x

    def time_add_values(self, x, h):
        x = _d(x).reshape(-1, self.dim)
        h = _d(h)
        return float(self.L.time_add_values(self.h, x.shape[0], _dp(x), _dp(h)))

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            self.L.gauss_destroy(self.h)
            self.h = None


PARAM_NAMES = ["dim", "b_tempering", "b_targeting", "global_tempering", "bias_factor", "boltzmann_factor",
               "hill_prefactor", "bias_per_step", "hill_density", "cum_bias", "total_volume", "expected_target",
               "b_outofbounds", "steps"]


def parse_edm_text(text):
    """Minimal reader of an edm input file for driving the PORT (the product's parser is separate
    and is tested against the reference's, tests/test_host_api.py)."""
    kv = {}
    for line in text.splitlines():
        p = line.split()
        if p and p[0] not in kv:
            kv[p[0]] = p[1:]
    return kv


class Bias:
    """EDMBias (lib/edm_bias.h:29).  kind="ref" is built from an edm file; kind="port" takes the same
    file and feeds the parsed numbers to orc_bias_create."""

    def __init__(self, kind, edm_file):
        self.L = load(kind)
        self.kind = kind
        self._keep = []
        text = open(edm_file).read()
        kv = parse_edm_text(text)
        self.dim = int(kv["dimension"][0])
        self.hills_file = None
        if kind == "ref":
            self.h = self.L.bias_create(edm_file.encode())
            hf = kv.get("hills_filename", ["HILLS"])[0]
            self.hills_file = hf + "_0"
        else:
            D = self.dim
            temp = int(kv["tempering"][0])
            gt = float(kv.get("global_tempering", [0])[0]) if temp else 0.0
            bf = float(kv.get("bias_factor", [0])[0]) if temp else 0.0
            pref = float(kv["hill_prefactor"][0])
            bps = float(kv["bias_per_step"][0]) if "bias_per_step" in kv else pref
            dens = float(kv["hill_density"][0]) if "hill_density" in kv else -1.0
            g = lambda k: _d([float(v) for v in kv[k][:D]])
            self.h = self.L.bias_create(D, temp, gt, bf, pref, bps, dens, _dp(g("bias_spacing")), _dp(g("bias_sigma")),
                                        _dp(g("box_low")), _dp(g("box_high")))
            self.target = None
            if "target_filename" in kv:
                raise NotImplementedError("port: pass a target with set_target()")

    def set_target(self, grid):
        assert self.kind == "port"
        self.target = grid
        self.L.bias_set_target(self.h, grid.h, grid.expected_bias())

    def setup(self, T, kB):
        self.L.bias_setup(self.h, float(T), float(kB))

    def subdivide(self, sublo, subhi, boxlo, boxhi, periodic, skin):
        p3 = lambda a: _d(list(np.atleast_1d(a)) + [0.0] * (3 - len(np.atleast_1d(a))))
        per = _i(list(np.atleast_1d(periodic)) + [0] * (3 - len(np.atleast_1d(periodic))))
        self.L.bias_subdivide(self.h, _dp(p3(sublo)), _dp(p3(subhi)), _dp(p3(boxlo)), _dp(p3(boxhi)), _ip(per),
                              _dp(p3(skin)))

    @property
    def gauss(self):
        return GaussGrid(self.kind, dim=self.dim, handle=self.L.bias_gauss(self.h))

    @property
    def hist(self):
        return Grid(self.kind, dim=self.dim, handle=self.L.bias_hist(self.h))

    def params(self):
        out = np.zeros(14)
        self.L.bias_params(self.h, _dp(out))
        return dict(zip(PARAM_NAMES, out.tolist()))

    def set_cum_bias(self, v):
        self.L.bias_set_cum_bias(self.h, float(v))

    def backlog(self):
        l, r = C.c_long(0), C.c_long(0)
        buf = np.zeros(8192)
        self.L.bias_backlog(self.h, C.byref(l), C.byref(r), _dp(buf))
        return int(l.value), int(r.value), buf

    def set_mask(self, mask):
        m = _i(mask)
        self._keep.append(m)
        self.L.bias_set_mask(self.h, _ip(m))

    def update_forces(self, x, f, apply_mask=-1):
        """x, f: (n, stride) float64 arrays; f is updated in place.  Returns the energy."""
        assert x.flags.c_contiguous and f.flags.c_contiguous and f.dtype == np.float64
        n = x.shape[0]
        return float(self.L.bias_update_forces(self.h, n, _dp(x), x.shape[1], _dp(f), f.shape[1], int(apply_mask)))

    def add_hills(self, x, runiform, apply_mask=-1):
        assert x.flags.c_contiguous
        u = _d(runiform)
        self.L.bias_add_hills(self.h, x.shape[0], _dp(x), x.shape[1], _dp(u), int(apply_mask))

    def pre_add_hill(self, est):
        self.L.bias_pre_add_hill(self.h, int(est))

    def add_hill_many(self, x, u):
        x = _d(x).reshape(-1, self.dim)
        u = _d(u)
        self.L.bias_add_hill_many(self.h, x.shape[0], _dp(x), _dp(u))

    def post_add_hill(self):
        self.L.bias_post_add_hill(self.h)

    def pair_step(self, pi, pj, x, f, shift=None, do_hills=False, est=0, uniforms=None):
        pi, pj = _i(pi), _i(pj)
        npairs = pi.size
        r = np.zeros(npairs)
        sh = _dp(_d(shift).ravel()) if shift is not None else None
        un = _dp(_d(uniforms)) if uniforms is not None else None
        e = self.L.pair_step(self.h, npairs, _ip(pi), _ip(pj), _dp(x), _dp(f), sh, int(do_hills), int(est), un, _dp(r))
        return float(e), r

    def pair_step_ghost(self, pi, pj, x, f, nlocal, shift=None, do_hills=False, est=0, uniforms=None):
        """pair_step for one rank of a decomposed system: atoms >= nlocal are ghosts (no force, one proposal)."""
        pi, pj = _i(pi), _i(pj)
        npairs = pi.size
        r = np.zeros(max(npairs, 1))
        nc = C.c_long(0)
        sh = _dp(_d(shift).ravel()) if shift is not None else None
        un = _dp(_d(uniforms)) if uniforms is not None else None
        e = self.L.pair_step_ghost(self.h, npairs, _ip(pi), _ip(pj), _dp(x), _dp(f), sh, int(nlocal), int(do_hills),
                                   int(est), un, _dp(r), C.byref(nc))
        return float(e), r[:npairs], nc.value

    def time_fix_pair(self, pi, pj, img, box, x, f, do_hills, est, seed=0, step=0):
        """The reference's own pair loop (lammps/fix_edm_pair.cpp:173-247, evaluation and hill proposals interleaved)
        over pairs (pi, pj, img) from build_half_list_fast; returns (seconds, energy, hill proposals made)."""
        assert pi.dtype == np.int32 and pj.dtype == np.int32 and img.dtype == np.int8
        assert x.flags.c_contiguous and f.flags.c_contiguous
        e = C.c_double(0)
        nc = C.c_long(0)
        t = self.L.time_fix_pair(self.h, pi.size, _ip(pi), _ip(pj), img.ctypes.data_as(C.POINTER(C.c_byte)), _dp(_d(box)),
                                 _dp(x), _dp(f), int(do_hills), int(est), seed, step, C.byref(e), C.byref(nc))
        return float(t), e.value, nc.value

    def time_pair_eval(self, r, repeats=1):
        r = _d(r)
        return float(self.L.time_pair_eval(self.h, r.size, _dp(r), int(repeats)))

    def log(self):
        """Hill log as a structured array (port) or parsed from the HILLS text file (ref)."""
        if self.kind == "port":
            n = int(self.L.bias_log_size(self.h))
            out = np.zeros(n, dtype=HILL_EVENT)
            if n:
                self.L.bias_log_copy(self.h, out.ctypes.data_as(C.c_void_p))
            return out
        self.L.bias_flush_log(self.h)
        rows = []
        D = self.dim
        with open(self.hills_file) as fh:
            for line in fh:
                p = line.split()
                if not p:
                    continue
                pos = [float(v) for v in p[3:3 + D]] + [0.0] * (3 - D)
                rows.append((int(p[0]), ord(p[1]), int(p[2]), pos, float(p[3 + D]), float(p[4 + D]), float(p[5 + D])))
        return np.array(rows, dtype=HILL_EVENT)

    def log_enable(self, on):
        if self.kind == "port":
            self.L.bias_log_enable(self.h, int(on))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.bias_destroy(self.h)
            self.h = None


def build_half_list(x, box, cutoff):
    """Half neighbour list stand-in (oracle/edm_oracle.c:orc_build_half_list)."""
    L = load("port")
    x = _d(x).reshape(-1, 3)
    box = _d(box)
    n = x.shape[0]
    cnt = L.build_half_list(n, _dp(x), _dp(box), float(cutoff), 0, None, None, None)
    pi = np.zeros(max(cnt, 1), np.int32)
    pj = np.zeros(max(cnt, 1), np.int32)
    sh = np.zeros((max(cnt, 1), 3))
    L.build_half_list(n, _dp(x), _dp(box), float(cutoff), cnt, _ip(pi), _ip(pj), _dp(sh))
    return pi[:cnt], pj[:cnt], sh[:cnt]


def build_half_list_fast(x, box, cutoff):
    """The same pair set as build_half_list for boxes of >= 3 cutoffs per side, fast enough for 10^6 atoms;
    returns (pi, pj, img) with the image shift as int8 codes: separation = x[i] - x[j] - img * box."""
    L = load("port")
    x = _d(x).reshape(-1, 3)
    h = L.half_list_build(x.shape[0], _dp(x), _dp(_d(box)), float(cutoff))
    if not h:
        raise ValueError("box smaller than 3 cutoffs per side")
    n = L.half_list_size(h)
    pi, pj = np.zeros(n, np.int32), np.zeros(n, np.int32)
    img = np.zeros((n, 3), np.int8)
    L.half_list_copy(h, _ip(pi), _ip(pj), img.ctypes.data_as(C.POINTER(C.c_byte)))
    L.half_list_free(h)
    return pi, pj, img


def uniform_fill(seed, step, first, n):
    L = load("port")
    out = np.zeros(n)
    L.uniform_fill(seed, step, first, n, _dp(out))
    return out


def pair_uniforms(seed, step, pi, pj, natoms):
    """The two uniforms of every pair (i < j) as the device draws them: key i*natoms+j (edm_uniform_pair)."""
    L = load("port")
    keys = np.ascontiguousarray(pi.astype(np.uint64) * np.uint64(natoms) + pj.astype(np.uint64))
    out = np.zeros(2 * keys.size)
    L.uniform_pair_fill(seed, step, keys.size, keys.ctypes.data_as(C.POINTER(C.c_ulonglong)), _dp(out))
    return out
