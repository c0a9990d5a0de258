"""ctypes binding of libedm_b200.so (include/edm_b200.h) — the test and bench driver.

The product is the CUDA library and the C++ host classes; this module only marshals numpy arrays
(or raw device pointers, e.g. torch tensors' data_ptr()) into the C ABI.  It never computes
anything itself and has no CPU fallback: a missing library or a missing GPU raises EdmError.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(os.path.dirname(HERE))
# EDM_B200_LIB: another build of the same library (A/B measurements of kernel variants, tools/experiments/)
LIB_PATH = os.environ.get("EDM_B200_LIB") or os.path.join(PKG, "lib", "libedm_b200.so")

EDM_BUFFER_DBLS = 8192

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_long)
vp = C.c_void_p


class EdmError(RuntimeError):
    pass


class BiasParams(C.Structure):
    _fields_ = [("dim", C.c_int), ("b_tempering", C.c_int), ("b_targeting", C.c_int),
                ("global_tempering", C.c_double), ("bias_factor", C.c_double), ("boltzmann_factor", C.c_double),
                ("hill_prefactor", C.c_double), ("bias_per_step", C.c_double), ("hill_density", C.c_double),
                ("expected_target", C.c_double), ("total_volume", C.c_double)]


class BiasState(C.Structure):
    _fields_ = [("cum_bias", C.c_double), ("temp_hill_cum", C.c_double), ("steps", C.c_longlong),
                ("hills_added", C.c_int), ("skipped", C.c_int), ("backlog_left", C.c_long),
                ("backlog_right", C.c_long), ("n_accepted", C.c_long), ("log_dropped", C.c_long)]


class PairDomain(C.Structure):
    _fields_ = [("lo", C.c_double * 3), ("hi", C.c_double * 3), ("periodic", C.c_int * 3), ("nlocal", C.c_long)]


class PairResult(C.Structure):
    _fields_ = [("energy", C.c_double), ("n_pairs", C.c_longlong), ("n_calls", C.c_longlong)]


HILL_EVENT = np.dtype(
    [("steps", "<i8"), ("type", "<i4"), ("hills_added", "<i4"), ("pos", "<f8", (3,)), ("height", "<f8"),
     ("bias_added", "<f8"), ("cum_over_vol", "<f8")],
    align=True,
)

EXPORTS = {
    # name: (restype, argtypes)
    "edm_last_error": (C.c_char_p, []),
    "edm_device_count": (C.c_int, [c_ip]),
    "edm_uniform": (C.c_double, [C.c_uint64, C.c_uint64, C.c_uint64]),
    "edm_uniform_pair": (C.c_double, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]),
    "edm_grid_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_int, C.c_int]),
    "edm_grid_create_from_header": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, c_ip, c_dp, c_dp, c_ip, C.c_int, C.c_int]),
    "edm_gauss_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_int, c_dp]),
    "edm_grid_destroy": (C.c_int, [vp]),
    "edm_grid_set_boundary": (C.c_int, [vp, c_dp, c_dp, c_ip]),
    "edm_grid_geometry": (C.c_int, [vp, c_ip, c_ip, c_dp, c_dp, c_dp, c_ip, c_ip, C.POINTER(C.c_size_t)]),
    "edm_grid_flags": (C.c_int, [vp, c_ip, c_ip, c_ip]),
    "edm_grid_boundary": (C.c_int, [vp, c_dp, c_dp, c_ip, c_dp]),
    "edm_grid_set_interpolation": (C.c_int, [vp, C.c_int]),
    "edm_grid_upload": (C.c_int, [vp, c_dp, c_dp]),
    "edm_grid_download": (C.c_int, [vp, c_dp, c_dp]),
    "edm_grid_clear": (C.c_int, [vp]),
    "edm_grid_eval": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp, c_dp]),
    "edm_grid_eval_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, vp, vp, vp]),
    "edm_grid_eval_plain": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp, c_dp]),
    "edm_grid_get_value": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp]),
    "edm_grid_hist_add": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp]),
    "edm_grid_add": (C.c_int, [vp, vp, C.c_double, C.c_double]),
    "edm_grid_minmax": (C.c_int, [vp, c_dp, c_dp]),
    "edm_grid_remap": (C.c_int, [vp, c_dp]),
    "edm_gauss_deposit": (C.c_int, [vp, C.c_long, c_dp, c_dp, c_dp]),
    "edm_gauss_deposit_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, vp]),
    "edm_bias_create": (C.c_int, [C.POINTER(vp), vp, vp, vp, C.POINTER(BiasParams)]),
    "edm_bias_destroy": (C.c_int, [vp]),
    "edm_bias_state": (C.c_int, [vp, C.POINTER(BiasState)]),
    "edm_bias_set_cum_bias": (C.c_int, [vp, C.c_double]),
    "edm_bias_backlog_get": (C.c_int, [vp, c_lp, c_lp, c_dp]),
    "edm_bias_backlog_set": (C.c_int, [vp, C.c_long, C.c_long, c_dp]),
    "edm_bias_log_read": (C.c_int, [vp, vp, C.c_long, c_lp]),
    "edm_bias_update_forces": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp, C.c_long, c_ip, C.c_int, c_dp]),
    "edm_bias_step_coords": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp, C.c_long, c_ip, C.c_int, C.c_int, c_dp,
                                       C.c_uint64, C.c_uint64, c_dp]),
    "edm_bias_step_coords_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, vp, C.c_long, vp, C.c_int, C.c_int, vp,
                                           C.c_uint64, C.c_uint64, vp, vp]),
    "edm_bias_round_after": (C.c_int, [vp, vp]),
    "edm_bias_energy_dev": (C.c_int, [vp, vp, vp]),
    "edm_bias_energy_with_round": (C.c_int, [vp, vp]),
    "edm_bias_round_commit_on": (C.c_int, [vp, vp]),
    "edm_bias_update_forces_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, vp, C.c_long, vp, C.c_int, vp, vp]),
    "edm_bias_add_hills": (C.c_int, [vp, C.c_long, c_dp, C.c_long, c_dp, c_ip, C.c_int, C.c_uint64, C.c_uint64]),
    "edm_bias_add_hills_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, vp, vp, C.c_int, C.c_uint64, C.c_uint64, vp]),
    "edm_bias_pre_add_hill": (C.c_int, [vp, C.c_int]),
    "edm_bias_add_hill_batch": (C.c_int, [vp, C.c_long, c_dp, c_dp]),
    "edm_bias_post_add_hill": (C.c_int, [vp]),
    "edm_pair_step_cells": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_int, C.c_int, c_dp, C.c_double, C.c_int,
                                      C.c_longlong, C.c_uint64, C.c_uint64, C.POINTER(PairResult)]),
    "edm_pair_step_cells_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_int, C.c_int, c_dp, C.c_double, C.c_int,
                                          C.c_longlong, C.c_uint64, C.c_uint64, C.POINTER(PairResult), vp]),
    "edm_pair_step_cells_domain": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_int, C.c_int, C.POINTER(PairDomain), C.c_double,
                                             C.c_int, C.c_longlong, C.c_uint64, C.c_uint64, C.POINTER(PairResult)]),
    "edm_pair_step_cells_domain_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_int, C.c_int, C.POINTER(PairDomain),
                                                 C.c_double, C.c_int, C.c_longlong, C.c_uint64, C.c_uint64,
                                                 C.POINTER(PairResult), vp]),
    "edm_pair_select_cells_domain_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_int, C.c_int, C.POINTER(PairDomain),
                                                   C.c_double, C.c_longlong, C.c_uint64, C.c_uint64, vp, vp]),
    "edm_pair_step_list": (C.c_int, [vp, C.c_long, C.c_long, c_dp, c_dp, c_ip, C.c_int, C.c_int, C.c_long, c_ip,
                                     c_lp, c_ip, C.c_int, C.c_longlong, c_dp, C.c_uint64, C.c_uint64,
                                     C.POINTER(PairResult)]),
    "edm_hill_block_doubles": (C.c_size_t, [C.c_int, C.c_long]),
    "edm_bias_select_dev": (C.c_int, [vp, C.c_long, vp, C.c_long, vp, vp, C.c_int, C.c_longlong, C.c_uint64,
                                      C.c_uint64, C.c_uint64, vp]),
    "edm_pair_select_cells_dev": (C.c_int, [vp, C.c_long, vp, vp, vp, C.c_int, C.c_int, c_dp, C.c_double,
                                            C.c_longlong, C.c_uint64, C.c_uint64, vp, vp]),
    "edm_bias_hills_pack_dev": (C.c_int, [vp, vp, C.c_long, vp]),
    "edm_launch_count": (C.c_int, [C.POINTER(C.c_longlong)]),
    "edm_bias_set_profiling": (C.c_int, [vp, C.c_int]),
    "edm_bias_profile_ms": (C.c_int, [vp, c_dp]),
    "edm_bias_profile_pair_ms": (C.c_int, [vp, c_dp, c_dp]),
    "edm_bias_profile_e2e_ms": (C.c_int, [vp, c_dp, c_dp, c_dp, c_dp]),
    "edm_bias_round_times_us": (C.c_int, [vp, c_dp]),
    "edm_bias_exchange_times_us": (C.c_int, [vp, c_dp]),
    "edm_host_pin": (C.c_int, [vp, C.c_size_t]),
    "edm_host_unpin": (C.c_int, [vp]),
    "edm_pair_list_set": (C.c_int, [vp, C.c_long, c_ip, C.POINTER(C.c_long), c_ip]),
    "edm_pair_step_listed": (C.c_int, [vp, C.c_long, C.c_long, c_dp, c_dp, c_ip, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                       c_dp, C.c_uint64, C.c_uint64, C.POINTER(PairResult)]),
    "edm_pair_search_info": (C.c_int, [vp, c_ip, c_dp, C.POINTER(C.c_longlong)]),
    "edm_bias_round_info": (C.c_int, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "edm_bias_hills_commit_dev": (C.c_int, [vp, vp, C.c_int, C.c_long, C.c_longlong, vp]),
    "edm_comm_nccl_version": (C.c_int, [c_ip]),
    "edm_comm_unique_id": (C.c_int, [C.c_char_p]),
    "edm_comm_init_rank": (C.c_int, [C.POINTER(vp), C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "edm_comm_init_file": (C.c_int, [C.POINTER(vp), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double]),
    "edm_comm_init_all": (C.c_int, [C.POINTER(vp), C.c_int, c_ip]),
    "edm_comm_from_nccl": (C.c_int, [C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int]),
    "edm_comm_destroy": (C.c_int, [vp]),
    "edm_comm_info": (C.c_int, [vp, c_ip, c_ip, c_ip]),
    "edm_comm_peer_windows": (C.c_int, [vp, c_ip]),
    "edm_comm_group_start": (C.c_int, []),
    "edm_comm_group_end": (C.c_int, []),
    "edm_comm_allreduce_sum_dev": (C.c_int, [vp, vp, C.c_long, vp]),
    "edm_bias_exchange_dev": (C.c_int, [vp, vp, C.c_long, C.c_longlong, vp]),
    "edm_bias_exchange_all_dev": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_long, C.c_longlong, C.POINTER(vp)]),
    "edm_bias_set_comm": (C.c_int, [vp, vp, C.c_long]),
    "edm_bias_check": (C.c_int, [vp]),
}

_lib = None


def lib():
    """Loads the CUDA library; raises EdmError when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EdmError("libedm_b200.so is missing: run `python electronic-dance-music_b200/build.py` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise EdmError("edm_b200 error %d: %s" % (rc, lib().edm_last_error().decode()))


def device_count():
    n = C.c_int(0)
    lib().edm_device_count(C.byref(n))
    return n.value


def launch_count():
    n = C.c_longlong(0)
    lib().edm_launch_count(C.byref(n))
    return n.value


def uniform(seed, step, counter):
    return lib().edm_uniform(seed, step, counter)


def uniform_pair(seed, step, pairkey, which):
    return lib().edm_uniform_pair(seed, step, pairkey, which)


def _d(a):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64)))


def _i(a):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.int32)))


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


class Comm:
    """One rank's end of the hill exchange (an NCCL communicator owned by the library)."""

    def __init__(self, handle):
        self.L = lib()
        self.h = handle

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(lib().edm_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def init_rank(cls, uid, nranks, rank, device):
        h = vp()
        check(lib().edm_comm_init_rank(C.byref(h), uid, nranks, rank, device))
        return cls(h)

    @classmethod
    def init_file(cls, path, nranks, rank, device, timeout_s=120.0):
        h = vp()
        check(lib().edm_comm_init_file(C.byref(h), path.encode(), nranks, rank, device, timeout_s))
        return cls(h)

    @classmethod
    def init_all(cls, ndev, devices=None):
        hs = (vp * ndev)()
        d = _i(devices) if devices is not None else None
        check(lib().edm_comm_init_all(hs, ndev, _ip(d) if d is not None else None))
        return [cls(vp(h)) for h in hs]

    def info(self):
        n, r, d = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.L.edm_comm_info(self.h, C.byref(n), C.byref(r), C.byref(d)))
        return dict(nranks=n.value, rank=r.value, device=d.value)

    def peer_windows(self):
        """True when the exchange runs over NVLink peer windows (one kernel), False when it is an ncclAllGather."""
        e = C.c_int(0)
        check(self.L.edm_comm_peer_windows(self.h, C.byref(e)))
        return bool(e.value)

    def destroy(self):
        if self.h:
            self.L.edm_comm_destroy(self.h)
            self.h = None


def nccl_version():
    v = C.c_int(0)
    check(lib().edm_comm_nccl_version(C.byref(v)))
    return v.value


class Grid:
    """Device-resident Grid / GaussGrid (lib/grid.h:185, lib/gaussian_grid.h:59)."""

    def __init__(self, dim=None, mn=None, mx=None, spacing=None, periodic=None, b_deriv=0, b_interp=0, sigma=None,
                 device=0, header_bins=None, handle=None):
        self.L = lib()
        self.owned = handle is None
        if handle is not None:
            self.h = vp(handle) if not isinstance(handle, vp) else handle
            d = C.c_int(0)
            check(self.L.edm_grid_geometry(self.h, C.byref(d), None, None, None, None, None, None, None))
            self.dim = d.value
            return
        self.dim = dim
        h = vp()
        if sigma is not None:
            check(self.L.edm_gauss_create(C.byref(h), device, dim, _dp(_d(mn)), _dp(_d(mx)), _dp(_d(spacing)),
                                          _ip(_i(periodic)), int(b_interp), _dp(_d(sigma))))
        elif header_bins is not None:
            check(self.L.edm_grid_create_from_header(C.byref(h), device, dim, _ip(_i(header_bins)), _dp(_d(mn)),
                                                     _dp(_d(mx)), _ip(_i(periodic)), int(b_deriv), int(b_interp)))
        else:
            check(self.L.edm_grid_create(C.byref(h), device, dim, _dp(_d(mn)), _dp(_d(mx)), _dp(_d(spacing)),
                                         _ip(_i(periodic)), int(b_deriv), int(b_interp)))
        self.h = h

    def __del__(self):
        if getattr(self, "owned", False) and getattr(self, "h", None):
            self.L.edm_grid_destroy(self.h)
            self.h = None

    def set_boundary(self, mn, mx, periodic):
        check(self.L.edm_grid_set_boundary(self.h, _dp(_d(mn)), _dp(_d(mx)), _ip(_i(periodic))))

    def info(self):
        n = np.zeros(3, np.int32)
        per = np.zeros(3, np.int32)
        mini = np.zeros(3, np.int32)
        dx, mn, mx = np.zeros(3), np.zeros(3), np.zeros(3)
        sz = C.c_size_t(0)
        check(self.L.edm_grid_geometry(self.h, None, _ip(n), _dp(dx), _dp(mn), _dp(mx), _ip(per), _ip(mini), C.byref(sz)))
        D = self.dim
        return dict(n=n[:D].copy(), dx=dx[:D].copy(), min=mn[:D].copy(), max=mx[:D].copy(), periodic=per[:D].copy(),
                    minisize=mini[:D].copy(), size=sz.value)

    @property
    def size(self):
        return self.info()["size"]

    def set_interpolation(self, b):
        check(self.L.edm_grid_set_interpolation(self.h, int(b)))

    def set_arrays(self, v, d=None):
        v = _d(v)
        dd = _d(d).ravel() if d is not None else None
        check(self.L.edm_grid_upload(self.h, _dp(v), _dp(dd) if dd is not None else None))

    def get_arrays(self):
        sz = self.size
        v = np.zeros(sz)
        d = np.zeros(sz * self.dim)
        check(self.L.edm_grid_download(self.h, _dp(v), _dp(d)))
        return v, d.reshape(sz, self.dim)

    def clear(self):
        check(self.L.edm_grid_clear(self.h))

    def eval(self, x, xstride=None):
        x = _d(x)
        xs = xstride or self.dim
        x = x.reshape(-1, xs)
        n = x.shape[0]
        val = np.zeros(n)
        der = np.zeros((n, self.dim))
        check(self.L.edm_grid_eval(self.h, n, _dp(x), xs, _dp(val), _dp(der)))
        return val, der

    def get_value(self, x):
        x = _d(x).reshape(-1, self.dim)
        val = np.zeros(x.shape[0])
        check(self.L.edm_grid_get_value(self.h, x.shape[0], _dp(x), self.dim, _dp(val)))
        return val

    def hist_add(self, x, v):
        x = _d(x).reshape(-1, self.dim)
        v = _d(v)
        check(self.L.edm_grid_hist_add(self.h, x.shape[0], _dp(x), self.dim, _dp(v)))

    def add(self, other, scale, offset):
        check(self.L.edm_grid_add(self.h, other.h, float(scale), float(offset)))

    def minmax(self):
        a, b = C.c_double(0), C.c_double(0)
        check(self.L.edm_grid_minmax(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def remap(self, x):
        x = _d(x).copy()
        check(self.L.edm_grid_remap(self.h, _dp(x)))
        return x

    def add_value(self, x, h):
        return float(self.add_values(_d(x).reshape(1, self.dim), [h])[0])

    def add_values(self, x, h):
        """GaussGrid::add_value for a list of hills (host buffers); returns each hill's bias_added."""
        x = _d(x).reshape(-1, self.dim)
        h = _d(h)
        ba = np.zeros(x.shape[0])
        check(self.L.edm_gauss_deposit(self.h, x.shape[0], _dp(x), _dp(h), _dp(ba)))
        return ba


def GaussGrid(dim, mn, mx, spacing, periodic, interp, sigma, device=0):
    return Grid(dim, mn, mx, spacing, periodic, b_deriv=1, b_interp=interp, sigma=sigma, device=device)


class Bias:
    """Device-resident EDMBias step state (lib/edm_bias.h:29) over existing device grids."""

    def __init__(self, bias_grid, hist_grid, params, target=None):
        self.L = lib()
        self.bias_grid, self.hist_grid, self.target = bias_grid, hist_grid, target
        self.dim = bias_grid.dim
        self.params = params
        p = BiasParams(**params)
        h = vp()
        check(self.L.edm_bias_create(C.byref(h), bias_grid.h, hist_grid.h if hist_grid else None,
                                     target.h if target else None, C.byref(p)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.edm_bias_destroy(self.h)
            self.h = None

    def state(self):
        s = BiasState()
        check(self.L.edm_bias_state(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in BiasState._fields_}

    def set_cum_bias(self, v):
        check(self.L.edm_bias_set_cum_bias(self.h, float(v)))

    def backlog(self):
        l, r = C.c_long(0), C.c_long(0)
        buf = np.zeros(EDM_BUFFER_DBLS)
        check(self.L.edm_bias_backlog_get(self.h, C.byref(l), C.byref(r), _dp(buf)))
        return l.value, r.value, buf

    def set_backlog(self, left, right, buf):
        buf = _d(buf)
        assert buf.size == EDM_BUFFER_DBLS
        check(self.L.edm_bias_backlog_set(self.h, left, right, _dp(buf)))

    def log(self):
        n = C.c_long(0)
        check(self.L.edm_bias_log_read(self.h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=HILL_EVENT)
        check(self.L.edm_bias_log_read(self.h, out.ctypes.data_as(vp), out.size, C.byref(n)))
        return out[:n.value]

    def update_forces(self, x, f, mask=None, apply_mask=-1):
        assert x.flags.c_contiguous and f.flags.c_contiguous and f.dtype == np.float64 and x.dtype == np.float64
        e = C.c_double(0)
        m = _i(mask) if mask is not None else None
        check(self.L.edm_bias_update_forces(self.h, x.shape[0], _dp(x), x.shape[1], _dp(f), f.shape[1],
                                            _ip(m) if m is not None else None, int(apply_mask), C.byref(e)))
        return e.value

    def step_coords(self, x, f, runiform=None, mask=None, apply_mask=-1, do_hills=True, seed=0, step=0):
        """update_forces + add_hills over the same atoms in one pipelined call (fix edm's post_force)."""
        assert x.flags.c_contiguous and f.flags.c_contiguous and f.dtype == np.float64 and x.dtype == np.float64
        e = C.c_double(0)
        u = _d(runiform) if runiform is not None else None
        m = _i(mask) if mask is not None else None
        check(self.L.edm_bias_step_coords(self.h, x.shape[0], _dp(x), x.shape[1], _dp(f), f.shape[1],
                                          _ip(m) if m is not None else None, int(apply_mask), 1 if do_hills else 0,
                                          _dp(u) if u is not None else None, seed, step, C.byref(e)))
        return e.value

    def add_hills(self, x, runiform=None, mask=None, apply_mask=-1, seed=0, step=0):
        assert x.flags.c_contiguous and x.dtype == np.float64
        u = _d(runiform) if runiform is not None else None
        m = _i(mask) if mask is not None else None
        check(self.L.edm_bias_add_hills(self.h, x.shape[0], _dp(x), x.shape[1], _dp(u) if u is not None else None,
                                        _ip(m) if m is not None else None, int(apply_mask), seed, step))

    def pre_add_hill(self, est):
        check(self.L.edm_bias_pre_add_hill(self.h, int(est)))

    def add_hill_many(self, x, u):
        x = _d(x).reshape(-1, self.dim)
        u = _d(u)
        check(self.L.edm_bias_add_hill_batch(self.h, x.shape[0], _dp(x), _dp(u)))

    def post_add_hill(self):
        check(self.L.edm_bias_post_add_hill(self.h))

    def pair_step_cells(self, x, f, box, cutoff, do_hills=False, est=0, seed=0, step=0, types=None, itype=0, jtype=0):
        assert x.flags.c_contiguous and f.flags.c_contiguous and x.shape[1] == 3 and f.shape[1] == 3
        r = PairResult()
        t = _i(types) if types is not None else None
        check(self.L.edm_pair_step_cells(self.h, x.shape[0], _dp(x), _dp(f), _ip(t) if t is not None else None, itype,
                                         jtype, _dp(_d(box)), float(cutoff), int(do_hills), int(est), seed, step,
                                         C.byref(r)))
        return dict(energy=r.energy, n_pairs=r.n_pairs, n_calls=r.n_calls)

    def pair_step_cells_domain(self, x, f, lo, hi, periodic, nlocal, cutoff, do_hills=False, est=0, seed=0, step=0,
                               types=None, itype=0, jtype=0):
        """One rank's share: x has nall rows (local atoms first, then ghosts), f has nlocal rows."""
        assert x.flags.c_contiguous and f.flags.c_contiguous and x.shape[1] == 3 and f.shape == (nlocal, 3)
        dom = PairDomain((C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_int * 3)(*[int(p) for p in periodic]), nlocal)
        r = PairResult()
        t = _i(types) if types is not None else None
        check(self.L.edm_pair_step_cells_domain(self.h, x.shape[0], _dp(x), _dp(f), _ip(t) if t is not None else None,
                                                itype, jtype, C.byref(dom), float(cutoff), int(do_hills), int(est), seed,
                                                step, C.byref(r)))
        return dict(energy=r.energy, n_pairs=r.n_pairs, n_calls=r.n_calls)

    def pair_search_info(self):
        bd = (C.c_int * 3)()
        sc = C.c_double(0)
        fb = C.c_longlong(0)
        check(self.L.edm_pair_search_info(self.h, bd, C.byref(sc), C.byref(fb)))
        return dict(bricks=tuple(bd), density_scale=sc.value, fallbacks=fb.value)

    def pair_list_set(self, ilist, first, jlist):
        """Uploads a half neighbour list (CSR) once; pair_step_listed then reuses it until the next call."""
        il, jl = _i(ilist), _i(jlist)
        fi = np.ascontiguousarray(np.asarray(first, dtype=np.int64))
        check(self.L.edm_pair_list_set(self.h, il.size, _ip(il), fi.ctypes.data_as(C.POINTER(C.c_long)), _ip(jl)))

    def pair_step_listed(self, x, f, nlocal, do_hills=False, est=0, runiform=None, seed=0, step=0, types=None, itype=0,
                         jtype=0):
        assert x.flags.c_contiguous and f.flags.c_contiguous and x.shape[1] == 3 and f.shape[1] == 3
        r = PairResult()
        t = _i(types) if types is not None else None
        u = _d(runiform) if runiform is not None else None
        check(self.L.edm_pair_step_listed(self.h, x.shape[0], int(nlocal), _dp(x), _dp(f),
                                          _ip(t) if t is not None else None, itype, jtype, 1 if do_hills else 0, int(est),
                                          _dp(u) if u is not None else None, seed, step, C.byref(r)))
        return dict(energy=r.energy, n_pairs=r.n_pairs, n_calls=r.n_calls)

    def set_comm(self, comm, cap=0):
        check(self.L.edm_bias_set_comm(self.h, comm.h if comm is not None else None, cap))

    def check(self):
        check(self.L.edm_bias_check(self.h))

    def round_times_us(self):
        out = np.zeros(16)
        check(self.L.edm_bias_round_times_us(self.h, _dp(out)))
        return out

    def exchange_times_us(self):
        out = np.zeros(5)
        check(self.L.edm_bias_exchange_times_us(self.h, _dp(out)))
        return out

    def round_info(self):
        a, m, b = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        check(self.L.edm_bias_round_info(self.h, C.byref(a), C.byref(m), C.byref(b)))
        return dict(parallel=a.value, split=m.value, in_order=b.value)

    def pair_step_list(self, x, f, nlocal, ilist, first, jlist, do_hills=False, est=0, runiform=None, seed=0, step=0,
                       types=None, itype=0, jtype=0):
        assert x.flags.c_contiguous and f.flags.c_contiguous and x.shape[1] == 3 and f.shape[1] == 3
        r = PairResult()
        il, jl = _i(ilist), _i(jlist)
        fi = np.ascontiguousarray(np.asarray(first, dtype=np.int64))
        t = _i(types) if types is not None else None
        u = _d(runiform) if runiform is not None else None
        check(self.L.edm_pair_step_list(self.h, x.shape[0], int(nlocal), _dp(x), _dp(f),
                                        _ip(t) if t is not None else None, itype, jtype, il.size, _ip(il),
                                        fi.ctypes.data_as(c_lp), _ip(jl), int(do_hills), int(est),
                                        _dp(u) if u is not None else None, seed, step, C.byref(r)))
        return dict(energy=r.energy, n_pairs=r.n_pairs, n_calls=r.n_calls)


def parse_edm_text(text):
    kv = {}
    for line in text.splitlines():
        p = line.split()
        if p and p[0] not in kv:
            kv[p[0]] = p[1:]
    return kv


def bias_from_edm(edm_file, temperature, boltz, sublo, subhi, boxlo, boxhi, periodic, skin, device=0, target=None,
                  expected_target=0.0):
    """Python rendering of `new EDMBias(file); setup(T, kB); subdivide(...)` (lib/edm_bias.cpp:34-69,
    264-269, 98-222) for test/bench drivers; the C++ EDM::EDMBias class does the same natively."""
    kv = parse_edm_text(open(edm_file).read())
    D = int(kv["dimension"][0])
    temp = int(kv["tempering"][0])
    gt = float(kv.get("global_tempering", [0])[0]) if temp else 0.0
    bf = float(kv.get("bias_factor", [0])[0]) if temp else 0.0
    pref = float(kv["hill_prefactor"][0])
    bps = float(kv["bias_per_step"][0]) if "bias_per_step" in kv else pref
    dens = float(kv["hill_density"][0]) if "hill_density" in kv else -1.0
    arr = lambda k: np.array([float(v) for v in kv[k][:D]])
    dx, sg, lo, hi = arr("bias_spacing"), arr("bias_sigma"), arr("box_low"), arr("box_high")
    sublo, subhi, boxlo, boxhi, skin = (np.atleast_1d(np.asarray(a, float)) for a in (sublo, subhi, boxlo, boxhi, skin))
    periodic = np.atleast_1d(np.asarray(periodic, int))
    bper = np.zeros(D, np.int32)
    gper = np.zeros(D, np.int32)
    mn, mx = np.zeros(D), np.zeros(D)
    for i in range(D):
        if abs(boxlo[i] - lo[i]) < 0.000001 and abs(boxhi[i] - hi[i]) < 0.000001:
            bper[i] = periodic[i]
    for i in range(D):
        mn[i], mx[i] = sublo[i], subhi[i]
        if abs(sublo[i] - lo[i]) < 0.000001 and abs(subhi[i] - hi[i]) < 0.000001:
            gper[i] = periodic[i]
        else:
            mn[i] -= skin[i]
            mx[i] += skin[i]
    g = GaussGrid(D, mn, mx, dx, gper, 1, sg, device=device)
    hist = Grid(D, mn, mx, sg, gper, 0, 0, device=device)
    g.set_boundary(lo, hi, bper)
    vol = 1.0
    for i in range(D):
        vol *= hi[i] - lo[i]
    tv = 0.0
    tv += vol
    params = dict(dim=D, b_tempering=temp, b_targeting=1 if target is not None else 0, global_tempering=gt,
                  bias_factor=bf, boltzmann_factor=boltz * temperature, hill_prefactor=pref, bias_per_step=bps,
                  hill_density=dens, expected_target=float(expected_target), total_volume=tv)
    return Bias(g, hist, params, target=target)
