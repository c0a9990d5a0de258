"""The reference's Python API, `edm.EDMBias` (python/edm/__init__.py:1-8 over the boost::python class
EDMBias_Py, python/edm/edm_python.cxx:6-19), on the B200 engine.

    from edm_b200.compat import EDMBias
    bias = EDMBias("input.edm", 1.0, 1.0)
    bias.set_box([0], [10], [0])
    bias.add_hill([0.25])
    energy, dvdx = bias.get_force([0.24])      # python-example/EDM.ipynb:103

Same method names and argument meaning; ctypes over libedm.so (EDM::EDMBias, C++) instead of
Boost.Python.  Every call lands on the GPU through the C ABI — there is no CPU path.
"""
import ctypes as C
import os
import random

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libedm.so"))
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError("libedm.so is missing: run `python electronic-dance-music_b200/build.py --host`")
        L = C.CDLL(_LIB)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.edm_py_new.restype = vp
        L.edm_py_new.argtypes = [C.c_char_p, C.c_double, C.c_double]
        L.edm_py_delete.argtypes = [vp]
        L.edm_py_dim.argtypes = [vp]
        L.edm_py_set_box.argtypes = [vp, C.c_int, dp, dp, ip]
        L.edm_py_pre_add_hill.argtypes = [vp, C.c_int]
        L.edm_py_post_add_hill.argtypes = [vp]
        L.edm_py_add_hill.argtypes = [vp, dp, C.c_double]
        L.edm_py_write_bias.argtypes = [vp, C.c_char_p]
        L.edm_py_write_lammps_table.argtypes = [vp, C.c_char_p]
        L.edm_py_write_histogram.argtypes = [vp]
        L.edm_py_clear_histogram.argtypes = [vp]
        L.edm_py_get_force.restype = C.c_double
        L.edm_py_get_force.argtypes = [vp, dp, dp]
        _lib = L
    return _lib


def _doubles(seq, n):
    if len(seq) > n:   # convert_list, python/edm/edm_bias_py.cpp:8-15
        raise ValueError("Tried to convert a list that was too big")
    return (C.c_double * n)(*[float(v) for v in seq])


class EDMBias_Py(object):
    def __init__(self, input_filename, temperature, boltzmann_constant):
        self._L = _load()
        self._h = self._L.edm_py_new(os.fsencode(input_filename), float(temperature), float(boltzmann_constant))
        self._dim = self._L.edm_py_dim(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.edm_py_delete(self._h)
            self._h = None

    def set_box(self, boxlo, boxhi, periodic):
        n = len(boxlo)
        per = (C.c_int * 3)(*([int(p) for p in periodic] + [0] * (3 - len(periodic))))
        self._L.edm_py_set_box(self._h, n, _doubles(boxlo, 3), _doubles(boxhi, 3), per)

    def pre_add_hill(self, est_hill_count):
        self._L.edm_py_pre_add_hill(self._h, int(est_hill_count))

    def post_add_hill(self):
        self._L.edm_py_post_add_hill(self._h)

    def add_hill_r(self, position, runiform):
        self._L.edm_py_add_hill(self._h, _doubles(position, self._dim), float(runiform))

    def write_bias(self, output):
        self._L.edm_py_write_bias(self._h, os.fsencode(output))

    def write_lammps_table(self, output):
        self._L.edm_py_write_lammps_table(self._h, os.fsencode(output))

    def write_histogram(self):
        self._L.edm_py_write_histogram(self._h)

    def clear_histogram(self):
        self._L.edm_py_clear_histogram(self._h)

    def get_force(self, position):
        f = (C.c_double * self._dim)()
        e = self._L.edm_py_get_force(self._h, _doubles(position, self._dim), f)
        return e, list(f)


class EDMBias(EDMBias_Py):   # python/edm/__init__.py:4-8
    def add_hill(self, position):
        self.pre_add_hill(1)
        self.add_hill_r(position, random.random())
        self.post_add_hill()
