"""The C++ host API (EDM::Grid / GaussGrid / EDMBias over the C ABI) and the LAMMPS entry points.

CPU part: everything compiles — libedm.so, the restated reference test binary, the two B200 fixes
against the mock LAMMPS headers, and (where /root/reference is present) the REFERENCE's own
lammps/fix_edm.cpp and fix_edm_pair.cpp against this repo's <edm/edm_bias.h>, which is the drop-in
claim in its strongest form: the reference's callers build unchanged.
GPU part: the restated tests/edm_test.cpp cases run green on the device.
"""
import importlib.util
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "electronic-dance-music_b200")
REF = "/root/reference"


def build_module():
    spec = importlib.util.spec_from_file_location("edm_b200_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def host_test_binary():
    return build_module().build_host_tests()


def test_host_library_and_fixes_compile(host_test_binary):
    assert os.path.exists(host_test_binary)
    for f in ("libedm.so", "fix_edm.o", "fix_edm_pair.o"):
        assert os.path.exists(os.path.join(PKG, "lib", f)), f


def test_host_api_keeps_the_reference_surface():
    """Names a consumer of <edm/edm_bias.h> relies on."""
    bias_h = open(os.path.join(PKG, "edm", "edm_bias.h")).read()
    for name in ("class EDMBias", "void subdivide(", "void setup(", "double update_forces(", "double update_force(",
                 "void set_mask(", "void add_hills(", "void pre_add_hill(", "void add_hill(", "void post_add_hill(",
                 "void write_bias(", "void write_histogram(", "void clear_histogram(", "void write_lammps_table(",
                 "dim_", "bias_", "b_tempering_", "hill_prefactor_", "bias_sigma_", "bias_dx_", "cum_bias_",
                 "BIAS_BUFFER_SIZE 2048"):
        assert name in bias_h, name
    grid_h = open(os.path.join(PKG, "edm", "grid.h")).read()
    for name in ("class Grid", "class DimmedGrid", "make_grid(", "read_grid(", "get_value_deriv(", "multi2one(",
                 "one2multi(", "grid_number_", "b_interpolate_", "grid_deriv_"):
        assert name in grid_h, name
    gauss_h = open(os.path.join(PKG, "edm", "gaussian_grid.h")).read()
    for name in ("class GaussGrid", "class DimmedGaussGrid", "make_gauss_grid(", "read_gauss_grid(", "set_boundary(",
                 "get_volume(", "in_bounds(", "remap(", "lammps_multi_write(", "minisize_", "boundary_min_"):
        assert name in gauss_h, name


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lammps")), reason="reference tree absent")
@pytest.mark.parametrize("fix", ["fix_edm.cpp", "fix_edm_pair.cpp"])
def test_reference_fixes_compile_unchanged_against_this_api(fix, tmp_path):
    """The reference's own LAMMPS fixes, compiled where they lie, against this repo's EDM headers."""
    mock = os.path.join(PKG, "lammps", "mock")
    cmd = ["/usr/bin/g++", "-std=c++11", "-fsyntax-only", "-w", "-I" + PKG, "-I" + mock,
           "-I" + os.path.join(REF, "lammps"), os.path.join(REF, "lammps", fix)]
    subprocess.check_call(cmd)


def write_plumed_grid(path, dim, n, dx, mn, mx, per, force, values, derivs):
    """PLUMED-1 grid text as the reference writes it (lib/grid.h:448-503), from stored arrays."""
    import numpy as np
    with open(path, "w") as fh:
        fh.write("#! FORCE %d\n#! NVAR %d\n#! TYPE %s\n" % (force, dim, " ".join(["32"] * dim) + " "))
        fh.write("#! BIN %s \n" % " ".join(str(int(n[i] if per[i] else n[i] - 1)) for i in range(dim)))
        fh.write("#! MIN %s \n" % " ".join("%g" % mn[i] for i in range(dim)))
        fh.write("#! MAX %s \n" % " ".join("%g" % (mx[i] if per[i] else mx[i] - dx[i]) for i in range(dim)))
        fh.write("#! PBC %s \n" % " ".join(str(int(per[i])) for i in range(dim)))
        idx = np.zeros(dim, dtype=np.int64)
        for p in range(values.size):
            t = p
            for j in range(dim - 1):
                idx[j] = t % n[j]
                t = (t - idx[j]) // n[j]
            idx[dim - 1] = t
            row = ["%.8f" % (mn[j] + dx[j] * idx[j]) for j in range(dim)] + ["%.8f" % values[p]]
            if force:
                row += ["%.8f" % (-derivs[p, j]) for j in range(dim)]
            fh.write(" ".join(row) + " \n")
            if idx[0] == n[0] - 1:
                fh.write("\n")


def make_fixture_dir(tmp_path):
    """1.grid 2.grid 3.grid read_test.edm rebuilt from tests/golden/plumed_grids.npz (the arrays the
    reference's reader produced from its own fixtures), so the file-based cases run without the tree."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "plumed_grids.npz"))
    d = tmp_path / "fixtures"
    d.mkdir()
    for dim in (1, 2, 3):
        s = str(dim)
        write_plumed_grid(str(d / (s + ".grid")), dim, z["n" + s], z["dx" + s], z["min" + s], z["max" + s],
                          z["periodic" + s], int(z["b_derivatives" + s]), z["grid" + s], z["deriv" + s])
    (d / "read_test.edm").write_text("dimension 2\ntempering 0\nhill_prefactor 1.0\nhill_density 1\n"
                                     "bias_spacing 1.0 1.0\nbias_sigma 2 1\ntarget_filename 2.grid.test\n"
                                     "box_low 0 0 \nbox_high 5 5\n")
    return str(d)


@pytest.mark.gpu
def test_restated_reference_unit_tests_on_gpu(host_test_binary, tmp_path):
    """tests/edm_test.cpp restated (tests_host/edm_host_test.cpp), run against the device."""
    src = make_fixture_dir(tmp_path)
    r = subprocess.run([host_test_binary, src], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:]
    assert " 0 failed" in r.stdout


@pytest.mark.gpu
def test_fixes_driven_by_a_fake_md_loop_on_gpu(host_test_binary, tmp_path):
    """fix edm / fix edm_pair constructed from a fix command, stepped like LAMMPS would (mock headers), every step's
    energy and forces checked against a CPU accumulation over the batched grid evaluation."""
    drv = os.path.join(PKG, "lib", "fix_driver_test")
    assert os.path.exists(drv)
    r = subprocess.run([drv], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:]
    assert " 0 failed" in r.stdout


@pytest.mark.gpu
def test_python_api_reproduces_the_notebook_vector(host_test_binary, tmp_path, monkeypatch):
    """edm.EDMBias (python/edm/__init__.py) on the B200 engine: python-example/EDM.ipynb:86-103."""
    import sys
    sys.path.insert(0, os.path.join(PKG, "python"))
    from edm_b200.compat import EDMBias
    monkeypatch.chdir(tmp_path)   # HILLS / HIST files open relative to the working directory
    f = tmp_path / "input.edm"
    f.write_text("tempering 0\nhill_prefactor 1.0\ndimension 1\nbox_low 0.0\nbox_high 1.0\nbias_spacing 0.01\n"
                 "bias_sigma 0.5\n")
    bias = EDMBias(str(f), 1, 1)
    bias.set_box([0], [10], [0])
    bias.pre_add_hill(1)
    bias.add_hill_r([0.25], 0.0)
    bias.post_add_hill()
    e, der = bias.get_force([0.24])
    assert abs(e - 1.1002417338159258) <= 1e-10 * 1.1002417338159258
    assert abs(der[0] - (-0.6144025830861709)) <= 1e-10 * 0.6144025830861709
    bias.add_hill([0.5])                       # the wrapper's own convenience method
    e2, _ = bias.get_force([0.24])
    assert e2 > e
    bias.write_bias(str(tmp_path / "BIAS"))
    bias.write_histogram()
    bias.clear_histogram()
    assert (tmp_path / "BIAS").exists()
    with pytest.raises(ValueError):
        bias.get_force([0.1, 0.2])
