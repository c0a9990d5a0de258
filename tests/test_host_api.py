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


@pytest.mark.gpu
def test_restated_reference_unit_tests_on_gpu(host_test_binary, tmp_path):
    """tests/edm_test.cpp restated (tests_host/edm_host_test.cpp), run against the device."""
    src = os.path.join(REF, "tests") if os.path.isdir(os.path.join(REF, "tests")) else "-"
    r = subprocess.run([host_test_binary, src], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:]
    assert " 0 failed" in r.stdout
