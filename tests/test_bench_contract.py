"""The bench.py contract the driver relies on, checked on the CPU through the reference arm (the only arm
that runs without a GPU): one JSON line with the agreed keys, same metric/unit/config as the GPU arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def reference_line(port):
    import pyoracle
    if not (pyoracle.available("ref") or pyoracle.available("port")):
        pytest.skip("no oracle library built")
    env = dict(os.environ)
    # a bounded run: two worker processes instead of one per core keeps the CPU suite short
    code = ("import os, sys; os.sched_setaffinity(0, set(sorted(os.sched_getaffinity(0))[:2])); "
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1']; "
            "import runpy; runpy.run_path(%r, run_name='__main__')" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_one_line_with_the_agreed_keys(reference_line):
    d = reference_line
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "CV bias+force evals/sec" and d["unit"] == "evals/s" and d["dtype"] == "f64"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"] == "c2_pair_rdf"
    # the same config object as the GPU arm prints (bench.c2_config): the driver's same_config check
    for k in ("atoms_per_gpu", "number_density", "cutoff", "pairs_per_gpu_per_step", "grid_points", "hill_density",
              "step", "parallelism"):
        assert k in d["config"], k
    assert abs(d["config"]["pairs_per_gpu_per_step"] - 26179939) < 30000   # N (4/3) pi rc^3 rho / 2
    assert d["value"] > 1e6
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: the product arm of bench.py refuses to run where no CUDA device is visible."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_roofline_record_is_keyed_by_the_sources_of_its_kernel():
    """bench.py reports ncu traffic / binding resource only for a capture taken on the sources the kernel is compiled
    from (its translation unit + every header): a change elsewhere in csrc/ must not silence it, a change there must."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.kernel_unit("block_eval_kernel") == "edm_pair.cu"
    assert bench.kernel_unit("forces_kernel<2>") == "edm_bias.cu"
    assert bench.kernel_unit("no_such_kernel") is None
    h_pair, h_bias, h_tree = bench.csrc_hash("block_eval_kernel"), bench.csrc_hash("forces_kernel<3>"), bench.csrc_hash()
    assert len({h_pair, h_bias, h_tree}) == 3
    rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_roofline.json")))
    traffic, binding, note = bench.ncu_roofline("c2_pair_rdf", "block_eval_kernel")
    if rec["c2_pair_rdf"]["block_eval_kernel"]["csrc_hash"] == h_pair:
        assert traffic and binding["bound"] in ("lsu", "issue", "fp64", "hbm", "l2")
    else:   # the sources moved on since the capture (tools/capture_roofline.sh refreshes it): nothing may be reported
        assert traffic is None and binding is None and "not reported" in note
