"""The hill exchange (flush_buffers / check_for_flush -> all-gather, lib/edm_bias.cpp:614-706)
on ONE GPU: everything but the wire.

  * test_sharded_blocks_*: the candidates of every step are split into shards, each shard is selected and packed
    into its own block on the device (edm_bias_select_dev with the job-wide est_hill_count and the shard's first
    counter, edm_bias_hills_pack_dev), and the rank-major concatenation of the blocks is committed
    (edm_bias_hills_commit_dev, nblocks > 1) — exactly what every rank does after ncclAllGather.  The result is
    checked against the single-rank oracle over the concatenated candidates: decisions bit-exact, values 1e-10.
  * test_library_exchange_single_rank: edm_bias_exchange_dev / edm_bias_set_comm on a 1-rank communicator.
  * test_exchange_from_cpp: the C++ driver (tests_host/exchange_test.cpp): ncclCommInitAll over every visible
    device, one host thread per device, EDM::EDMBias::post_add_hill exchanging inside the library.
The N-process form (torchrun, NCCL between processes) is tests/test_multi_rank.py.
"""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from test_gpu_parity import BIAS_CASES, compare_bias, write_edm

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "electronic-dance-music_b200")


@pytest.fixture(scope="module")
def edm():
    import edm_b200
    if edm_b200.device_count() == 0:
        pytest.fail("no CUDA device visible: the GPU tests cannot fall back to the CPU")
    return edm_b200


def make_biases(edm, port, tmp_path, name, n_dev):
    cfg = BIAS_CASES[name]
    f = write_edm(tmp_path, name + ".edm", cfg["text"])
    sublo, subhi = cfg["sub"]
    bo = port.Bias("port", f)
    bo.setup(cfg["T"], cfg["kB"])
    bo.subdivide(sublo, subhi, sublo, subhi, cfg["periodic"], cfg["skin"])
    devs = [edm.bias_from_edm(f, cfg["T"], cfg["kB"], sublo, subhi, sublo, subhi, cfg["periodic"], cfg["skin"])
            for _ in range(n_dev)]
    return cfg, bo, devs


@pytest.mark.parametrize("name,nshards", [("c5_rdf_tight_limiter_backlog", 2), ("c2_rdf_threshold_tempering", 3),
                                          ("c3_2d_local_tempering_sparse", 2), ("c4_3d_density", 4),
                                          ("2d_limiter_cuts_the_round", 2)])
def test_sharded_blocks_commit_matches_single_rank_oracle(edm, port, tmp_path, name, nshards):
    import torch
    L = edm.lib()
    cfg, bo, devs = make_biases(edm, port, tmp_path, name, nshards + 1)
    main, shards = devs[0], devs[1:]
    D = bo.dim
    n = cfg["n"]
    cap = 1024
    bw = L.edm_hill_block_doubles(D, cap)
    blocks = torch.zeros(bw * nshards, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    bounds = np.linspace(0, n, nshards + 1).astype(int)
    bounds[1] = max(1, bounds[1] - 7)           # ragged shards
    for step in range(cfg["steps"]):
        x = np.ascontiguousarray(rng.uniform(cfg["lo"], cfg["hi"], size=(n, 3)))
        u = rng.uniform(0, 1, n)
        xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
        for s, b in enumerate(shards):
            lo, hi = int(bounds[s]), int(bounds[s + 1])
            b.pre_add_hill(n)                    # fresh candidate list on this "rank"
            edm.check(L.edm_bias_select_dev(b.h, hi - lo, xd[lo:].data_ptr(), 3, ud[lo:].data_ptr(), None, -1, n, 0, step,
                                            lo, st))
            edm.check(L.edm_bias_hills_pack_dev(b.h, blocks[s * bw:].data_ptr(), cap, st))
        edm.check(L.edm_bias_hills_commit_dev(main.h, blocks.data_ptr(), nshards, cap, n, st))
        torch.cuda.synchronize()
        main.check()
        # block contents: counts add up to what the oracle accepts, centres are the accepted candidates in order
        host = blocks.cpu().numpy()
        counts = [int(host[s * bw]) for s in range(nshards)]
        if "hill_density" in cfg["text"]:
            dens = float(cfg["text"].split("hill_density")[1].split()[0])
            acc = u < dens / n
            assert sum(counts) == int(acc.sum())
            got = np.concatenate([host[s * bw + 1: s * bw + 1 + counts[s] * D].reshape(-1, D) for s in range(nshards)])
            assert np.array_equal(got, x[acc][:, :D])
        bo.add_hills(x, u, -1)
    log = compare_bias(main, bo)
    assert len(log) > 0


def test_pack_reports_a_block_that_is_too_small(edm, port, tmp_path):
    """Accepted hills beyond the block capacity are an error (EDM_ERR_CAPACITY), never a silent truncation."""
    import torch
    L = edm.lib()
    cfg, bo, (b,) = make_biases(edm, port, tmp_path, "c2_rdf_threshold_tempering", 1)
    n = cfg["n"]
    rng = np.random.default_rng(5)
    x = torch.from_numpy(np.ascontiguousarray(rng.uniform(2.0, 4.5, size=(n, 3)))).cuda()
    u = torch.from_numpy(rng.uniform(0, 1, n)).cuda()
    cap = 8
    block = torch.zeros(L.edm_hill_block_doubles(1, cap), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    b.pre_add_hill(n)
    edm.check(L.edm_bias_select_dev(b.h, n, x.data_ptr(), 3, u.data_ptr(), None, -1, n, 0, 0, 0, st))
    edm.check(L.edm_bias_hills_pack_dev(b.h, block.data_ptr(), cap, st))
    torch.cuda.synchronize()
    assert int(block[0].item()) == cap
    with pytest.raises(edm.EdmError, match="exhausted"):
        b.check()


@pytest.mark.parametrize("attached", [False, True])
def test_library_exchange_single_rank(edm, port, tmp_path, attached):
    """edm_bias_exchange_dev (explicit) and edm_bias_set_comm (attached: add_hills exchanges by itself) on a
    1-rank communicator: pack -> commit inside the library, against the oracle."""
    import torch
    L = edm.lib()
    name = "c5_rdf_tight_limiter_backlog"
    cfg, bo, (b,) = make_biases(edm, port, tmp_path, name, 1)
    comm = edm.Comm.init_rank(bytes(128), 1, 0, 0)
    assert comm.info() == dict(nranks=1, rank=0, device=0)
    n = cfg["n"]
    rng = np.random.default_rng(11)
    st = torch.cuda.current_stream().cuda_stream
    if attached:
        b.set_comm(comm, 2048)
    for step in range(cfg["steps"]):
        x = np.ascontiguousarray(rng.uniform(cfg["lo"], cfg["hi"], size=(n, 3)))
        u = rng.uniform(0, 1, n)
        if attached:
            b.add_hills(x, u)
        else:
            xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
            b.pre_add_hill(n)
            edm.check(L.edm_bias_select_dev(b.h, n, xd.data_ptr(), 3, ud.data_ptr(), None, -1, n, 0, step, 0, st))
            edm.check(L.edm_bias_exchange_dev(b.h, comm.h, 2048, n, st))
            torch.cuda.synchronize()
            b.check()
        bo.add_hills(x, u, -1)
    compare_bias(b, bo)
    b.set_comm(None)
    comm.destroy()


def test_exchange_from_cpp(edm, tmp_path):
    """tests_host/exchange_test.cpp: EDM::EDMBias replicas on every visible device (ncclCommInitAll, a host thread
    per device) against one single-rank EDMBias over the concatenated shards; bit-identical replicas."""
    exe = os.path.join(PKG, "lib", "exchange_test")
    if not os.path.exists(exe):
        import importlib.util
        spec = importlib.util.spec_from_file_location("edm_b200_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_host_tests()
    r = subprocess.run([exe, str(tmp_path)], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:]
    assert "EXCHANGE_TEST_OK ranks=%d" % min(edm.device_count(), 8) in r.stdout
