"""The N > 1 path.

CPU (gloo, world_size 2, runs in this container): the hill exchange — block packing, one all-gather,
rank-major commit — leaves both replicas bit-identical and equal to a single-rank run over the
rank-major concatenation of the shards (the parity oracle for P GPUs).
GPU (needs 2 devices, skipped otherwise): the same check with the CUDA path and NCCL.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EDM_TEXT = ("tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.004\n"
            "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def shard_inputs(seed, rank, step, n):
    rng = np.random.default_rng([seed, rank, step])
    return rng.uniform(0.5, 5.5, n), rng.uniform(0, 1, n)


def _gloo_worker(rank, world, port, tmp, n, steps, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))
    import torch
    import torch.distributed as dist
    import pyoracle
    from edm_b200 import exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = os.path.join(tmp, "r%d" % rank)
    os.makedirs(d, exist_ok=True)
    f = os.path.join(d, "c.edm")
    open(f, "w").write(EDM_TEXT + "hills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (d, d))
    b = pyoracle.Bias("port", f)
    b.setup(300.0, 0.0019872)
    b.subdivide([1.68], [5.0], [1.68], [5.0], [0], [0.0])
    cap = 1024
    est_total = n * world
    thresh = 250.0 / est_total          # lib/edm_bias.cpp:543 with the job-wide candidate count
    for step in range(steps):
        r, u = shard_inputs(7, rank, step, n)
        mine = r[u < thresh]            # local selection, candidate order kept
        blk = torch.from_numpy(exchange.pack_block(mine, 1, cap))
        allb = exchange.all_gather_blocks(blk).numpy()
        hills = exchange.unpack_blocks(allb, 1, cap)
        b.pre_add_hill(est_total)
        b.add_hill_many(hills, -np.ones(len(hills)))   # already selected: always accepted
        b.post_add_hill()
    v, dv = b.gauss.get_arrays()
    left, right, buf = b.backlog()
    np.savez(out % rank, grid=v, deriv=dv, backlog=buf, lr=[left, right], cum=b.params()["cum_bias"],
             types=b.log()["type"], pos=b.log()["pos"])
    dist.destroy_process_group()


def test_exchange_two_ranks_gloo(tmp_path, port):
    import torch.multiprocessing as mp
    world, n, steps = 2, 20000, 6
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_gloo_worker, args=(world, free_port(), str(tmp_path), n, steps, out), nprocs=world, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    for k in ("grid", "deriv", "backlog", "lr", "cum", "types", "pos"):
        assert np.array_equal(r0[k], r1[k]), "replicas differ in " + k
    # single-rank oracle over the rank-major concatenation with the job-wide est_hill_count
    f = tmp_path / "single.edm"
    f.write_text(EDM_TEXT + "hills_filename %s/H\nhistogram_filename %s/G\n" % (tmp_path, tmp_path))
    b = port.Bias("port", str(f))
    b.setup(300.0, 0.0019872)
    b.subdivide([1.68], [5.0], [1.68], [5.0], [0], [0.0])
    for step in range(steps):
        rs, us = zip(*[shard_inputs(7, rank, step, n) for rank in range(world)])
        b.pre_add_hill(n * world)
        b.add_hill_many(np.concatenate(rs), np.concatenate(us))
        b.post_add_hill()
    v, dv = b.gauss.get_arrays()
    assert np.array_equal(v, r0["grid"]) and np.array_equal(dv, r0["deriv"])
    assert b.params()["cum_bias"] == float(r0["cum"])
    assert np.array_equal(b.log()["type"], r0["types"]) and np.array_equal(b.log()["pos"], r0["pos"])
    assert {"h", "u", "b", "v"} <= set(chr(t) for t in r0["types"])      # the limiter was exercised
    left, right, buf = b.backlog()
    assert [left, right] == list(r0["lr"]) and np.array_equal(buf, r0["backlog"])


def test_block_layout_round_trip():
    sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))
    from edm_b200 import exchange
    rng = np.random.default_rng(0)
    for dim in (1, 2, 3):
        parts = [rng.uniform(size=(k, dim)) for k in (0, 5, 17)]
        blocks = np.concatenate([exchange.pack_block(p, dim, 32) for p in parts])
        assert blocks.size == 3 * exchange.block_doubles(dim, 32)
        assert np.array_equal(exchange.unpack_blocks(blocks, dim, 32), np.concatenate(parts))
    with pytest.raises(ValueError):
        exchange.pack_block(np.zeros((40, 1)), 1, 32)


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["peer_windows", "nccl_allgather"])
def test_two_gpu_replicas_match_single_rank_oracle(tmp_path, transport):
    """Both transports of the hill exchange against the single-rank oracle: NVLink peer windows (the default on one
    node) and the ncclAllGather path (EDM_B200_NO_P2P=1)."""
    import edm_b200
    if edm_b200.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = os.path.join(ROOT, "tests", "multi_gpu_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), script, str(tmp_path)]
    env = dict(os.environ, EDM_B200_NO_P2P="1" if transport == "nccl_allgather" else "0", EDM_B200_PEER_TIMEOUT="20")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MULTI_GPU_CHECK_OK" in r.stdout
    assert "MULTI_GPU_COORD_CHECK_OK" in r.stdout
    if transport == "nccl_allgather":
        assert "PEER_WINDOWS=0" in r.stdout
