"""Run under torchrun on N GPUs: every rank evaluates its own atoms' pairs, selects hills locally,
and calls the library's own exchange (edm_bias_exchange_dev on a communicator bootstrapped with
edm_comm_init_rank: over NVLink peer windows -- one kernel stores the hills into every peer and waits for
theirs -- or, with EDM_B200_NO_P2P=1, pack -> ncclAllGather -> commit; torch.distributed only ships
the 128-byte id and compares the replicas afterwards).  Checks:
replicas bit-identical across ranks; rank 0 equal (1e-10) to a single-rank oracle that sees the
rank-major concatenation of all ranks' pairs."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import edm_b200 as edm  # noqa: E402
import pyoracle  # noqa: E402

EDM_TEXT = ("tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.004\n"
            "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n")


def main():
    tmp = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [edm.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    comm = edm.Comm.init_rank(uid[0], world, rank, local)
    assert comm.info() == dict(nranks=world, rank=rank, device=local)
    if rank == 0:  # how the hills travel: NVLink peer windows (one kernel) or ncclAllGather (EDM_B200_NO_P2P=1)
        print("PEER_WINDOWS=%d" % int(comm.peer_windows()))
    d = os.path.join(tmp, "r%d" % rank)
    os.makedirs(d, exist_ok=True)
    f = os.path.join(d, "c.edm")
    open(f, "w").write(EDM_TEXT + "hills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (d, d))
    L = edm.lib()
    b = edm.bias_from_edm(f, 300.0, 0.0019872, [1.68], [5.0], [1.68], [5.0], [0], [0.0], device=local)
    n, box_len, rc, cap, steps, seed = 5000, 34.0, 5.0, 2048, 5, 11
    box = np.array([box_len] * 3)
    boxp = box.ctypes.data_as(C.POINTER(C.c_double))
    blk_n = L.edm_hill_block_doubles(1, cap)
    block = torch.zeros(blk_n, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(blk_n * world, dtype=torch.float64, device="cuda")
    fdev = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    edev = torch.zeros(1, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    est_total = 2 * 140000 * world
    xs = []
    for step in range(steps):
        x = np.ascontiguousarray(np.random.default_rng([seed, rank, step]).uniform(0, box_len, size=(n, 3)))
        xs.append(x)
        xd = torch.from_numpy(x).cuda()
        edm.check(L.edm_pair_select_cells_dev(b.h, n, xd.data_ptr(), fdev.data_ptr(), None, 0, 0, boxp, rc, est_total,
                                              seed + rank, step, edev.data_ptr(), st))
        edm.check(L.edm_bias_exchange_dev(b.h, comm.h, cap, est_total, st))
    torch.cuda.synchronize()
    b.check()
    v, dv = b.bias_grid.get_arrays()
    # replicas identical across ranks
    mine = torch.from_numpy(np.concatenate([v, dv.ravel()])).cuda()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(mine, ref)
    flags = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    assert int(flags.item()) == 1, "replicas differ across ranks"
    # gather every rank's positions on rank 0 for the single-rank oracle
    allx = [None] * world
    dist.all_gather_object(allx, xs)
    if rank == 0:
        fo = os.path.join(tmp, "o.edm")
        open(fo, "w").write(EDM_TEXT + "hills_filename %s/H\nhistogram_filename %s/G\n" % (tmp, tmp))
        bo = pyoracle.Bias("port", fo)
        bo.setup(300.0, 0.0019872)
        bo.subdivide([1.68], [5.0], [1.68], [5.0], [0], [0.0])
        for step in range(steps):
            rs, us = [], []
            for r in range(world):
                x = allx[r][step]
                pi, pj, sh = pyoracle.build_half_list(x, box, rc)
                force = np.zeros((n, 3))
                _, rr = bo.pair_step(pi, pj, x, force, shift=sh, do_hills=False)
                rs.append(rr)
                us.append(pyoracle.pair_uniforms(seed + r, step, pi, pj, n))
            bo.pre_add_hill(est_total)
            bo.add_hill_many(np.repeat(np.concatenate(rs), 2), np.concatenate(us))
            bo.post_add_hill()
        vo, do = bo.gauss.get_arrays()
        scale = np.abs(vo).max()
        assert np.abs(v - vo).max() <= 1e-10 * scale, np.abs(v - vo).max() / scale
        ld, lo = b.log(), bo.log()
        assert len(ld) == len(lo) and np.array_equal(ld["type"], lo["type"]) and np.array_equal(ld["pos"], lo["pos"])
        assert b.backlog()[:2] == bo.backlog()[:2]
        print("MULTI_GPU_CHECK_OK world=%d events=%d" % (world, len(ld)))
    coord_check(tmp, rank, world, local, comm)
    comm.destroy()
    dist.destroy_process_group()


COORD_TEXT = ("tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\n"
              "hill_density 200\ndimension 2\nbox_low 0 0\nbox_high 16 16\nbias_spacing 0.03125 0.03125\n"
              "bias_sigma 0.0625 0.0625\n")


def coord_check(tmp, rank, world, local, comm):
    """fix edm sharded: every rank owns a block of atoms (2-D coordinate CV, local well-tempering), selects
    with the job-wide est_hill_count and atom-index counters, and commits the all-gathered hills; checked
    against a single-rank oracle run on the rank-major concatenation."""
    d = os.path.join(tmp, "c%d" % rank)
    os.makedirs(d, exist_ok=True)
    f = os.path.join(d, "c.edm")
    open(f, "w").write(COORD_TEXT + "hills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (d, d))
    L = edm.lib()
    b = edm.bias_from_edm(f, 300.0, 0.0019872, [0, 0], [16, 16], [0, 0], [16, 16], [1, 1], [0.0, 0.0], device=local)
    n, cap, steps = 20000, 1024, 4
    blk_n = L.edm_hill_block_doubles(2, cap)
    block = torch.zeros(blk_n, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(blk_n * world, dtype=torch.float64, device="cuda")
    fdev = torch.zeros((n, 2), dtype=torch.float64, device="cuda")
    edev = torch.zeros(1, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    xs, us, es, fs = [], [], [], []
    side = torch.cuda.Stream(priority=-1)
    ev_fork, ev_k1, ev_join = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    for step in range(steps):
        rng = np.random.default_rng([77, rank, step])
        x = np.ascontiguousarray(rng.uniform(-1, 17, size=(n, 2)))
        u = rng.uniform(0, 1, n)
        xs.append(x)
        us.append(u)
        xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
        fdev.zero_()
        # the application's pattern: force update on the main stream; selection, exchange and the round on a
        # side stream, whose deposit waits (edm_bias_round_after) for the force update to finish reading
        main = torch.cuda.current_stream()
        ev_fork.record(main)
        side.wait_event(ev_fork)
        edm.check(L.edm_bias_update_forces_dev(b.h, n, xd.data_ptr(), 2, fdev.data_ptr(), 2, None, -1, edev.data_ptr(), st))
        ev_k1.record(main)
        with torch.cuda.stream(side):
            sst = side.cuda_stream
            edm.check(L.edm_bias_select_dev(b.h, n, xd.data_ptr(), 2, ud.data_ptr(), None, -1, n * world, 0, step,
                                            rank * n, sst))
            edm.check(L.edm_bias_round_after(b.h, ev_k1.cuda_event))
            edm.check(L.edm_bias_exchange_dev(b.h, comm.h, cap, n * world, sst))
            ev_join.record(side)
        main.wait_event(ev_join)
        es.append(float(edev.item()))
        fs.append(fdev.cpu().numpy().copy())
    torch.cuda.synchronize()
    v, dv = b.bias_grid.get_arrays()
    mine = torch.from_numpy(np.concatenate([v, dv.ravel()])).cuda()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    flags = torch.tensor([1 if torch.equal(mine, ref) else 0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    assert int(flags.item()) == 1, "2-D replicas differ across ranks"
    allx, allu, alle, allf = [None] * world, [None] * world, [None] * world, [None] * world
    dist.all_gather_object(allx, xs)
    dist.all_gather_object(allu, us)
    dist.all_gather_object(alle, es)
    dist.all_gather_object(allf, fs)
    if rank == 0:
        fo = os.path.join(tmp, "oc.edm")
        open(fo, "w").write(COORD_TEXT + "hills_filename %s/HC\nhistogram_filename %s/GC\n" % (tmp, tmp))
        bo = pyoracle.Bias("port", fo)
        bo.setup(300.0, 0.0019872)
        bo.subdivide([0, 0], [16, 16], [0, 0], [16, 16], [1, 1], [0.0, 0.0])
        for step in range(steps):
            x = np.ascontiguousarray(np.concatenate([allx[r][step] for r in range(world)]))
            u = np.concatenate([allu[r][step] for r in range(world)])
            x3 = np.zeros((x.shape[0], 3))
            x3[:, :2] = x
            fo3 = np.zeros_like(x3)
            eo = bo.update_forces(x3, fo3, -1)
            ed = sum(alle[r][step] for r in range(world))
            fd = np.concatenate([allf[r][step] for r in range(world)])
            if step:
                assert abs(ed - eo) <= 1e-10 * abs(eo), (ed, eo)
                assert np.abs(fd - fo3[:, :2]).max() <= 1e-10 * np.abs(fo3).max()
            bo.add_hills(x3, u, -1)
        vo, do = bo.gauss.get_arrays()
        assert np.abs(v - vo).max() <= 1e-10 * np.abs(vo).max(), np.abs(v - vo).max() / np.abs(vo).max()
        ld, lo = b.log(), bo.log()
        assert len(ld) == len(lo) and np.array_equal(ld["pos"], lo["pos"])
        assert np.abs(ld["height"] - lo["height"]).max() <= 1e-10 * np.abs(lo["height"]).max()
        print("MULTI_GPU_COORD_CHECK_OK world=%d events=%d rounds=%s" % (world, len(ld), b.round_info()))


if __name__ == "__main__":
    main()
