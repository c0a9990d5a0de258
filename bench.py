#!/usr/bin/env python
"""bench.py — EDM per-timestep bias engine on B200: CV bias+force evaluations/s and hills/s.

Workload (BASELINE.json configs[1], "c2_pair_rdf"): 1-D pair-distance EDM, bias box 1.68-5.0,
spacing 0.00025 (13 281 grid points), sigma 0.025, hill_density 250, threshold tempering;
10^6 synthetic atoms per GPU, uniform in a periodic cube at number density 0.1, every pair with
minimum-image distance < 5.0 evaluated once (about 2.6e7 pairs per GPU per step).

One step = one MD step of fix edm_pair with hill addition: cell binning, pair search, bias
energy + force per pair with force scatter, two hill proposals per pair, selection, (N > 1:
all-gather of the accepted hills), height scaling, bias_per_step limiter, deposit.

  value  evaluations/s over all ranks with positions and forces already resident in HBM
  e2e    the same step through the host-buffer C ABI call (edm_pair_step_cells): positions and
         forces copied host->device from pinned memory and forces + result copied back, every step
  --impl reference   the reference's own CPU code (oracle/_ref, unmodified lib/ compiled here) running the
         SAME step on the SAME 10^6-atom configuration on the host cores: the half neighbour list is built
         once (untimed, as LAMMPS would hand it over), its rows are sharded over P single-rank EDMBias
         instances, one per core, each with a full grid replica (the reference's scaling model for a
         replicated grid), and every step runs fix edm_pair's own loop — r = sqrt, update_force, +-scatter,
         two add_hill per pair, post_add_hill (lammps/fix_edm_pair.cpp:173-247)
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# Peer-window exchange: how long a rank's exchange kernel waits for a peer's hills before it reports EDM_ERR_COMM
# (library default 10 s).  Ranks of a benchmark run can sit in set-up code far apart; give them a minute.
os.environ.setdefault("EDM_B200_PEER_TIMEOUT", "60")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "electronic-dance-music_b200", "python"))

EDM_TEXT = ("tempering 1\nglobal_tempering 2.0\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.1\n"
            "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n")
N_ATOMS = 1_000_000
DENSITY = 0.1
CUTOFF = 5.0
TEMPERATURE, BOLTZ = 300.0, 0.0019872
PREWARM_HILLS = 20000        # hills deposited before timing so the evaluated bias is not flat
DEPOSIT_BATCH = 1 << 20      # hills in the batched-deposit throughput measurement
HILL_CAP = 4096              # records per rank in the exchange block
ALG_BYTES_PER_ATOM = 76      # 24 B x + 48 B f read-modify-write + 4 B type, per atom per step
SEED = 20261018
# ncu-derived numbers (DRAM traffic of the dominant kernel, the resource that binds it) are never typed in here:
# tools/ncu_summary.py --roofline writes them to this file together with a hash of the csrc/ tree they were
# measured on, and they are reported only while that hash matches the tree being benchmarked.
NCU_ROOFLINE_JSON = os.path.join(ROOT, "profiles", "ncu_roofline.json")


def _csrc_dir():
    return os.path.join(ROOT, "electronic-dance-music_b200", "csrc")


def kernel_unit(kernel):
    """The .cu file that defines `kernel` (template arguments ignored), or None."""
    base = kernel.split("<")[0]
    d = _csrc_dir()
    for name in sorted(os.listdir(d)):
        if name.endswith(".cu") and (" " + base + "(") in open(os.path.join(d, name)).read():
            return name
    return None


def csrc_hash(kernel=None):
    """sha1 over CUDA sources (file names + contents, sorted): the whole csrc/ tree, or -- for a kernel -- the
    translation unit that defines it plus every header, i.e. exactly what that kernel is compiled from."""
    d = _csrc_dir()
    unit = kernel_unit(kernel) if kernel else None
    h = hashlib.sha1()
    for name in sorted(os.listdir(d)):
        if name.endswith((".cuh", ".h")) or (name.endswith(".cu") and (unit is None or name == unit)):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def ncu_roofline(workload, kernel):
    """(traffic bytes per launch or None, binding-resource dict, note) from the last ncu capture of `kernel`."""
    try:
        rec = json.load(open(NCU_ROOFLINE_JSON))[workload][kernel]
    except Exception:
        return None, None, "no ncu capture on record for this kernel (profiles/ncu_roofline.json)"
    if rec.get("csrc_hash") != csrc_hash(kernel):
        return None, None, ("ncu capture on record (%s) was taken on sources %s, this build's are %s: not reported"
                            % (rec.get("source"), rec.get("csrc_hash"), csrc_hash(kernel)))
    return rec.get("dram_bytes_per_launch"), rec.get("binding"), rec.get("source")


def c2_config(workload, world, pairs_per_step):
    """The `config` object of both arms (ours and --impl reference): one definition, so they cannot drift."""
    return {"workload": workload, "atoms_per_gpu": N_ATOMS, "number_density": DENSITY, "cutoff": CUTOFF,
            "pairs_per_gpu_per_step": int(pairs_per_step), "grid_points": 13281, "hill_density": 250,
            "step": "fix edm_pair post_force with hill addition: every pair inside the cutoff evaluated once, force "
                    "scatter, two hill proposals per pair, limiter, deposit",
            "parallelism": "atoms sharded %d-way, grid replicated, hills all-gathered" % world}


ONE_BOX_ATOMS = 8_000_000   # c2_one_box: ONE periodic system of this many atoms, cut into z-slabs over the ranks


def one_box_slab(rank, world):
    """This rank's share of the single 8e6-atom box (number density 0.1, the same coordinates on every rank):
    local atoms of the slab z in [zlo, zhi) first, then the ghost atoms within the cutoff of its two faces
    (periodic images shifted), as LAMMPS hands them to fix edm_pair.  Returns (x, nlocal, lo, hi, periodic)."""
    L = (ONE_BOX_ATOMS / DENSITY) ** (1.0 / 3.0)
    x = np.random.default_rng(1234 + 7).uniform(0, L, size=(ONE_BOX_ATOMS, 3))
    if world == 1:
        return np.ascontiguousarray(x), ONE_BOX_ATOMS, [0.0] * 3, [L] * 3, [1, 1, 1]
    zlo, zhi = rank * L / world, (rank + 1) * L / world
    z = x[:, 2]
    rows = [x[(z >= zlo) & (z < zhi)]]
    nlocal = rows[0].shape[0]
    for shift in (-L, 0.0, L):
        zs = z + shift
        g = ((zs >= zlo - CUTOFF) & (zs < zlo)) | ((zs >= zhi) & (zs < zhi + CUTOFF))
        if g.any():
            xg = x[g].copy()
            xg[:, 2] += shift
            rows.append(xg)
    return np.ascontiguousarray(np.concatenate(rows)), nlocal, [0.0, 0.0, zlo - CUTOFF], [L, L, zhi + CUTOFF], [1, 1, 0]


def c2_positions(rank, n_sets):
    """Synthetic coordinates of one rank: uniform in a periodic cube at number density 0.1."""
    box_len = (N_ATOMS / DENSITY) ** (1.0 / 3.0)
    rng = np.random.default_rng(1234 + 1 + 1000 * rank)
    return box_len, [rng.uniform(0, box_len, size=(N_ATOMS, 3)) for _ in range(n_sets)]


# The other BASELINE.json configs (parity-test cases first; measurable on request with --workload).
# Coordinate workloads (fix edm): one step = update_forces over every atom + add_hills (stride 1).
COORD_WORKLOADS = {
    "c1_coord_1d": dict(
        text="tempering 0\nhill_prefactor 0.25\ndimension 1\nbox_low 0\nbox_high 10\nbias_spacing 0.009765625\n"
             "bias_sigma 0.025\nhill_density 250\n",
        T=1.0, kB=1.0, lo=[0.0], hi=[10.0], atoms=100_000, atoms_scale_with_gpus=False, prewarm=2000, warm_h=0.001),
    "c3_coord_2d": dict(
        text="tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\n"
             "hill_density 250\ndimension 2\nbox_low 0 0\nbox_high 64 64\nbias_spacing 0.015625 0.015625\n"
             "bias_sigma 0.0625 0.0625\n",
        T=300.0, kB=0.0019872, lo=[0.0, 0.0], hi=[64.0, 64.0], atoms=10_000_000, atoms_scale_with_gpus=False,
        prewarm=20000, warm_h=0.02 / 250),
    "c4_coord_3d": dict(
        text="tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 250\ndimension 3\nbox_low 0 0 0\n"
             "box_high 64 64 64\nbias_spacing 0.125 0.125 0.125\nbias_sigma 0.25 0.25 0.25\n",
        T=300.0, kB=0.0019872, lo=[0.0] * 3, hi=[64.0] * 3, atoms=10_000_000, atoms_scale_with_gpus=False,
        prewarm=20000, warm_h=0.02 / 250),
}
C2_LOCAL_TEXT = ("tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.1\n"
                 "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n")
C5_TEXT = ("tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.0002\n"
           "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n")


def write_edm(tmpdir, text=None):
    f = os.path.join(tmpdir, "bench.edm")
    with open(f, "w") as fh:
        fh.write((text or EDM_TEXT) + "hills_filename %s/HILLS\nhistogram_filename %s/HIST\n" % (tmpdir, tmpdir))
    return f


def prewarm_hills(rng):
    c = rng.uniform(1.68, 5.0, PREWARM_HILLS)
    h = np.full(PREWARM_HILLS, 0.02 / 250)
    return c, h


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.skip = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:   # nvidia-smi needs a moment to start
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples taken before this call are not part of the measured region."""
        self.skip = len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines[self.skip:]:
            p = [t.strip() for t in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI device), so the pinned host
    buffers of the end-to-end path are first-touched on the GPU's own NUMA node.  With one process per GPU and
    eight GPUs on two sockets, half of the host<->device copies otherwise cross the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        dev = "%s:%s" % (dom[-4:], rest)
        path = "/sys/bus/pci/devices/%s/local_cpulist" % dev.lower()
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# ------------------------------------------------------------------ reference arm / CPU baseline

def _farm_worker(conn, kind, edm_text, tmp, k, core, lo, hi, est, shared):
    """One single-rank reference instance pinned to one core, owning pairs [lo, hi) of the half list."""
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    d = os.path.join(tmp, "w%d" % k)
    os.makedirs(d, exist_ok=True)
    b = pyoracle.Bias(kind, write_edm(d, edm_text))
    b.log_enable(False)
    b.setup(TEMPERATURE, BOLTZ)
    b.subdivide([1.68], [5.0], [1.68], [5.0], [0], [0.0])
    warm_c, warm_h = shared["warm"]
    b.gauss.add_values(warm_c[:2000], warm_h[:2000])
    x, box = shared["x"], shared["box"]
    pi, pj, img = shared["pi"][lo:hi], shared["pj"][lo:hi], shared["img"][lo:hi]
    f = np.zeros_like(x)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg[0] == "stop":
            break
        _, do_hills, step = msg
        t, e, ncalls = b.time_fix_pair(pi, pj, img, box, x, f, do_hills, est, seed=SEED + k, step=step)
        conn.send((t, e, ncalls, hi - lo))
    conn.close()


class CpuFarm:
    """The reference's scaling model for a replicated grid: P single-rank EDMBias instances, one per core, the
    rows of ONE half neighbour list sharded over them (MPI itself is not in the image).  The list is built once,
    outside every timed region, as LAMMPS hands fix edm_pair a ready NeighList."""

    def __init__(self, kind, edm_text, x, box_len, cores, pair_limit=None):
        import multiprocessing as mp
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle
        t0 = time.perf_counter()
        pi, pj, img = pyoracle.build_half_list_fast(x, [box_len] * 3, CUTOFF)
        self.list_build_s = time.perf_counter() - t0
        self.pairs_total = int(pi.size)
        n = self.pairs_total if pair_limit is None else min(pair_limit, self.pairs_total)
        self.pairs = n
        self.cores = cores
        self.tmp = tempfile.mkdtemp()
        rng = np.random.default_rng(1234 + 1)
        shared = {"x": np.ascontiguousarray(x), "box": np.array([box_len] * 3), "pi": pi, "pj": pj, "img": img,
                  "warm": prewarm_hills(rng)}
        # equal pair counts, cut at row boundaries (a row = every listed partner of one atom i)
        cuts = [0]
        for k in range(1, cores):
            c = int(n * k / cores)
            while c < n and c > 0 and pi[c] == pi[c - 1]:
                c += 1
            cuts.append(max(c, cuts[-1]))
        cuts.append(n)
        # every instance thins with the job's proposal count: 2 per listed pair (fix_edm_pair.cpp:230-236, 245);
        # the reference's MPI build divides hill_density by mpi_size_ instead (lib/edm_bias.cpp:175-180) — same rate
        est = 2 * n
        avail = sorted(os.sched_getaffinity(0))
        ctx = mp.get_context("fork")
        self.procs, self.conns = [], []
        for k in range(cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_farm_worker, args=(b, kind, edm_text, self.tmp, k, avail[k % len(avail)], cuts[k],
                                                       cuts[k + 1], est, shared), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        for c in self.conns:
            assert c.recv() == "ready"

    def step(self, step, do_hills=True):
        """One MD step of fix edm_pair on every instance at once.  Returns (wall seconds on this host from release
        to the last instance's answer, slowest instance's own loop time, pairs evaluated, hill proposals)."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("step", 1 if do_hills else 0, step))
        res = [c.recv() for c in self.conns]
        wall = time.perf_counter() - t0
        return wall, max(r[0] for r in res), sum(r[3] for r in res), sum(r[2] for r in res)

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=10)


def _cpu_worker(args):
    """Bare EDMBias::update_force over pre-drawn pair distances (no list, no sqrt, no scatter, no hills): the
    secondary CPU figure."""
    kind, edm_file, r, warm_c, warm_h, repeats, core = args
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    b = pyoracle.Bias(kind, edm_file)
    b.log_enable(False)
    b.setup(TEMPERATURE, BOLTZ)
    b.subdivide([1.68], [5.0], [1.68], [5.0], [0], [0.0])
    b.gauss.add_values(warm_c[:2000], warm_h[:2000])
    return b.time_pair_eval(r, repeats)


def cpu_pair_rate(kind, edm_file, r, warm, cores, repeats):
    """evaluations/s of the reference's update_force over `r`, sharded over `cores` processes."""
    import multiprocessing as mp
    shards = np.array_split(r, cores)
    avail = sorted(os.sched_getaffinity(0))
    jobs = [(kind, edm_file, np.ascontiguousarray(s), warm[0], warm[1], repeats, avail[i % len(avail)])
            for i, s in enumerate(shards)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.map(_cpu_worker, jobs)
    # every instance runs concurrently; the job is done when the slowest shard is
    return r.size * repeats / max(times)


def cpu_hill_rate(kind, n_hills, warm_rng):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    g = pyoracle.GaussGrid(kind, 1, [1.68], [5.0], [0.00025], [0], 1, [0.025])
    c = warm_rng.uniform(1.68, 5.0, n_hills)
    h = np.full(n_hills, 1e-6)
    t = g.time_add_values(c, h)
    return n_hills / t


def oracle_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    if pyoracle.available("ref"):
        return "ref", "reference"
    if not pyoracle.available("port"):
        pyoracle.build(("port",))
    return "port", "port"


def sample_pair_distances(rng, n):
    """r of uniformly distributed pairs inside the cutoff sphere: p(r) ~ r^2 on [0, CUTOFF)."""
    return CUTOFF * rng.uniform(0, 1, n) ** (1.0 / 3.0)


def cpu_pair_loop(kind, edm_text, steps, warmup, cores=None, pair_limit=None):
    """Times the reference's fix edm_pair step on rank 0's first coordinate set.  Returns a dict."""
    cores = cores or len(os.sched_getaffinity(0))
    box_len, (x,) = c2_positions(0, 1)
    farm = CpuFarm(kind, edm_text, x, box_len, cores, pair_limit)
    try:
        for w in range(warmup):
            farm.step(w)
        walls, loops, pairs = [], [], 0
        for k in range(steps):
            wall, loop, n, _ = farm.step(warmup + k)
            walls.append(wall)
            loops.append(loop)
            pairs += n
    finally:
        farm.close()
    return {"evals_per_s": pairs / sum(walls), "ms_per_step": 1e3 * sum(walls) / steps,
            "slowest_instance_ms_per_step": 1e3 * sum(loops) / steps, "pairs_per_step": farm.pairs,
            "pairs_in_list": farm.pairs_total, "cores": cores, "list_build_s": farm.list_build_s}


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload != "c2_pair_rdf":
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm runs the benchmark workload c2_pair_rdf"}))
        return
    kind, kind_name = oracle_kind()
    cores = len(os.sched_getaffinity(0))
    res = cpu_pair_loop(kind, EDM_TEXT, args.steps, args.warmup, cores)
    value = res["evals_per_s"]
    rng = np.random.default_rng(1234 + 1)
    hills = cpu_hill_rate(kind, 20000, rng)
    sample = ("the full workload: all %d listed pairs of the 10^6-atom configuration (rank 0's first coordinate set) per "
              "step, list rows sharded over %d single-rank instances pinned one per core; fix edm_pair's own loop "
              "(sqrt, update_force, +-scatter, 2 add_hill per pair, post_add_hill); neighbour list built once, untimed "
              "(%.1f s)" % (res["pairs_per_step"], cores, res["list_build_s"]))
    out = {
        "impl": "reference", "metric": "CV bias+force evals/sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": c2_config(args.workload, args.gpus, res["pairs_in_list"]),
        "hills_per_s": hills,
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind_name, "sample": sample,
                         "slowest_instance_ms_per_step": res["slowest_instance_ms_per_step"],
                         "hills_per_s_1core": hills},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------ GPU arm

def run_gpu(args, rank, local_rank, world):
    import torch
    import ctypes as C
    import edm_b200 as edm

    if edm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        bind_to_gpu_numa_node(local_rank)
        # the hill exchange is the library's own (NCCL called from C++): torch.distributed only ships the 128-byte
        # id once, synchronises the ranks around the timed region and reduces the final statistics
        uid = [edm.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        comm = edm.Comm.init_rank(uid[0], world, rank, local_rank)
    else:
        comm = edm.Comm.init_rank(bytes(128), 1, 0, local_rank)
    L = edm.lib()
    tmp = tempfile.mkdtemp()
    edm_text = {"c5_pair_rdf_backlog": C5_TEXT, "c2_pair_rdf_local_tempering": C2_LOCAL_TEXT}.get(args.workload, EDM_TEXT)
    edm_file = write_edm(tmp, edm_text)
    bias = edm.bias_from_edm(edm_file, TEMPERATURE, BOLTZ, [1.68], [5.0], [1.68], [5.0], [0], [0.0], device=local_rank)
    warm_rng = np.random.default_rng(1234 + 1)
    warm = prewarm_hills(warm_rng)          # same hills on every rank: replicas start identical
    bias.bias_grid.add_values(*warm)
    edm.check(L.edm_bias_set_profiling(bias.h, 1))

    one_box = args.workload == "c2_one_box"
    dom = None
    if one_box:
        # ONE system over all ranks: z-slabs with ghost atoms (strong scaling of a fixed 8e6-atom box)
        n_sets = 1
        xs, nlocal, dlo, dhi, dper = one_box_slab(rank, world)
        sets = [xs]
        n_rows = xs.shape[0]
        box_len = dhi[0] - dlo[0]
        dom = edm.PairDomain((C.c_double * 3)(*dlo), (C.c_double * 3)(*dhi), (C.c_int * 3)(*dper), nlocal)
    else:
        n_sets = 3                           # rotate position sets; L2 is flushed between steps anyway
        box_len, sets = c2_positions(rank, n_sets)
        n_rows = nlocal = N_ATOMS
    box = np.array([box_len] * 3)
    xs_host = [torch.from_numpy(x).pin_memory() for x in sets]
    xs_dev = [x.cuda(non_blocking=True) for x in xs_host]
    f_dev = torch.zeros((n_rows, 3), dtype=torch.float64, device="cuda")
    f_host = torch.zeros((nlocal, 3), dtype=torch.float64).pin_memory()
    energy_dev = torch.zeros(1, dtype=torch.float64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    boxp = box.ctypes.data_as(C.POINTER(C.c_double))
    expected_pairs = nlocal * (4.0 / 3.0) * np.pi * CUTOFF ** 3 * DENSITY / 2
    est_local = int(2 * expected_pairs)
    seed = SEED

    def step_resident(step):
        x = xs_dev[step % n_sets]
        est_total = est_local * world
        if one_box:
            edm.check(L.edm_pair_select_cells_domain_dev(bias.h, n_rows, x.data_ptr(), f_dev.data_ptr(), None, 0, 0,
                                                         C.byref(dom), CUTOFF, est_total, seed + rank, step,
                                                         energy_dev.data_ptr(), stream))
        else:
            edm.check(L.edm_pair_select_cells_dev(bias.h, N_ATOMS, x.data_ptr(), f_dev.data_ptr(), None, 0, 0, boxp,
                                                  CUTOFF, est_total, seed + rank, step, energy_dev.data_ptr(), stream))
        # pack -> ncclAllGather -> commit inside the library (one rank: pack -> commit)
        edm.check(L.edm_bias_exchange_dev(bias.h, comm.h, HILL_CAP, est_total, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()   # the first exchange waits for every rank's hills: ranks that set up at different speeds meet here first
    clocks = ClockSampler(local_rank)
    clocks.start()
    step_no = 0
    for _ in range(max(args.warmup, 3)):
        step_resident(step_no)
        step_no += 1
    barrier()
    st0 = bias.state()
    launches0 = edm.launch_count()
    clocks.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pair_ms, find_ms, eval_ms = [], [], []
    barrier()
    for k in range(args.steps):
        flush.zero_()                         # evict L2 between timed iterations (outside the event pair)
        ev[k][0].record()
        step_resident(step_no)
        ev[k][1].record()
        step_no += 1
        ms = C.c_double(0)
        edm.check(L.edm_bias_profile_ms(bias.h, C.byref(ms)))   # waits for this step's pair kernel
        pair_ms.append(ms.value)
        ms_a, ms_b = C.c_double(0), C.c_double(0)
        edm.check(L.edm_bias_profile_pair_ms(bias.h, C.byref(ms_a), C.byref(ms_b)))
        find_ms.append(ms_a.value)
        eval_ms.append(ms_b.value)
    barrier()
    launches = edm.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    st1 = bias.state()
    hills_timed = len(bias.log())             # events logged since creation (prewarm deposits are not logged)

    bias.check()                              # nothing was dropped or overflowed in the timed rounds

    # ---- e2e through the host-buffer C ABI.  With the communicator attached the call exchanges the hills by
    # itself and takes this rank's est_hill_count, as fix edm_pair passes it.
    bias.set_comm(comm, HILL_CAP)

    def step_e2e(step):
        xh = xs_host[step % n_sets]
        r = edm.PairResult()
        if one_box:
            edm.check(L.edm_pair_step_cells_domain(bias.h, n_rows, xh.data_ptr(), f_host.data_ptr(), None, 0, 0,
                                                   C.byref(dom), CUTOFF, 1, est_local, seed + rank, step, C.byref(r)))
        else:
            edm.check(L.edm_pair_step_cells(bias.h, N_ATOMS, xh.data_ptr(), f_host.data_ptr(), None, 0, 0, boxp, CUTOFF,
                                            1, est_local, seed + rank, step, C.byref(r)))
        return r

    r0 = step_e2e(step_no)
    step_no += 1
    pairs_per_step = [0] * n_sets
    for s in range(n_sets):
        rr = step_e2e(step_no)
        pairs_per_step[step_no % n_sets] = rr.n_pairs
        step_no += 1
    barrier()
    e2e_steps = args.steps                    # the same K as the device-timed figure
    e2e_pairs = 0
    split = np.zeros(4)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        rr = step_e2e(step_no)
        e2e_pairs += rr.n_pairs
        step_no += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for k in range(3):                        # the device-side split, outside the timed loop (event queries synchronise)
        step_e2e(step_no)
        step_no += 1
        if not one_box:                       # the domain entry point is not instrumented
            v = [C.c_double(0) for _ in range(4)]
            edm.check(L.edm_bias_profile_e2e_ms(bias.h, *[C.byref(t) for t in v]))
            split += np.array([t.value for t in v]) / 3.0
    bias.set_comm(None)

    pairs_timed = sum(pairs_per_step[(st0["steps"] + k) % n_sets] for k in range(args.steps))

    # ---- batched deposit throughput (hills/s) on a scratch replica of the bias grid
    dep_grid = edm.GaussGrid(1, [1.68], [5.0], [0.00025], [0], 1, [0.025], device=local_rank)
    drng = np.random.default_rng(99 + rank)
    dc = torch.from_numpy(drng.uniform(1.68, 5.0, DEPOSIT_BATCH)).cuda()
    dh = torch.full((DEPOSIT_BATCH,), 1e-6, dtype=torch.float64, device="cuda")
    dba = torch.zeros(DEPOSIT_BATCH, dtype=torch.float64, device="cuda")
    dep_launch0 = edm.launch_count()
    for _ in range(2):
        edm.check(L.edm_gauss_deposit_dev(dep_grid.h, DEPOSIT_BATCH, dc.data_ptr(), dh.data_ptr(), dba.data_ptr(), stream))
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dep_reps = 3
    d0.record()
    for _ in range(dep_reps):
        edm.check(L.edm_gauss_deposit_dev(dep_grid.h, DEPOSIT_BATCH, dc.data_ptr(), dh.data_ptr(), dba.data_ptr(), stream))
    d1.record()
    torch.cuda.synchronize()
    dep_ms = d0.elapsed_time(d1) / dep_reps
    hills_per_s = DEPOSIT_BATCH / (dep_ms * 1e-3)
    clk = clocks.stop()   # covers the timed steps, the e2e steps and the deposit batches

    # ---- reduce over ranks: max time, summed work
    stats = torch.tensor([total_ms, e2e_s, float(pairs_timed), float(e2e_pairs), hills_per_s, float(launches)] +
                         list(split), dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_s = float(mx[0]), float(mx[1])
        split = mx[6:10].cpu().numpy()
        mn = stats.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        # hills are NOT sharded: every replica deposits every hill, so the job's rate is one replica's
        pairs_all, e2e_pairs_all, hills_all, launches_all = float(sm[2]), float(sm[3]), float(mn[4]), float(sm[5])
    else:
        pairs_all, e2e_pairs_all, hills_all, launches_all = float(pairs_timed), float(e2e_pairs), hills_per_s, float(launches)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = pairs_all / (total_ms * 1e-3)
        kernel_ms = float(np.mean(eval_ms)) if np.mean(eval_ms) > 0 else float(np.mean(pair_ms))
        alg_bytes = ALG_BYTES_PER_ATOM * n_rows
        info = bias.pair_search_info()
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        st_hills = st1["steps"] - st0["steps"]
        traffic, binding, traffic_src = ncu_roofline(args.workload, "block_eval_kernel")
        h2d, d2h = (n_rows + nlocal) * 24, nlocal * 24
        config = c2_config(args.workload, world, pairs_per_step[0])
        if one_box:
            config.update({"atoms_total": ONE_BOX_ATOMS, "atoms_per_gpu": nlocal, "rows_per_gpu_with_ghosts": n_rows,
                           "parallelism": "ONE periodic box cut into %d z-slabs, ghost atoms within the cutoff of the "
                                          "faces (newton off: a pair across a face is evaluated on both sides), grid "
                                          "replicated, hills all-gathered" % world})
        out = {
            "metric": "CV bias+force evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if one_box else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config,
            "timing": {"hill_rounds_timed": st_hills,
                       "l2": "flushed between timed steps (512 MiB memset outside the per-step CUDA-event pair)",
                       "exchange": ("edm_bias_exchange_dev inside the library: " +
                                    ("one kernel stores the accepted hills into every rank's NVLink peer window and "
                                     "waits for theirs -> commit" if (world > 1 and comm.peer_windows()) else
                                     "pack -> ncclAllGather -> commit" if world > 1 else "single rank, nothing travels"))},
            "hills_per_s": hills_all,
            "hills": {"batched_deposit_hills_per_s": hills_all, "batch": DEPOSIT_BATCH, "ms_per_batch": dep_ms,
                      "in_situ_hill_events": int(hills_timed),
                      "in_situ_hills_per_s": (st1["steps"] - st0["steps"]) * 250.0 / (total_ms * 1e-3) if total_ms else None,
                      "rounds": bias.round_info(),
                      "backlog": list(bias.backlog()[:2])},
            # achieved/peak/frac are the HBM figures the contract asks for (algorithmic bytes over the measured copy
            # peak); `bound` names the resource that actually binds the kernel, read from the ncu capture on record
            # for THIS build (null when there is none), never typed in
            "roofline": {"bound": (binding or {}).get("bound", "hbm"), "kernel": "block_eval_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "hbm_frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "search_kernel_ms": float(np.mean(find_ms)), "pair_kernels_ms": float(np.mean(pair_ms)),
                         "bricks": list(info["bricks"]), "fallbacks": info["fallbacks"],
                         "traffic_source": traffic_src, "binding_resource": binding, "csrc_hash": csrc_hash("block_eval_kernel"),
                         "note": "not HBM bound by design (SURVEY 8d: 2.9 B/pair of compulsory traffic)"},
            "e2e": {"value": e2e_pairs_all / e2e_s, "unit": "evals/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + 32,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "split_ms": {"x_h2d": float(split[0]), "kernels": float(split[1]), "f_d2h": float(split[2]),
                                 "device_span": float(split[3]),
                                 "host_overhead": 1e3 * e2e_s / e2e_steps - float(split[3]),
                                 "note": "max over ranks; the 24 MB f upload and the f download run on a copy stream "
                                         "beside the search and the hill round"},
                    "pcie_gbs": {"x_h2d": 24e-6 * N_ATOMS / float(split[0]) if split[0] > 0 else None,
                                 "f_d2h": 24e-6 * N_ATOMS / float(split[2]) if split[2] > 0 else None,
                                 "note": "slowest rank's achieved host-link rate; ranks share the host's root complexes"}},
            "gpu_launches": int(launches_all),
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == "c2_pair_rdf":
            kind, kind_name = oracle_kind()
            cores = len(os.sched_getaffinity(0))
            crng = np.random.default_rng(4321)
            res = cpu_pair_loop(kind, edm_text, 4, 1, cores)                    # the real loop, all cores
            res1 = cpu_pair_loop(kind, edm_text, 2, 1, 1, pair_limit=2_000_000)  # and one core on a 2e6-pair sample
            r = sample_pair_distances(crng, 1_000_000 * cores)
            bare = cpu_pair_rate(kind, edm_file, r, warm, cores, 2)
            hills = cpu_hill_rate(kind, 20000, crng)
            out["cpu_baseline"] = {
                "value": res["evals_per_s"], "unit": "evals/s", "cores": cores, "kind": kind_name,
                "sample": "4 steps of fix edm_pair's own loop (lammps/fix_edm_pair.cpp:173-247: sqrt, update_force, "
                          "+-scatter, 2 add_hill per pair, post_add_hill) over all %d listed pairs of this workload, list "
                          "rows sharded over %d single-rank instances, one per core; list built once, untimed"
                          % (res["pairs_per_step"], cores),
                "ms_per_step": res["ms_per_step"], "evals_per_s_1core": res1["evals_per_s"],
                "bare_update_force_evals_per_s": bare,
                "bare_update_force_note": "EDMBias::update_force alone over pre-drawn distances (no list walk, sqrt, "
                                          "scatter or hills), %d instances: the round-1 figure, kept for continuity" % cores,
                "hills_per_s_1core": hills}
        print(json.dumps(out))
    comm.destroy()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ GPU arm, coordinate workloads

def run_coord(args, rank, local_rank, world):
    """fix edm (lammps/fix_edm.cpp:134-162) on synthetic coordinates: K1 over every atom, then the hill round."""
    import torch
    import ctypes as C
    import edm_b200 as edm

    if edm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [edm.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, 0)
        comm = edm.Comm.init_rank(uid[0], world, rank, local_rank)
    else:
        comm = edm.Comm.init_rank(bytes(128), 1, 0, local_rank)
    cfg = COORD_WORKLOADS[args.workload]
    L = edm.lib()
    tmp = tempfile.mkdtemp()
    edm_file = write_edm(tmp, cfg["text"])
    D = len(cfg["lo"])
    bias = edm.bias_from_edm(edm_file, cfg["T"], cfg["kB"], cfg["lo"], cfg["hi"], cfg["lo"], cfg["hi"], [1] * D,
                             [0.0] * D, device=local_rank)
    geo = bias.bias_grid.info()
    n_pts = int(np.prod(geo["n"][:D]))
    warm_rng = np.random.default_rng(1234 + 3)
    wc = warm_rng.uniform(cfg["lo"][0], cfg["hi"][0], size=(cfg["prewarm"], D))
    bias.bias_grid.add_values(np.ascontiguousarray(wc), np.full(cfg["prewarm"], cfg["warm_h"]))

    n_atoms = cfg["atoms"] // world          # strong split of the config's atom count: BASELINE names the total
    rng = np.random.default_rng(1234 + 3 + 1000 * rank)
    n_sets = 2

    def ordered(x):
        """--input-order: how the caller's atom array is ordered.  random = as drawn (the default, the hard case);
        cell = sorted by grid cell, row-major (what an MD code with spatial sorting hands over); strip = sorted only
        by a coarse strip of the slowest dimension whose records fit L2 (what a device-side binning pass would buy)."""
        if args.input_order == "random":
            return x
        dx = np.array(geo["dx"][:D])
        cell = np.floor((x - np.array(cfg["lo"])) / dx).astype(np.int64)
        n = np.array(geo["n"][:D], dtype=np.int64)
        if args.input_order == "cell":
            key = cell[:, D - 1]
            for d in range(D - 2, -1, -1):
                key = key * n[d] + cell[:, d]
        else:
            rows_per_strip = max(1, int((48 << 20) // (32 * np.prod(n[:D - 1]))))   # ~48 MB of records per strip
            key = cell[:, D - 1] // rows_per_strip
        return np.ascontiguousarray(x[np.argsort(key, kind="stable")])

    xs_host = [torch.from_numpy(ordered(rng.uniform(cfg["lo"][0], cfg["hi"][0], size=(n_atoms, D)))).pin_memory()
               for _ in range(n_sets)]
    xs_dev = [x.cuda(non_blocking=True) for x in xs_host]
    f_dev = torch.zeros((n_atoms, D), dtype=torch.float64, device="cuda")
    f_host = torch.zeros((n_atoms, D), dtype=torch.float64).pin_memory()
    energy_dev = torch.zeros(1, dtype=torch.float64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    seed = SEED
    est_total = n_atoms * world

    def forces(step):
        x = xs_dev[step % n_sets]
        edm.check(L.edm_bias_update_forces_dev(bias.h, n_atoms, x.data_ptr(), D, f_dev.data_ptr(), D, None, -1,
                                               energy_dev.data_ptr(), stream))

    def hills(step):
        x = xs_dev[step % n_sets]
        if world == 1:
            edm.check(L.edm_bias_add_hills_dev(bias.h, n_atoms, x.data_ptr(), D, None, None, -1, seed, step, stream))
        else:
            edm.check(L.edm_bias_select_dev(bias.h, n_atoms, x.data_ptr(), D, None, None, -1, est_total, seed, step,
                                            rank * n_atoms, stream))
            edm.check(L.edm_bias_exchange_dev(bias.h, comm.h, HILL_CAP, est_total, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    side = torch.cuda.Stream(priority=-1)   # the round's small kernels must not queue behind the force update's CTAs
    ev_fork, ev_k1, ev_join = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

    def step_fused(step):
        """One step as an application would issue it: the round's read-only part beside the force update."""
        x = xs_dev[step % n_sets]
        if world == 1:
            edm.check(L.edm_bias_step_coords_dev(bias.h, n_atoms, x.data_ptr(), D, f_dev.data_ptr(), D, None, -1, 1, None,
                                                 seed, step, energy_dev.data_ptr(), stream))
            return
        main = torch.cuda.current_stream()
        ev_fork.record(main)
        side.wait_event(ev_fork)
        with torch.cuda.stream(side):     # the selection first: launched behind the force update it waits for SM slots
            sst = side.cuda_stream
            edm.check(L.edm_bias_select_dev(bias.h, n_atoms, x.data_ptr(), D, None, None, -1, est_total, seed, step,
                                            rank * n_atoms, sst))
        edm.check(L.edm_bias_update_forces_dev(bias.h, n_atoms, x.data_ptr(), D, f_dev.data_ptr(), D, None, -1, None, stream))
        ev_k1.record(main)
        with torch.cuda.stream(side):
            # the round's writers (deposit, tail) follow the force update on the main stream: nothing to join
            edm.check(L.edm_bias_round_commit_on(bias.h, main.cuda_stream))
            edm.check(L.edm_bias_energy_with_round(bias.h, energy_dev.data_ptr()))   # summed by an idle deposit CTA
            edm.check(L.edm_bias_exchange_dev(bias.h, comm.h, HILL_CAP, est_total, sst))

    barrier()   # the first exchange waits for every rank's hills: ranks that set up at different speeds meet here first
    clocks = ClockSampler(local_rank)
    clocks.start()
    step_no = 0
    for _ in range(max(args.warmup, 3)):
        step_fused(step_no)
        step_no += 1
    barrier()
    launches0 = edm.launch_count()
    info0 = bias.round_info()
    clocks.mark()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step_fused(step_no)
        ev[k][1].record()
        step_no += 1
    barrier()
    launches = edm.launch_count() - launches0
    info1 = bias.round_info()
    stamps_fused = bias.round_times_us()       # device-clock stamps of the last overlapped step
    xstamps_fused = bias.exchange_times_us()   # selection / exchange of that step, same clock
    total_ms = float(sum(e[0].elapsed_time(e[1]) for e in ev))
    # the two halves one after the other, for the breakdown only
    evb = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(5)]
    for k in range(5):
        flush.zero_()
        evb[k][0].record()
        forces(step_no)
        evb[k][1].record()
        hills(step_no)
        evb[k][2].record()
        step_no += 1
    barrier()
    k1_ms = [e[0].elapsed_time(e[1]) for e in evb]
    round_ms = [e[1].elapsed_time(e[2]) for e in evb]

    # ---- e2e through the host-buffer C ABI: edm_bias_step_coords (update_forces + add_hills) on pinned host arrays
    def step_e2e(step):
        dp = C.POINTER(C.c_double)
        xh = C.cast(xs_host[step % n_sets].data_ptr(), dp)
        e = C.c_double(0)
        edm.check(L.edm_bias_step_coords(bias.h, n_atoms, xh, D, C.cast(f_host.data_ptr(), dp), D, None, -1, 1, None,
                                         seed, step, C.byref(e)))
        return e.value

    # (N > 1: the communicator is attached, so the host-buffer call exchanges the hills itself, as the fixes do)
    if world > 1:
        bias.set_comm(comm, HILL_CAP)
    step_e2e(step_no)
    step_no += 1
    torch.cuda.synchronize()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        step_e2e(step_no)
        step_no += 1
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        bias.set_comm(None)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    # ---- batched deposit throughput (GaussGrid::add_value equivalents per second) on a scratch replica
    dep_hills_per_s = None
    if world == 1:
        nb = 1 << 17
        dgrid = edm.GaussGrid(D, cfg["lo"], cfg["hi"], list(geo["dx"][:D]), [1] * D, 1,
                              [float(v) for v in cfg["text"].split("bias_sigma")[1].split()[:D]], device=local_rank)
        drng = np.random.default_rng(77)
        dc = torch.from_numpy(drng.uniform(cfg["lo"][0], cfg["hi"][0], size=(nb, D))).cuda()
        dh = torch.full((nb,), 1e-6, dtype=torch.float64, device="cuda")
        dba = torch.zeros(nb, dtype=torch.float64, device="cuda")
        edm.check(L.edm_gauss_deposit_dev(dgrid.h, nb, dc.data_ptr(), dh.data_ptr(), dba.data_ptr(), stream))
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(2):
            edm.check(L.edm_gauss_deposit_dev(dgrid.h, nb, dc.data_ptr(), dh.data_ptr(), dba.data_ptr(), stream))
        d1.record()
        torch.cuda.synchronize()
        dep_ms = d0.elapsed_time(d1) / 2
        dep_hills_per_s = {"batched_deposit_hills_per_s": nb / (dep_ms * 1e-3), "batch": nb, "ms_per_batch": dep_ms,
                           "note": "edm_gauss_deposit_dev: n pre-selected hills in list order, one CTA per hill"}
        del dgrid
    clk = clocks.stop()

    stats = torch.tensor([total_ms, float(np.mean(k1_ms)), float(np.mean(round_ms)), float(launches)],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        total_ms, k1, rnd = float(mx[0]), float(mx[1]), float(mx[2])
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        launches_all = float(sm[3])
    else:
        k1, rnd, launches_all = float(stats[1]), float(stats[2]), float(launches)
    if rank == 0:
        peak, peak_src = measured_peak()
        evals = float(n_atoms) * world * args.steps
        grid_bytes = n_pts * (1 + D) * 8
        alg = 24 * D * n_atoms + min(n_atoms * (2 ** D) * (1 + D) * 8, grid_bytes)   # SURVEY 8(d), per launch
        achieved = alg / (k1 * 1e-3) / 1e9
        # the capture on record depends on how the atoms were ordered: random order under the workload's own name
        roof_key = args.workload if args.input_order == "random" else "%s@%s" % (args.workload, args.input_order)
        out = {
            "metric": "CV bias+force evals/sec", "value": evals / (total_ms * 1e-3), "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "input_order": args.input_order, "atoms_total": n_atoms * world,
                       "atoms_per_gpu": n_atoms,
                       "grid_points": n_pts, "grid_bytes_in_hbm": n_pts * (2 if D == 1 else 4) * 8, "hill_density": 250,
                       "l2": "flushed between timed steps (512 MiB memset outside the CUDA-event pairs)",
                       "parallelism": "atoms sharded %d-way, grid replicated, hills all-gathered" % world},
            "step_breakdown_ms": {"update_forces": k1, "hill_round": rnd,
                                  "note": "measured back to back; in the timed step the round's selection, exchange, plan, "
                                          "integrals and decision run beside update_forces on a second stream"},
            "round_stamps_us": {
                "legend": "us since the plan began: [0-6] plan phases, [7,8] decision, [9,10] first deposit taken / last "
                          "deposit done, [11,12] in-order kernel begin/end, [13] last deposit CTA left, [14] force update began, [15] in-order kernel resident",
                "overlapped_step": [round(float(v), 2) for v in stamps_fused],
                "overlapped_step_exchange": {
                    "legend": "same clock: selection began / its last CTA left / exchange kernel past its predecessor / "
                              "own block delivered to every peer / every peer's block in",
                    "us": [round(float(v), 2) for v in xstamps_fused]},
                "back_to_back": [round(float(v), 2) for v in bias.round_times_us()]},
            "batched_deposit": dep_hills_per_s,
            "hills": {"rounds_parallel": info1["parallel"] - info0["parallel"],
                      "rounds_split": info1["split"] - info0["split"],
                      "rounds_in_order": info1["in_order"] - info0["in_order"]},
            "roofline": {"bound": "hbm", "kernel": "forces_kernel<%d>" % D, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_roofline(roof_key, "forces_kernel<%d>" % D)[0],
                         "traffic_source": ncu_roofline(roof_key, "forces_kernel<%d>" % D)[2],
                         "binding_resource": ncu_roofline(roof_key, "forces_kernel<%d>" % D)[1],
                         "peak_source": peak_src,
                         "kernel_ms": k1, "algorithmic_bytes_per_launch": alg,
                         "note": "kernel_ms spans forces_kernel + the 1-CTA energy sum (CUDA events on the launch stream)"},
            "gpu_launches": int(launches_all),
            "clocks": clk,
        }
        if e2e_ms is not None:
            out["e2e"] = {"value": n_atoms * world / (e2e_ms * 1e-3), "unit": "evals/s",
                          "h2d_bytes_per_step": 2 * n_atoms * D * 8, "d2h_bytes_per_step": n_atoms * D * 8 + 8,
                          "ms_per_step": e2e_ms,
                          "note": "edm_bias_step_coords on pinned host arrays; bytes are per rank; max over ranks"}
        print(json.dumps(out))
    bias.check()
    comm.destroy()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--csrc-hash", action="store_true", help="print the hash of the csrc/ tree and of every translation unit, and exit")
    ap.add_argument("--input-order", default="random", choices=["random", "cell", "strip"],
                    help="coordinate workloads: order of the caller's atom array (see run_coord.ordered)")
    ap.add_argument("--workload", default="c2_pair_rdf",
                    choices=["c2_pair_rdf", "c5_pair_rdf_backlog", "c2_pair_rdf_local_tempering", "c2_one_box"] +
                    sorted(COORD_WORKLOADS),
                    help="c2_pair_rdf is the benchmark (BASELINE.json configs[1]); the others are the remaining configs")
    args = ap.parse_args()
    if args.csrc_hash:   # first line: the whole tree; then one line per translation unit (unit + headers)
        print(csrc_hash())
        for name in sorted(os.listdir(_csrc_dir())):
            if name.endswith(".cu"):
                h = hashlib.sha1()
                for n2 in sorted(os.listdir(_csrc_dir())):
                    if n2.endswith((".cuh", ".h")) or n2 == name:
                        h.update(n2.encode())
                        h.update(open(os.path.join(_csrc_dir(), n2), "rb").read())
                print(name, h.hexdigest()[:16])
        return
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        if args.workload in COORD_WORKLOADS:
            run_coord(args, rank, local_rank, world)
        else:
            run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
