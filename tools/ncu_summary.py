"""Condenses an .ncu-rep (one kernel launch, --set full) into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_selected.txt
    python tools/ncu_summary.py gpurun_out/x.ncu-rep --roofline WORKLOAD KERNEL CSRC_HASH SOURCE.txt
        additionally records the kernel's DRAM traffic and busiest resource in profiles/ncu_roofline.json under
        the hash of the csrc/ tree the capture was taken on (bench.py reports them only while that hash matches
        the build it is measuring; the hash is printed by `python bench.py --csrc-hash` ON THE SNAPSHOT that ran).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_atom.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    u = unit.strip().lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def unit_hash(hash_file_or_value, kernel):
    """The hash bench.py compares against for `kernel`: the line of its translation unit in the file written by
    `python bench.py --csrc-hash` on the snapshot that ran (first line = whole tree, kept for old files)."""
    if not os.path.exists(hash_file_or_value):
        return hash_file_or_value
    sys.path.insert(0, ROOT)
    import bench
    unit = bench.kernel_unit(kernel)
    lines = open(hash_file_or_value).read().split("\n")
    for ln in lines[1:]:
        parts = ln.split()
        if len(parts) == 2 and parts[0] == unit:
            return parts[1]
    return lines[0].strip()


def record_roofline(hdr, units, vals, workload, kernel, csrc, source):
    csrc = unit_hash(csrc, kernel)
    get = lambda k: (vals[hdr.index(k)], units[hdr.index(k)]) if k in hdr else (None, None)
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    pct = {
        "lsu": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")[0],
        "issue": get("smsp__issue_active.avg.pct_of_peak_sustained_active")[0],
        "fp64": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")[0],
        "dram": None,
        "l2": get("lts__throughput.avg.pct_of_peak_sustained_elapsed")[0],
    }
    pct = {k: float(v.replace(",", "")) for k, v in pct.items() if v not in (None, "")}
    rdp, wrp = get("dram__bytes_read.sum.pct_of_peak_sustained_elapsed")[0], get("dram__bytes_write.sum.pct_of_peak_sustained_elapsed")[0]
    if rdp not in (None, "") and wrp not in (None, ""):  # read + write share the same pins: their percentages add
        pct["dram"] = float(rdp.replace(",", "")) + float(wrp.replace(",", ""))
    names = {"lsu": "L1/shared-memory data pipe (l1tex LSU wavefronts)", "issue": "warp issue slots",
             "fp64": "fp64 pipe", "dram": "HBM (dram throughput)", "l2": "L2 (lts throughput)"}
    top = max(pct, key=pct.get)
    path = os.path.join(ROOT, "profiles", "ncu_roofline.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data.setdefault(workload, {})[kernel] = {
        "csrc_hash": csrc, "source": source,
        "dram_bytes_per_launch": int(to_bytes(*rd) + to_bytes(*wr)),
        "dram_bytes_read": int(to_bytes(*rd)), "dram_bytes_write": int(to_bytes(*wr)),
        "gpu_time_us": float(get("gpu__time_duration.sum")[0].replace(",", "")) *
                       {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(get("gpu__time_duration.sum")[1].strip(), 1.0),
        "binding": {"bound": {"dram": "hbm"}.get(top, top), "name": names[top], "pct_of_peak": pct[top],
                    "all_pct": pct, "source": source},
    }
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    roof = sys.argv[sys.argv.index("--roofline") + 1:] if "--roofline" in sys.argv else None
    print("# %s  (ncu --set full --clock-control none, one launch; values as ncu reports them)" % rep)
    if roof:
        print("# sources %s (%s)" % (unit_hash(roof[2], roof[1]), roof[1]))
    for vals in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-82s %s %s" % (w, vals[i], units[i]))
        print()
    if roof:
        record_roofline(hdr, units, rows[2], *roof)


if __name__ == "__main__":
    main()
