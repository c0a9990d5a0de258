"""Top CUDA source lines by warp-stall samples of the kernel in an .ncu-rep captured with --import-source on
(-lineinfo build).  Per line: samples, share, the dominant stall reasons.

    python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep [N]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = {}
    hdr, fname = None, ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():   # a CUDA source line row (its SASS rows follow with an empty first column)
            cs = hdr.index("Warp Stall Sampling (All Samples)")
            try:
                v = float(r[cs])
            except ValueError:
                continue
            if v <= 0:
                continue
            stalls = {h: float(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h and r[i] not in ("", "-")}
            key = (fname, int(r[0]))
            e = agg.setdefault(key, [0.0, r[1].strip()[:110], {}, 0.0])
            e[0] += v
            e[3] += float(r[hdr.index("Instructions Executed")] or 0)
            for k, x in stalls.items():
                e[2][k] = e[2].get(k, 0.0) + x
    tot = sum(e[0] for e in agg.values())
    print("# %s: %d warp-stall samples" % (rep, tot))
    for (f, l), e in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join("%s %.0f" % (k.replace("stall_", ""), x) for k, x in sorted(e[2].items(), key=lambda kv: -kv[1])[:3] if x > 0)
        print("%7.0f %5.1f%%  %s:%-5d %s   [inst %d; %s]" % (e[0], 100 * e[0] / tot, f, l, e[1], e[3], why))


if __name__ == "__main__":
    main()
