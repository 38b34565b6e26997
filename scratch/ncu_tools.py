"""Helpers for the ncu evidence under profiles/ (read here, on the CPU box, from files brought back in gpurun_out/):
    python scratch/ncu_tools.py launches <launch_list.csv> [top]     aggregate a `--metrics gpu__time_duration.sum` list by kernel
    python scratch/ncu_tools.py summary <file.ncu-rep> <out.csv>     the roofline-relevant columns of a `--set full` capture"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.max", "gpu__dram_throughput", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_active.avg")


def launches(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki])[:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-72s %5d %9.3f ms %5.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
    print("total %.3f ms in %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))


def summary(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or any(h.startswith(k) for k in KEYS)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in cols if i < len(r)])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
    else:
        summary(sys.argv[2], sys.argv[3])
