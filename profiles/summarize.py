"""Turns gpurun_out/ ncu artefacts into the committed summaries under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.md>     # per-kernel share of a run
    python profiles/summarize.py full <report.ncu-rep> <out.csv>      # key metrics per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "").replace("st::", "").replace("<unnamed>::", "")


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) > vi and r[vi]:
            a = agg.setdefault(short(r[ki]), [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", ""))
            n += 1
    tot = sum(v for _, v in agg.values())
    with open(out, "w") as f:
        f.write(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k[:70]}` | {c} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |\n")
        f.write(f"\n{n} launches, {tot / 1e6:.2f} ms of kernel time (ncu-serialised, cold cache: compare shares, not absolutes)\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    idx = [i for i, h in enumerate(hdr) if h in KEYS]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        for r in rows[1:]:
            w.writerow([short(r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
