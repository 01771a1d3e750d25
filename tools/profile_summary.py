"""Summarise ncu captures into profiles/ (tracked): launch shares, key metrics, hottest source lines.

    python tools/profile_summary.py <tag> <launches.csv> <full.ncu-rep> [kernel-regex]

Writes profiles/<tag>_launches.md, profiles/<tag>_metrics.csv, profiles/<tag>_lines.txt and updates
profiles/traffic.json (per-launch DRAM bytes of the step kernel, read by bench.py).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    out = ["# {} — launch list (ncu --metrics gpu__time_duration.sum --clock-control none)".format(tag), "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append("| `{}` | {} | {:.1f} | {:.3f} | {:.2f} |".format(name, cnt[name], v / 1e3, v / total,
                                                                  v / cnt[name] / 1e3))
    open(os.path.join(PROF, tag + "_launches.md"), "w").write("\n".join(out) + "\n")
    return tot, cnt


def metrics(tag, rep, which):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [["kernel", "metric", "value", "unit"]]
    dram = None
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                out.append([name[:60], key, r[i], units[i]])
        rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
        wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        dram = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
    with open(os.path.join(PROF, tag + "_metrics.csv"), "w", newline="") as fh:
        csv.writer(fh).writerows(out)
    if dram is not None:
        tpath = os.path.join(PROF, "traffic.json")
        data = json.load(open(tpath)) if os.path.exists(tpath) else {}
        data[which] = dram
        json.dump(data, open(tpath, "w"), indent=1)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "40"],
                           capture_output=True, text=True).stdout
    open(os.path.join(PROF, tag + "_lines.txt"), "w").write(lines)


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    which = sys.argv[4] if len(sys.argv) > 4 else "tcgen05"
    os.makedirs(PROF, exist_ok=True)
    launches(tag, lcsv)
    metrics(tag, rep, which)
    print("wrote profiles/{}_*".format(tag))
