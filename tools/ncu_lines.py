"""Rank CUDA source lines of an ncu report by warp-stall samples.

    python tools/ncu_lines.py report.ncu-rep [top_n]
Needs -lineinfo at compile time and --import-source on at capture time.
"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter()); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try: n = int(r[hdr.index("# Samples")])
    except ValueError: continue
    key = (cur_file, r[0])
    if r[1].strip(): src[key] = r[1].strip()
    agg[key]["samples"] += n
    agg[key]["inst"] += int(r[hdr.index("Instructions Executed")] or 0)
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: agg[key][h] += int(r[i])
            except ValueError: pass
# the combined view repeats lines once per SASS instruction: dedupe happened via sum of per-SASS rows
tot = sum(v["samples"] for v in agg.values())
print("total samples", tot)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    stalls = sorted(((h[6:], c) for h, c in v.items() if h.startswith("stall_") and c), key=lambda t: -t[1])[:3]
    print("%5.1f%% %-14s:%-4s inst=%-8d %-60s %s" % (100.0 * v["samples"] / max(tot, 1), key[0], key[1], v["inst"],
          src.get(key, "")[:60], " ".join("%s=%d" % s for s in stalls)))
