"""Per-kernel SASS mnemonic counts of liblasso_b200.so (what proves tcgen05 / TMEM / TMA use):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM),
UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copy (cp.async.bulk), SYNCS = mbarrier ops,
FFMA / FFMA2 / DFMA = CUDA-core fp32 / packed fp32 / fp64 multiply-adds, LDL / STL = local memory (spills).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("LASSO_B200_LIB") or os.path.join(ROOT, "pytorch-lasso_b200", "csrc", "liblasso_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "DFMA",
        "HMMA", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
counts, order, cur, total = {}, [], None, collections.Counter()
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it, m.group(1))
        cur = cur.replace("(int)", "").replace("lasso::<unnamed>::", "").replace("lasso::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        if cur in counts:
            cur += " #%d" % len(order)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op == k or (k in ("LDTM", "STTM", "SYNCS", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA") and op.startswith(k)):
                counts[cur][k] += 1
                break
print("# SASS mnemonic counts per kernel, %s (cuobjdump -sass, sm_100a)" % os.path.basename(LIB))
print("%-58s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEYS))
for name in order:
    c = counts[name]
    print("%-58s %7d " % (name[:58], total[name]) + " ".join("%7d" % c[k] for k in KEYS))
