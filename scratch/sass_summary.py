"""profiles/r02_sass_summary.txt: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / TMEM use (cuobjdump
-sass of the objects linked into libryolo.so) plus register / stack usage.  Run after build():  python scratch/sass_summary.py"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(REPO, "build")
MN = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "FADD2", "FMUL2", "FFMA2", "HMMA",
      "STG.E.EF", "RED.E", "ATOMS", "SHFL", "MUFU"]
out = ["# SASS evidence, round 2 (cuobjdump -sass of the sm_100a objects linked into libryolo.so; nvcc 12.9, -O3 -lineinfo)",
       "# counts of instruction mnemonics per kernel: UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA bulk tensor load /",
       "# store, UBLKCP = cp.async.bulk (1-D TMA copy, the BN row pipes), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR =",
       "# tcgen05.commit, SYNCS = mbarrier ops, FADD2 / FMUL2 / FFMA2 = packed fp32x2.",
       "# Regenerate: python scratch/sass_summary.py", ""]
for o in sorted(f for f in os.listdir(BUILD) if f.endswith(".o") and not f.startswith("rbox_oracle")):
    path = os.path.join(BUILD, o)
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    regs, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (m.group(1), m.group(2))
    counts, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            counts[cur]["_total"] += 1
            for k in MN:
                if m.group(1).startswith(k):
                    counts[cur][k] += 1
    out.append("## %s" % o)
    for fn, c in counts.items():
        dem = re.sub(r"\(.*", "", subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip())
        r = regs.get(fn, ("?", "?"))
        out.append("%-60s instr=%-6d regs=%-4s stack=%-4s | %s" % (dem[:60], c["_total"], r[0], r[1],
                                                                    " ".join("%s=%d" % (k, c[k]) for k in MN if c[k])))
    out.append("")
open(os.path.join(REPO, "profiles", "r02_sass_summary.txt"), "w").write("\n".join(out))
print("wrote profiles/r02_sass_summary.txt (%d lines)" % len(out))
