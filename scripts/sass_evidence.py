#!/usr/bin/env python
"""Counts the Blackwell-native SASS mnemonics per kernel of libspx.so (cuobjdump -sass; runs here, no GPU): usage sass_evidence.py <tag>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2a"
lib = os.path.join(ROOT, "self_play_reinforcement_learning_b200", "libspx.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "SHFL", "DFMA|DMUL|DADD", "LDG.256|STG.256", "LDL|STL"]
cnt, name = collections.defaultdict(collections.Counter), None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", ln)
    if not m or name is None:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    c = cnt[name]
    if base == "UTCHMMA":
        c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
    elif base in ("UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "SHFL"):
        c[base] += 1
    elif base in ("DFMA", "DMUL", "DADD"):
        c["DFMA|DMUL|DADD"] += 1
    elif base in ("LDG", "STG") and ".256" in op:
        c["LDG.256|STG.256"] += 1
    elif base in ("LDL", "STL"):
        c["LDL|STL"] += 1
want = ["advance_kernel", "tower_kernel", "heads_kernel", "tttnet_kernel", "env_step_quad_kernel", "env_step_kernel", "conv_tf32_kernel", "wgrad_kernel",
        "bn_fwd_kernel", "bn_bwd"]
lines = [f"# {tag} SASS evidence (cuobjdump -sass libspx.so): Blackwell-native instructions per kernel",
         "# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, HMMA = mma.sync, "
         "LDGSTS = cp.async, LDG.256 = 256-bit global access, LDL|STL = local-memory (spill) traffic",
         "kernel," + ",".join(cols)]
for k in sorted(cnt):
    if any(w in k for w in want):
        lines.append(k.replace(",", ";") + "," + ",".join(str(cnt[k][c]) for c in cols))
open(os.path.join(ROOT, "profiles", f"{tag}_sass_evidence.csv"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
