"""Executed-instruction mix by SASS opcode for one kernel of an ncu report (needs --import-source on / --set full).
    python tools/ncu_opmix.py report.ncu-rep kernel_regex [topn]
"""
import csv, sys, collections, subprocess
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-count", "1",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
ops = collections.Counter()
for r in rows:
    if r and r[0] in ("Address", "#"):
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        n = int(r[hdr["Instructions Executed"]])
    except (ValueError, KeyError):
        continue
    src = r[hdr["Source"]].strip()
    tok = src.split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "DMMA", "BAR")) and "." in op else "")] += n
tot = sum(ops.values())
print("total warp instructions", tot)
for op, n in ops.most_common(topn):
    print(f"{op:14s} {n:12d} {n / tot:6.3f}")
