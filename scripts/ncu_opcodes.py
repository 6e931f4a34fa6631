"""Executed warp-instructions and stall samples of one kernel aggregated by SASS opcode (from `ncu --page source`)."""
import csv, collections, re, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
nwin = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; col = {n: i for i, n in enumerate(h)}
ops = collections.defaultdict(lambda: [0, 0]); tot = 0; tots = 0
for r in rows[hi + 1:]:
    if len(r) != len(h): break
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]].strip())
    op = m.group(2) if m else r[col["Source"]]
    op = ".".join(op.split(".")[:2]) if op.startswith(("DMMA", "ATOMS", "LDS", "STS", "SHFL", "BAR", "LDG", "STG")) else op.split(".")[0]
    n = int(r[col["Instructions Executed"]] or 0); s = int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    ops[op][0] += n; ops[op][1] += s; tot += n; tots += s
for k, v in sorted(ops.items(), key=lambda x: -x[1][0])[:32]:
    print(f"{k:14s} {v[0] / nwin:12.0f} {100 * v[0] / tot:5.1f}% inst  {100 * v[1] / max(tots, 1):5.1f}% samples")
print("total", tot / nwin)
