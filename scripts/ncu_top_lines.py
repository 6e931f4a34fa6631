"""Aggregate ncu warp-stall samples of one kernel by CUDA source line.
The CSV of `ncu --page source` has per-SASS-instruction samples but no line numbers, so the SASS rows are aligned by
index with `nvdisasm --print-line-info` of the same function from the in-tree .so (build with -lineinfo)."""
import csv, glob, os, re, subprocess, sys, tempfile
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ground-fusion2_b200", "libgf2_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; sass = []
for r in rows:
    if r and r[0] == "Address": 
        if hdr is not None: break   # only the first launch
        hdr = r; continue
    if hdr and len(r) == len(hdr): sass.append(r)
col = {n: i for i, n in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = None
dis_name = sys.argv[5] if len(sys.argv) > 5 else kernel  # (part of) the mangled name when the kernel is a template instance
lines = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    # split by function
    m = re.search(r"\n\.text\.[^\n]*" + re.escape(dis_name) + r"[^\n]*:\n(.*?)(?=\n\s*\.section|\Z)", dis, re.S)
    if not m: continue
    cur = ("?", 0); lines = []
    for ln in m.group(1).splitlines():
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm: cur = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
    break
assert lines, "function not found in cubin"
if len(lines) != len(sass): print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} ncu rows", file=sys.stderr)
agg = {}
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for (f, l), r in zip(lines, sass):
    a = agg.setdefault((f, l), [0, 0, {}])
    a[0] += int(r[col["Warp Stall Sampling (All Samples)"]] or 0); a[1] += int(r[col["Instructions Executed"]] or 0)
    for n in reasons:
        v = int(r[col[n]] or 0)
        if v: a[2][n] = a[2].get(n, 0) + v
tot = sum(a[0] for a in agg.values()) or 1
print("kernel", kernel, "samples", tot, "warp-instructions", sum(a[1] for a in agg.values()))
src = {}
def text(f, l):
    if f not in src:
        p = glob.glob(os.path.join(os.path.dirname(so), "csrc", f)) + glob.glob(os.path.join(os.path.dirname(so), "..", "include", f))
        src[f] = open(p[0]).read().splitlines() if p else []
    return src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    rs = ",".join(f"{k[6:]}:{v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*a[0]/tot:5.1f}% inst {a[1]:>9} {f}:{l:<4} [{rs}] {text(f, l)}")
