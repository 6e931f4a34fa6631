"""Turns an `ncu --set full` report into the per-kernel summary bench.py reads (profiles/ncu_summary_r2.json):
    python scripts/ncu_summarize.py gpurun_out/ncu_solver_r2.ncu-rep 4096 "<command the capture ran>"
Per kernel (first captured launch): duration, DRAM bytes read + written (per launch and per window), fp64 tensor sub-pipe (DMMA) and fp64 FMA
pipe activity, warps active, registers, shared memory."""
import csv, datetime, io, json, subprocess, sys

rep, windows, command = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
num = lambda d, k: float(d[k].replace(",", "")) if d.get(k) not in (None, "", "n/a") else None
out = {"captured": datetime.date.today().isoformat(), "command": command, "windows_per_launch": windows, "report": rep.split("/")[-1], "kernels": {}}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0]
    if name in out["kernels"]:
        continue
    rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
    unit_r = dict(zip(hdr, rows[1])).get("dram__bytes_read.sum", "byte")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_r, 1.0)
    tot = (rd + wr) * scale if rd is not None and wr is not None else None
    dur = num(d, "gpu__time_duration.sum")
    dur_unit = dict(zip(hdr, rows[1])).get("gpu__time_duration.sum", "ns")
    dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(dur_unit, 1e-6) if dur is not None else None
    out["kernels"][name] = {
        "duration_ms_under_ncu": dur_ms, "dram_bytes_per_launch": tot, "dram_bytes_per_window": tot / windows if tot else None,
        "dmma_pipe_pct": num(d, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active") or num(d, "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": num(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or num(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": num(d, "launch__registers_per_thread"), "shared_mem_per_block": num(d, "launch__shared_mem_per_block_dynamic"),
        "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size"),
    }
json.dump(out, open("profiles/ncu_summary_r2.json", "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
