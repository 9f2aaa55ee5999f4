"""Condenses an .ncu-rep (ncu --set full) into the text summary kept under profiles/ (run where ncu is installed)."""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.DictReader(io.StringIO(det)))
keep_sections = ("GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis", "Launch Statistics",
                 "Occupancy", "GPU and Memory Workload Distribution", "Warp State Statistics", "Scheduler Statistics")
lines = []
if rows:
    r0 = rows[0]
    lines.append(f"kernel: {r0['Kernel Name']} | grid {r0['Grid Size']} block {r0['Block Size']}")
for r in rows:
    if r["Section Name"] in keep_sections and r["Metric Name"]:
        lines.append(f"{r['Section Name']:40s} | {r['Metric Name']:55s} | {r['Metric Unit']:14s} | {r['Metric Value']}")
rr = list(csv.reader(io.StringIO(raw)))
if len(rr) >= 3:
    hdr, units, vals = rr[0], rr[1], rr[2]
    want = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
            "launch__registers_per_thread", "smsp__cycles_active.avg", "sm__cycles_active.avg", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    lines.append("--- raw metrics")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) for w in want):
            lines.append(f"{h:80s} | {u:12s} | {v}")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(lines), "lines")
