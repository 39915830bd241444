"""Key metrics of every kernel in an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py file.ncu-rep [--csv out.csv]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
idx = [(w, h.index(w)) for w in WANT if w in h]
out = [[w for w, _ in idx] + ["(units)"]]
for r in data:
    out.append([r[i] for _, i in idx] + [" ".join(units[i] for _, i in idx)])
if "--csv" in sys.argv:
    with open(sys.argv[sys.argv.index("--csv") + 1], "w", newline="") as f:
        csv.writer(f).writerows([[w for w, _ in idx]] + [[units[i] for _, i in idx]] + [[r[i] for _, i in idx] for r in data])
for r in data:
    print("---", r[h.index("Kernel Name")][:80])
    for w, i in idx[1:]:
        print(f"  {w:75s} {r[i]:>18s} {units[i]}")
