"""Print the roofline-relevant metrics of every launch in an .ncu-rep (run here, no GPU needed):
    python scripts/ncu_raw_summary.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
        "smsp__cycles_active.avg", "sm__cycles_active.avg"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
kn = hdr.index("Kernel Name")
units = rows[1]
print("units:", {w: units[i] for w, i in idx})
for r in rows[2:]:
    print(r[kn][:70])
    print("   ", {w.split(".")[0]: r[i] for w, i in idx})
