"""Summarise an .ncu-rep (ncu --set full) as a markdown table: one row per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "duration"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts %"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = "`" + v.replace("void ", "")[:44] + "`"
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
