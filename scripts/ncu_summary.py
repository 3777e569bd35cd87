"""Summarise an .ncu-rep (ncu --set full) as a markdown table: one row per captured launch.
usage: ncu_summary.py report.ncu-rep [traffic.json key regex]  -- the optional arguments record the mean DRAM bytes per
launch of the kernels matching `regex` under `key` in the JSON file bench.py reads (profiles/ncu_traffic.json)."""
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

if len(sys.argv) > 4:
    import json, os, re
    path, key, rx = sys.argv[2], sys.argv[3], re.compile(sys.argv[4])
    ir, iw, ik, ig = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("Grid Size")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
            for r in rows[2:] if rx.search(r[ik] + " " + r[ig])]
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
              "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}: dram__bytes_read.sum + dram__bytes_write.sum, mean over {len(vals)} launches"}
    json.dump(d, open(path, "w"), indent=1)
