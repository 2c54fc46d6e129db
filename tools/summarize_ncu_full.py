"""Key metrics of an `ncu --set full` report as a markdown table.  Usage: summarize_ncu_full.py report.ncu-rep > out.md"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
idx = [(hdr.index(k), label) for k, label in cols if k in hdr]
print("| " + " | ".join(l for _, l in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, label in idx:
        v = r[i]
        if label == "kernel":
            v = "`" + v.replace("void ", "").replace("unnamed>::", "")[:48] + "`"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f} {units[i]}" if units[i] not in ("", "register/thread") else v.split(".")[0]
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
