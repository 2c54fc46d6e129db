"""Per kernel family: launches, device time and DRAM bytes from an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.
Usage: summarize_traffic.py launches.csv <pairs per step> > profiles/rNN_dram_traffic.json"""
import csv
import json
import sys
from collections import defaultdict

FAMILIES = ("k_gemm_tc", "k_gemm_tn", "k_chain", "k_res2net_front", "k_max_pool", "k_segnorm", "k_kpconv_gather_mma", "k_kpconv_c1", "k_grid_query_tq",
            "k_grid_query", "k_order", "k_row_positive", "k_pack_coarse", "k_overlap_pool")
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
per = defaultdict(lambda: defaultdict(float))
ids = defaultdict(set)
for r in csv.DictReader(lines):
    fam = next((f for f in FAMILIES if f + "<" in r["Kernel Name"] or f + "(" in r["Kernel Name"]), None)
    if fam is None:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        per[fam]["time_us"] += v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        ids[fam].add(r["ID"])
    elif r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        per[fam]["dram_read_MB" if "read" in r["Metric Name"] else "dram_write_MB"] += v * scale
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 f"bench.py --pairs {sys.argv[2]}, launches of one window (see the launch summary of the same round)",
       "pairs": int(sys.argv[2]),
       "families": {f: {"launches": len(ids[f]), "time_us": round(per[f]["time_us"], 1),
                        "dram_read_MB": round(per[f]["dram_read_MB"], 1), "dram_write_MB": round(per[f]["dram_write_MB"], 1)}
                    for f in FAMILIES if f in per}}
print(json.dumps(out, indent=1))
