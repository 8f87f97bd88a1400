"""Condense `ncu -i X.ncu-rep --page raw --csv` into the metrics DESIGN.md quotes: python tools/ncu_summary.py raw.csv [raw2.csv ...] > out.json"""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "dram__bytes.sum.per_second": "dram_bytes_per_second",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in data:
        e = {"kernel": r[idx["Kernel Name"]].strip()}
        for k, name in KEEP.items():
            if k in idx and r[idx[k]] != "":
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                if name == "duration":
                    e["duration_us"] = round(v * UNIT.get(u, 1.0), 2)
                elif u in UNIT and "byte" in u:
                    e[name + "_MB"] = round(v * UNIT[u] / 1e6, 1)
                else:
                    e[name] = round(v, 2)
        launches.append(e)
    out[path.split("/")[-1]] = launches
print(json.dumps(out, indent=1))
