"""Turn an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into the few numbers the roofline uses.
usage: python profiles/summarize.py profiles/<name>_raw.csv"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_fma.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = f"{r[i]} {units[i]}".strip()
    if "dram__bytes_read.sum" in hdr:
        i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        d["dram_bytes_total"] = to_bytes(r[i], units[i]) + to_bytes(r[j], units[j])
    out.append(d)
print(json.dumps(out, indent=1))
