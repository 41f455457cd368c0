"""Selects the metrics that profiles/*_ncu_raw.csv keep from `ncu -i X.ncu-rep --page raw --csv`
exports (one column per captured kernel).

    python tools/ncu_select.py out.csv name1=raw1.csv[:k] [name2=raw2.csv ...]      (k: k-th captured kernel, default 0)
"""
import csv
import re
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
    r"launch__(grid_size|block_size|registers_per_thread|occupancy_limit_\w+|shared_mem_per_block_dynamic)|"
    r"sm__inst_executed_pipe_(fp64|alu|fma|lsu|xu|tensor_subpipe_dmma)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__cycles_elapsed\.max|smsp__cycles_active\.avg|smsp__inst_executed\.sum|"
    r"smsp__thread_inst_executed_per_inst_executed\.ratio|"
    r"smsp__issue_active\.avg\.(pct_of_peak_sustained_active|per_cycle_active)|"
    r"smsp__warps_(active|eligible)\.avg\.per_cycle_active|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
    r"smsp__sass_thread_inst_executed_op_(dfma|dmul|dadd)_pred_on\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_ld|_op_st)?\.sum\.pct_of_peak_sustained_elapsed|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|lts__t_sector_hit_rate\.pct)$")


def load(path):
    k = 0
    if ":" in path and path.rsplit(":", 1)[1].isdigit():
        path, k = path.rsplit(":", 1)[0], int(path.rsplit(":", 1)[1])
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2 + k]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}, (vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")


# ncu scales every value of a capture on its own (Kbyte here, Mbyte there; us / ms; Gbyte/s / Tbyte/s): one
# unit per row, values of the other captures converted
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
         "s": 1.0, "second": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}


def convert(item, unit):
    u, v = item
    if u == unit or not v:
        return v
    num, den = (u.split("/") + [""])[:2], None
    tnum = (unit.split("/") + [""])[:2]
    try:
        f = SCALE[num[0]] / SCALE[tnum[0]]
        if num[1] != tnum[1]:
            f *= SCALE[tnum[1]] / SCALE[num[1]]
        return "%.6f" % (float(v.replace(",", "")) * f)
    except (KeyError, ValueError):
        return v + " " + u          # unknown unit pair: keep the value with its own unit


def main():
    out = sys.argv[1]
    cols = [a.split("=", 1) for a in sys.argv[2:]]
    data = [(n,) + load(p) for n, p in cols]
    names = sorted({h for _, d, _ in data for h in d if KEEP.match(h)})
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [n for n, _, _ in data])
        w.writerow(["kernel", ""] + [k for _, _, k in data])
        for h in names:
            unit = next((d[h][0] for _, d, _ in data if h in d), "")
            w.writerow([h, unit] + [convert(d.get(h, ("", "")), unit) for _, d, _ in data])


if __name__ == "__main__":
    main()
