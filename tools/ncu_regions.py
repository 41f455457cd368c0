"""Hot-spot table from `ncu -i X.ncu-rep --page source --csv`: stall samples and executed warp
instructions per window of N SASS instructions, with the dominant opcodes of the window (to label the
phase) and the average number of active threads.

    python tools/ncu_regions.py src.csv [window=250]
"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rows = list(csv.reader(open(path)))
hdr = rows[1]
ia, isrc, isamp, iex, ithr = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed",
                                                     "Thread Instructions Executed"))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= ithr or not r[isrc]:
        continue
    try:
        data.append((r[isrc], int(r[isamp] or 0), int(r[iex] or 0), int(r[ithr] or 0), [int(r[i] or 0) for i in stall_cols]))
    except ValueError:
        continue
tot_s = sum(d[1] for d in data) or 1
tot_i = sum(d[2] for d in data) or 1
print("instructions %d, samples %d, warp instructions executed %d" % (len(data), tot_s, tot_i))
print("%6s %7s %7s %6s  %-28s %s" % ("sass#", "samp%", "inst%", "thr", "top stalls", "top opcodes"))
for k in range(0, len(data), win):
    blk = data[k:k + win]
    s = sum(d[1] for d in blk)
    i = sum(d[2] for d in blk)
    t = sum(d[3] for d in blk)
    ops = Counter(d[0].split()[0].split(".")[0] if not d[0].startswith("@") else d[0].split()[1].split(".")[0] for d in blk)
    st = [sum(d[4][j] for d in blk) for j in range(len(stall_cols))]
    top = sorted(range(len(st)), key=lambda j: -st[j])[:3]
    print("%6d %6.1f%% %6.1f%% %6.1f  %-28s %s" % (
        k, 100.0 * s / tot_s, 100.0 * i / tot_i, t / i if i else 0,
        " ".join("%s:%d" % (hdr[stall_cols[j]][6:10], 100 * st[j] // max(sum(st), 1)) for j in top),
        " ".join("%s:%d" % kv for kv in ops.most_common(5))))
