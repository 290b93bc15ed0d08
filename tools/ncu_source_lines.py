"""Per-source-line totals (stall samples, warp instructions) of one kernel from an ncu report captured with
--import-source on:  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv -k regex:NAME > t.csv;
python tools/ncu_source_lines.py t.csv [top]"""
import csv
import sys


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg, f = [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] == "Line No" or r[0] == "":
        continue
    agg.append([f, num(r[0]), r[1][:100], num(r[4]), num(r[7])])
tot_s = sum(a[3] for a in agg) or 1
tot_i = sum(a[4] for a in agg) or 1
print("total samples", tot_s, "total warp instructions", tot_i)
agg.sort(key=lambda a: -a[3])
for a in agg[:top]:
    print(f"{a[0]:16s} {a[1]:5d} samp {a[3]:6d} ({100 * a[3] / tot_s:4.1f}%) inst {a[4]:9d} ({100 * a[4] / tot_i:4.1f}%)  {a[2]}")
