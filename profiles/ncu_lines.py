"""Summarise an ncu report per source line: python profiles/ncu_lines.py <rep> [topN]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = None
agg = collections.defaultdict(lambda: [0, 0, ""])
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1].split("(")[0][-24:]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0].isdigit() and hdr:
        ie, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
        key = (cur_fn, cur_file, int(r[0]))
        try:
            agg[key][0] += int(r[ie])
            agg[key][1] += int(r[ism])
            agg[key][2] = r[1][:100]
        except ValueError:
            pass
for fn in sorted({k[0] for k in agg}):
    tot = sum(v[0] for k, v in agg.items() if k[0] == fn)
    ts = sum(v[1] for k, v in agg.items() if k[0] == fn)
    print("=====", fn, "warp-instructions", tot, "samples", ts)
    items = sorted(((v[0], k, v) for k, v in agg.items() if k[0] == fn), reverse=True)[:top]
    for c, k, v in items:
        print(f"  inst {100 * c / max(tot, 1):5.1f}%  smp {100 * v[1] / max(ts, 1):5.1f}%  {k[1]}:{k[2]:<4d} {v[2]}")
