"""Aggregate an ncu source-page export (--page source --csv --print-source cuda,sass) per CUDA source line.
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K | python profiles/ncu_lines.py [N]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file = ""
data = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("", "Function Name"):
        continue
    try:
        data.append((int(r[ii]), int(r[si]), cur_file, int(r[0]), r[1].strip()[:100]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print(f"total warp-instructions {tot}, samples {tots}")
for n, s, f, l, src in sorted(data, reverse=True)[:top]:
    print(f"{n:>11} {100*n/tot:5.1f}%  samp {100*s/max(tots,1):5.1f}%  {f}:{l}: {src}")
