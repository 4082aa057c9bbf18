"""Keep a fixed set of columns of an `ncu --page raw --csv` export (the full export is ~2 MB per
capture).  usage: python profiles/pick_metrics.py raw.csv metric1,metric2,..."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = ["Kernel Name"] + sys.argv[2].split(",")
hdr = rows[0]
idx = [hdr.index(w) for w in want if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([hdr[i] for i in idx])
for r in rows[1:]:
    out.writerow([r[i] for i in idx])
