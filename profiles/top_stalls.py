"""Top stall locations + opcode mix of one kernel from an `ncu --page source --csv` export.
usage: python profiles/top_stalls.py source.csv [n]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ia, isrc = hdr.index("Address"), hdr.index("Source")
iss, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
seen, data = set(), []
for r in rows[hi + 1:]:
    try:
        if r[ia] in seen:
            continue
        seen.add(r[ia])
        data.append((int(r[iss]), r[isrc].strip()[:110], int(r[iex])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
print("total stall samples %d, warp instructions %d" % (tot, sum(d[2] for d in data)))
for i, d in enumerate(data):
    data[i] = d + (i,)
for d in sorted(data, reverse=True)[:n]:
    print("%6d %5.1f%%  [%4d] %-100s ex=%d" % (d[0], 100.0 * d[0] / tot, d[3], d[1], d[2]))
mix = Counter()
for d in data:
    tok = d[1].split()
    op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
    mix[op.split(".")[0]] += d[2]
print("opcode mix:", mix.most_common(22))
