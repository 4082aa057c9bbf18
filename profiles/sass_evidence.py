"""Counts of the Blackwell-specific SASS instructions per kernel of libbdpose.so:
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk),
SYNCS (mbarrier), LDGMC (multimem.ld_reduce).  usage: python profiles/sass_evidence.py > profiles/rN_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-modal-regression_b200", "bdpose", "libbdpose.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                       capture_output=True, text=True).stdout.splitlines()
short = {}
for raw, dem in zip(re.findall(r"Function : (\S+)", sass), names):
    dem = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", dem)
    cut = dem.rfind(">(")
    short[raw] = dem[:cut + 1] if cut > 0 else re.sub(r"\(.*", "", dem)
WANT = ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP", "SYNCS", "LDGMC", "UTCBAR", "ATOMS", "REDUX")
cnt = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = short[m.group(1)]
        cnt.setdefault(cur, collections.Counter())
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        cnt[cur]["total"] += 1
        for w in WANT:
            if op.startswith(w):
                cnt[cur][w] += 1
print("%-70s %6s " % ("kernel", "SASS") + " ".join("%7s" % w for w in WANT))
for k, c in cnt.items():
    if any(c[w] for w in WANT):
        print("%-70s %6d " % (k[:70], c["total"]) + " ".join("%7d" % c[w] for w in WANT))
