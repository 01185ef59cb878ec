"""Opcode mix of a kernel from `ncu -i report --page source --csv` (SASS view): warp-level executed instructions and
stall samples per opcode.   python scripts/sass_mix.py source.csv[.gz] [top]"""
import collections
import csv
import gzip
import sys

p = sys.argv[1]
f = gzip.open(p, "rt", errors="replace") if p.endswith(".gz") else open(p, errors="replace")
rows = list(csv.reader(f))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
iS, iE, iT, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ex, th, sm, n = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= iSm:
        continue
    s = r[iS].strip()
    if s.startswith("@"):
        s = s.split(None, 1)[1] if " " in s else s
    op = s.split()[0].rstrip(";") if s else "?"
    base = op.split(".")[0]
    if base in ("LDS", "STS", "LDG", "STG", "ATOMG", "RED", "LD", "ST", "ATOM"):
        key = op
    else:
        key = base
    try:
        e, t, q = int(r[iE] or 0), int(r[iT] or 0), int(r[iSm] or 0)
    except ValueError:
        continue
    ex[key] += e; th[key] += t; sm[key] += q; n[key] += 1
tot, tots = sum(ex.values()), sum(sm.values())
print("%d SASS instructions, %d warp-level executed, %d thread-level, %d stall samples" % (sum(n.values()), tot, sum(th.values()), tots))
print("%-16s %6s %14s %6s %7s %8s" % ("opcode", "static", "warp-exec", "%", "lanes", "%samples"))
for k, v in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("%-16s %6d %14d %6.2f %7.1f %8.2f" % (k, n[k], v, 100.0 * v / tot, th[k] / max(v, 1), 100.0 * sm[k] / max(tots, 1)))
