#!/usr/bin/env python
"""Aggregate `ncu --page source --csv` (SASS view): total stall-reason samples and the hottest instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = {hdr[i]: 0 for i in st}
isamp = hdr.index('# Samples'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
nxt = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] in ('Address', 'Kernel Name')), len(rows))
body = [r for r in rows[hi + 1:nxt] if len(r) == len(hdr)]
for r in body:
    for i in st:
        tot[hdr[i]] += int(r[i] or 0)
S = sum(tot.values())
print("total samples", S, " instructions executed (warp)", sum(int(r[iex] or 0) for r in body))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-28s %8d %5.1f%%" % (k, v, 100.0 * v / max(S, 1)))
print("hottest SASS:")
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = max(st, key=lambda i: int(r[i] or 0))
    print("  %6s  %-70s %s" % (r[isamp], r[isrc].strip()[:70], hdr[top]))
