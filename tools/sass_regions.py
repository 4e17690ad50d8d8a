"""Group an `ncu --page source --csv --print-source sass` export into contiguous regions of equal execution count.
usage: sass_regions.py sass.csv ITEMS   (ITEMS = work items of the launch; counts are printed per item)"""
import csv, itertools, sys
rows = list(csv.reader(open(sys.argv[1])))
items = float(sys.argv[2])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    ex = float(r[ix['Instructions Executed']] or 0)
    out.append((r[ix['Address']][-5:], ex / items, r[ix['Source']], int(float(r[ix['# Samples']] or 0))))
for k, g in itertools.groupby(out, key=lambda o: round(o[1], 1)):
    g = list(g)
    if k * len(g) < 20: continue
    print('%6.1f/item x %4d instr = %8.0f  [%s..%s] smp=%6d  %s' % (k, len(g), k * len(g), g[0][0], g[-1][0], sum(x[3] for x in g), g[0][2][:50]))
