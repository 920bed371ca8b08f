"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export over line ranges of one file (developer tool).
    python scripts/ncu_regions.py src.csv file.cuh name:lo-hi [name:lo-hi ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
regs = []
for a in sys.argv[3:]:
    n, rng = a.split(':'); lo, hi = rng.split('-'); regs.append((n, int(lo), int(hi)))
cur = None; agg = {n: [0, 0, 0] for n, _, _ in regs}; other = [0, 0, 0]
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No') or r[0] == '': continue
    try: ln, smp, inst, thr = int(r[0]), int(r[4]), int(r[7]), int(r[8]) if r[8].isdigit() else 0
    except Exception: continue
    tgt = other
    if cur == fname:
        for n, lo, hi in regs:
            if lo <= ln <= hi: tgt = agg[n]; break
    tgt[0] += smp; tgt[1] += inst; tgt[2] += thr
ts = sum(v[0] for v in agg.values()) + other[0]; ti = sum(v[1] for v in agg.values()) + other[1]
for n, v in list(agg.items()) + [('other files/lines', other)]:
    print('%-22s samples %5.1f%%  warp-inst %5.1f%%  thr/inst %4.1f' % (n, 100 * v[0] / ts, 100 * v[1] / ti, v[2] / max(v[1], 1)))
