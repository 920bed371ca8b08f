"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (developer tool).
    python scripts/ncu_lines.py src.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0] != '':
        try: out.append((cur, int(r[0]), r[1].strip()[:100], int(r[4]), int(r[7]), int(r[8]) if r[8].isdigit() else 0))
        except Exception: pass
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print('total samples', ts, 'warp inst', ti, 'avg threads/inst %.1f' % (sum(o[5] for o in out) / max(ti, 1)))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print('%-14s %4d smp %5.1f%% inst %5.1f%% thr %4.1f | %s' % (o[0][:14], o[1], 100 * o[3] / ts, 100 * o[4] / ti, o[5] / max(o[4], 1), o[2]))
