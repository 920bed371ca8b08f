"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (developer tool).
    python scripts/ncu_lines.py src.csv [top_n] [kernel-name substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else None
cur = None; fn = None; out = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': continue
    if want and (fn is None or want not in fn): continue
    if r[0] != '':
        try:
            key = (cur, int(r[0]))
            o = out.setdefault(key, [r[1].strip()[:110], 0, 0, 0])
            o[1] += int(r[4]); o[2] += int(r[7]); o[3] += int(r[8]) if r[8].isdigit() else 0
        except Exception: pass
ts = sum(o[1] for o in out.values()); ti = sum(o[2] for o in out.values())
print('total samples', ts, 'warp inst', ti, 'avg threads/inst %.1f' % (sum(o[3] for o in out.values()) / max(ti, 1)))
for (f, ln), o in sorted(out.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-14s %4d smp %5.1f%% inst %5.1f%% thr %4.1f | %s' % (f[:14], ln, 100 * o[1] / ts, 100 * o[2] / ti, o[3] / max(o[2], 1), o[0]))
