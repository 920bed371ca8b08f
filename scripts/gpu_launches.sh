#!/bin/bash
# ncu launch list (per-kernel durations) of one bench step for each library given.  Usage: scripts/gpu_launches.sh TAG lib1.so [lib2.so ...]
TAG=${1:-x}; shift
O=gpurun_out
mkdir -p $O
for LIB in "$@"; do
  B=$(basename $LIB .so)
  SALVE_BEV_LIB=$PWD/$LIB timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_${B}_launches.csv python scripts/profile_step.py > $O/${TAG}_${B}_ncu.log 2>&1
  echo "== $B"
  python - <<PY
import csv
rows = [r for r in csv.reader(open("$O/${TAG}_${B}_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ki][:50], []).append(float(r[vi].replace(",", "")))
tot = 0
for k, v in agg.items():
    print("%-52s n=%3d last=%10.1f us" % (k, len(v), v[-1] / 1e3)); tot += v[-1] / 1e3
print("sum of last launches: %.1f us" % tot)
PY
done
