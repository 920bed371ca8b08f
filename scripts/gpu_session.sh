#!/bin/bash
# One gpurun call: GPU tests, a short bench, the ncu launch list of one bench step.  Usage: scripts/gpu_session.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -25 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $O/${TAG}_bench.err; cat $O/${TAG}_bench.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d[k] for k in ('value', 'ms_per_step', 'value_no_dedup', 'no_dedup_matches', 'gpu_launches') if k in d}, d.get('e2e', {}).get('value'), d.get('e2e', {}).get('matches_device_path'))
    print({k: (round(v['ms_per_step'], 3), round(v['frac'], 4)) for k, v in d['roofline']['kernels'].items()})
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_step.py > $O/${TAG}_ncu.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("$O/${TAG}_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ki][:50], []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print("%-52s n=%3d last=%10.1f us" % (k, len(v), v[-1] / 1e3))
PY
