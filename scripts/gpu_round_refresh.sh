#!/bin/bash
# One gpurun call that refreshes everything profiles/ holds for the default config: GPU tests, the full bench line, the launch
# list, the ncu --set full capture of one step and the sanitizer logs.  Usage: scripts/gpu_round_refresh.sh TAG
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log
timeout 900 python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_step.py > $O/${TAG}_launch_ncu.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"stage_kernel|splat_pano|replicate|image_order" -s 9 -c 9 -f -o $O/${TAG}_step python scripts/profile_step.py > $O/${TAG}_step_ncu.log 2>&1; echo "ncu full rc=$?"
timeout 900 compute-sanitizer --tool memcheck python scripts/profile_step.py 160 10 1 > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 $O/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python scripts/profile_step.py 160 10 1 > $O/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -1 $O/${TAG}_racecheck.log
