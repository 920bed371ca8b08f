#!/bin/bash
# One gpurun call: `ncu --set full` of the kernels matching REGEX in the second pass of scripts/profile_step.py.
# Usage: scripts/gpu_ncu.sh TAG REGEX SKIP COUNT
TAG=${1:-x}; RE=${2:-stage_kernel}; SKIP=${3:-4}; CNT=${4:-4}
O=gpurun_out
mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o $O/${TAG} python scripts/profile_step.py > $O/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/${TAG}_ncu.log; ls -la $O/${TAG}.ncu-rep
