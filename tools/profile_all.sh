#!/bin/bash
# Per-kernel evidence for one proof (BLS12-381, 2^20 by default), from the repo root through gpurun:
#   gpurun --timeout 1200 -- 'bash tools/profile_all.sh r02n'
# Writes into gpurun_out/ (tools/ncu_summary.py <tag> then turns them into profiles/<tag>_*):
#   <tag>_launches.csv      every kernel launch of ONE proof (MSMs serialised) with its device time -> per-kernel shares
#   <tag>_full_raw.csv      ncu metrics of every kernel of one proof (raw page; sections: speed of light, memory, scheduler,
#                           warp states, occupancy, launch)
#   <tag>_ba_backward_kernel.ncu-rep + _raw.csv   `--set full --import-source on` of one launch of the dominant kernel
#   <tag>_meta.json         kernel_rev + library configuration of the capture
# Never read a bench number from these runs.
set -u
TAG=${1:-r02}; CURVE=${2:-bls12_381}; LOGN=${3:-20}
if [ $# -ge 3 ]; then shift 3; else shift $#; fi
O=gpurun_out; mkdir -p $O
export G16_PROFILE_TAG=$TAG
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${TAG}_launches.csv \
    python tools/profile_run.py $CURVE $LOGN "$@" > $O/${TAG}_launches.log 2>&1
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section Occupancy \
    --section LaunchStats --clock-control none --profile-from-start off -f -o $O/${TAG}_full \
    python tools/profile_run.py $CURVE $LOGN "$@" > $O/${TAG}_full.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>> $O/${TAG}_full.log
rm -f $O/${TAG}_full.ncu-rep
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ba_backward_kernel -c 1 -f \
    -o $O/${TAG}_ba_backward_kernel python tools/profile_run.py $CURVE $LOGN "$@" > $O/${TAG}_ba_backward_kernel.log 2>&1
ncu -i $O/${TAG}_ba_backward_kernel.ncu-rep --page raw --csv > $O/${TAG}_ba_backward_kernel_raw.csv 2>> $O/${TAG}_ba_backward_kernel.log
ls -la $O | grep $TAG
