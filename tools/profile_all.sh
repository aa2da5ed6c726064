#!/bin/bash
# Per-kernel evidence for one proof (BLS12-381, 2^20 by default), from the repo root through gpurun:
#   gpurun --timeout 900 -- 'bash tools/profile_all.sh r02'
# Writes into gpurun_out/:
#   <tag>_launches.csv      every kernel launch of ONE proof (MSMs serialised) with its device time -> per-kernel shares
#   <tag>_full_raw.csv      ncu --set full metrics of every kernel of one proof (raw page)
#   <tag>_<kernel>.ncu-rep  full report with source for the two dominant kernels (one launch each)
# Never read a bench number from these runs.
set -u
TAG=${1:-r02}; CURVE=${2:-bls12_381}; LOGN=${3:-20}; shift 3 2>/dev/null
O=gpurun_out; mkdir -p $O
export G16_PROFILE_TAG=$TAG
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${TAG}_launches.csv \
    python tools/profile_run.py $CURVE $LOGN "$@" > $O/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --profile-from-start off -f -o $O/${TAG}_full \
    python tools/profile_run.py $CURVE $LOGN "$@" > $O/${TAG}_full.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>> $O/${TAG}_full.log
rm -f $O/${TAG}_full.ncu-rep
for k in ba_backward_kernel; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -c 1 -f -o $O/${TAG}_$k \
      python tools/profile_run.py $CURVE $LOGN "$@" > $O/${TAG}_$k.log 2>&1
done
ls -la $O
