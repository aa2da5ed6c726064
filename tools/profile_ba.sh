#!/bin/bash
# Profile the batched-affine rounds (csrc/msm_ba.cuh) on the GPU box.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 900 -- 'bash tools/profile_ba.sh 3 0'        # rounds for G1 MSMs, rounds for the G2 MSM
# Writes into gpurun_out/: the launch list of one proof with serialised MSMs (per-kernel time shares) and the raw ncu
# metrics of one launch of every kernel of a round plus the final accumulation.  Never read a bench number from this run.
set -u
O=gpurun_out
mkdir -p $O
export G16_MSM_BA=${1:-3} G16_MSM_BA_G2=${2:-0}
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/ba_launches.csv \
    python tools/profile_run.py bls12_381 20 1 > $O/ba_launches.log 2>&1
for k in ba_forward_kernel ba_combine_kernel ba_backward_kernel msm_accum_l0; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/ba_$k \
      python tools/profile_run.py bls12_381 20 1 > $O/ba_$k.log 2>&1
  ncu -i $O/ba_$k.ncu-rep --page raw --csv > $O/ba_${k}_raw.csv 2>> $O/ba_$k.log
  rm -f $O/ba_$k.ncu-rep    # a full report with sources exceeds what gpurun brings back; the CSV is what gets read
done
ls -la $O
