#!/bin/bash
# Sweep the batched-affine knobs on one B200 (through gpurun, from the repo root):
#   gpurun --timeout 900 -- 'bash tools/sweep_ba.sh'
# One bench.py run per configuration (5 timed proofs, CPU baseline skipped except for the first and last, which also
# assert bit-identity with the CPU oracle); one line per configuration in gpurun_out/sweep_ba.txt.
set -u
O=gpurun_out
mkdir -p $O
: > $O/sweep_ba.txt
run() {   # name, extra bench args, env assignments...
  local name=$1 extra=$2; shift 2
  env "$@" timeout 170 python bench.py --steps 5 --warmup 3 $extra > $O/sweep_$name.json 2> $O/sweep_$name.err
  python - "$name" "$O/sweep_$name.json" >> $O/sweep_ba.txt <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms_per_step=%.2f" % d["ms_per_step"], "e2e_ms=%.2f" % d["e2e"]["ms_per_step"],
          "g1_span_ms=%.2f" % d["roofline"]["avg_launch_ms"], "g2_span_ms=%.2f" % d["kernels"]["msm_accum_l0_g2"]["launch_ms"],
          "launches=%d" % d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base         ""                  G16_MSM_BA=0
run r3_gcd       "--no-cpu-baseline" G16_MSM_BA=3 G16_BA_INV_GCD=1
run r3_gcd_lean  "--no-cpu-baseline" G16_MSM_BA=3 G16_BA_INV_GCD=1 G16_BA_LEAN=1
run r2_gcd_lean  "--no-cpu-baseline" G16_MSM_BA=2 G16_BA_INV_GCD=1 G16_BA_LEAN=1
run r4_gcd_lean  "--no-cpu-baseline" G16_MSM_BA=4 G16_BA_INV_GCD=1 G16_BA_LEAN=1
run r5_gcd_lean  "--no-cpu-baseline" G16_MSM_BA=5 G16_BA_INV_GCD=1 G16_BA_LEAN=1
run r3_m16       "--no-cpu-baseline" G16_MSM_BA=3 G16_BA_INV_GCD=1 G16_BA_LEAN=1 G16_BA_M=16
run r3_g16       "--no-cpu-baseline" G16_MSM_BA=3 G16_BA_INV_GCD=1 G16_BA_LEAN=1 G16_BA_G=16
run r3_g2r2      "--no-cpu-baseline" G16_MSM_BA=3 G16_MSM_BA_G2=2 G16_BA_INV_GCD=1 G16_BA_LEAN=1
run r4_final     ""                  G16_MSM_BA=4 G16_BA_INV_GCD=1 G16_BA_LEAN=1
cat $O/sweep_ba.txt
