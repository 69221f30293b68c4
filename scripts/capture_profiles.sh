#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun, one GPU): usage capture_profiles.sh <tag>
# Each command runs plain first (must exit 0), then under ncu; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r2a}
B="python bench.py --steps 1 --warmup 3 --ticks-per-step 40 --fused-chunk 40 --no-cpu-baseline --no-e2e --no-stagger --no-config4"
mkdir -p gpurun_out
$B > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain bench failed"; exit 1; }
# launch list of the default (fused) loop: 4 fused launches of 40 ticks, then the 64 instrumented separate ticks, then the aux legs
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
# full set: the fused tick kernel (2 warm launches of 40 ticks) ...
ncu --set full --clock-control none --import-source on -k 'regex:tower_kernel' -s 2 -c 2 -f -o gpurun_out/prof_fused_$TAG $B > gpurun_out/ncu_fused_$TAG.log 2>&1
# ... and the two kernels of the separate-launch loop (3 tower + 3 advance launches)
B2="$B --no-fused"
$B2 > gpurun_out/plain_nofused_$TAG.json 2> gpurun_out/plain_nofused_$TAG.err || { echo "plain --no-fused bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k 'regex:tower_kernel|advance_kernel' -s 250 -c 6 -f -o gpurun_out/prof_$TAG $B2 > gpurun_out/ncu_full_$TAG.log 2>&1
python scripts/dbg_env_ncu.py > gpurun_out/env_plain_$TAG.log 2>&1 || { echo "plain env run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:env_step -s 2 -c 2 -f -o gpurun_out/prof_env_$TAG python scripts/dbg_env_ncu.py > gpurun_out/ncu_env_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_fused_$TAG.log; tail -n 2 gpurun_out/ncu_full_$TAG.log; tail -n 2 gpurun_out/ncu_env_$TAG.log; cat gpurun_out/env_plain_$TAG.log
# the SGD step on the device (csrc/spx_train.cu): launch list of 2 steps (after 3 warm steps), then ncu --set full of the two tcgen05 GEMM kernels
T="python scripts/dbg_train_time.py 20 128 2"
$T > gpurun_out/train_plain_$TAG.log 2>&1 || { echo "plain train run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches_$TAG.csv $T > gpurun_out/ncu_train_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:conv_tf32_kernel|wgrad_kernel' -s 600 -c 6 -f -o gpurun_out/prof_train_$TAG $T > gpurun_out/ncu_train_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_train_$TAG.log; cat gpurun_out/train_plain_$TAG.log
