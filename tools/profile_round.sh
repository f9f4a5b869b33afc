#!/bin/bash
# Runs on the GPU box (under gpurun): the ncu evidence of a round. Usage: tools/profile_round.sh <out dir under gpurun_out>
# 1. launch lists (gpu__time_duration, cold-cache / serialised: compare shares) of the default workload and of --config 5
# 2. one `--set full` capture of K1a/K1b and the pose kernels for each workload
set -u
OUT=gpurun_out/$1
mkdir -p $OUT
B="python bench.py --no-e2e --no-cpu-baseline --no-extras --no-config5"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -k regex:'cape|pose' -c 40 --csv --log-file $OUT/launches.csv $B --steps 2 --warmup 3 > $OUT/launches_bench.log 2>&1
$NCU --metrics gpu__time_duration.sum -k regex:'cape|pose' -c 40 --csv --log-file $OUT/launches_c5.csv $B --config 5 --steps 2 --warmup 3 > $OUT/launches_c5_bench.log 2>&1
# skip the warm-up launches of each kernel (-s counts matching launches), take one
$NCU --set full --import-source on -k regex:'cape_cell_fit_kernel|cape_cell_finish_kernel' -s 6 -c 2 -o $OUT/k1 $B --steps 2 --warmup 3 > $OUT/k1.log 2>&1
$NCU --set full --import-source on -k regex:'pose_ransac_kernel|pose_variance_kernel|cape_segment_kernel' -s 9 -c 3 -o $OUT/pose $B --steps 2 --warmup 3 > $OUT/pose.log 2>&1
$NCU --set full --import-source on -k regex:'cape_cell_fit_kernel|cape_cell_finish_kernel' -s 6 -c 2 -o $OUT/k1_c5 $B --config 5 --steps 2 --warmup 3 > $OUT/k1_c5.log 2>&1
$NCU --set full --import-source on -k regex:'pose_fused_kernel|pose_ransac_kernel' -s 3 -c 2 -o $OUT/pose_c5 $B --config 5 --steps 2 --warmup 3 > $OUT/pose_c5.log 2>&1
ls -la $OUT
