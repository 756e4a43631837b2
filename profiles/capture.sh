#!/bin/bash
# Runs on the GPU box (through gpurun): one plain run of profiles/drive_kernels.py, then
# `ncu --set full` captures of every kernel it launches, exported to CSV on the box so that
# what travels back in gpurun_out/ stays small (a .ncu-rep with sources of 40 launches does not).
#   usage: bash profiles/capture.sh <tag> [<drive_kernels args>]
set -u
TAG=${1:-r02}
shift || true
OUT=gpurun_out
mkdir -p $OUT
python profiles/drive_kernels.py "$@" > $OUT/drive_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/drive_${TAG}.log; exit 1; }
# the second pass of the step API (launches 9 trace + 9 warm-up pass, then 9 to read), with SASS
ncu --set full --clock-control none --import-source on -k regex:^k_step -s 18 -c 10 \
    -o $OUT/prof_step_${TAG} python profiles/drive_kernels.py "$@" > $OUT/ncu_step_${TAG}.log 2>&1
ncu -i $OUT/prof_step_${TAG}.ncu-rep --page raw --csv > $OUT/raw_step_${TAG}.csv 2>/dev/null
ncu -i $OUT/prof_step_${TAG}.ncu-rep --page source --csv --print-source sass > $OUT/src_step_${TAG}.csv 2>/dev/null
# everything else, metrics only
ncu --set full --clock-control none -k 'regex:^k_(observe|features|qeval|rollout|sweep|mcts_run)' -c 12 \
    -o $OUT/prof_rest_${TAG} python profiles/drive_kernels.py "$@" > $OUT/ncu_rest_${TAG}.log 2>&1
ncu -i $OUT/prof_rest_${TAG}.ncu-rep --page raw --csv > $OUT/raw_rest_${TAG}.csv 2>/dev/null
ls -la $OUT
for f in $OUT/*.ncu-rep; do
    if [ $(stat -c %s "$f") -gt 25000000 ]; then rm -f "$f"; fi
done
gzip -f $OUT/src_step_${TAG}.csv
du -sh $OUT
