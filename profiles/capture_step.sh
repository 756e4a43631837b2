#!/bin/bash
# ncu --set full of the nine k_step launches of one lock-step pass (and the packed kernel), with
# the SASS source page exported on the box.   usage: bash profiles/capture_step.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python profiles/drive_kernels.py --only step,packed > $OUT/drive_step_${TAG}.log 2>&1 || { tail -20 $OUT/drive_step_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:^k_step -s 18 -c 10 \
    -o $OUT/prof_step_${TAG} python profiles/drive_kernels.py --only step,packed > $OUT/ncu_step_${TAG}.log 2>&1
ncu -i $OUT/prof_step_${TAG}.ncu-rep --page raw --csv > $OUT/raw_step_${TAG}.csv 2>/dev/null
ncu -i $OUT/prof_step_${TAG}.ncu-rep --page source --csv --print-source sass > $OUT/src_step_${TAG}.csv 2>/dev/null
gzip -f $OUT/src_step_${TAG}.csv
for f in $OUT/*.ncu-rep; do if [ $(stat -c %s "$f") -gt 25000000 ]; then rm -f "$f"; fi; done
tail -12 $OUT/drive_step_${TAG}.log
du -sh $OUT
