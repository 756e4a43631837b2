#!/bin/bash
# One `ncu --set full` capture of a single kernel family on the GPU box, exported to CSV (the
# report itself is dropped: gpurun returns at most 64 MiB).
#   usage: bash profiles/capture_one.sh <name> <kernel regex> <skip> <count> "<drive_kernels args>" [tag]
set -u
NAME=$1; RE=$2; SKIP=$3; COUNT=$4; ARGS=$5; TAG=${6:-r02}
OUT=gpurun_out
mkdir -p $OUT
python profiles/drive_kernels.py $ARGS > $OUT/drive_${NAME}_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/drive_${NAME}_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $COUNT -o $OUT/prof_${NAME}_${TAG} \
    python profiles/drive_kernels.py $ARGS > $OUT/ncu_${NAME}_${TAG}.log 2>&1
ncu -i $OUT/prof_${NAME}_${TAG}.ncu-rep --page raw --csv > $OUT/raw_${NAME}_${TAG}.csv 2>/dev/null
ncu -i $OUT/prof_${NAME}_${TAG}.ncu-rep --page source --csv --print-source sass > $OUT/src_${NAME}_${TAG}.csv 2>/dev/null
gzip -f $OUT/src_${NAME}_${TAG}.csv
rm -f $OUT/*.ncu-rep
tail -6 $OUT/drive_${NAME}_${TAG}.log
