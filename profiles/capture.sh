#!/bin/bash
# Runs on the GPU box (through gpurun): one plain run of profiles/drive_kernels.py, then
# `ncu --set full` captures of every kernel it launches, exported to CSV on the box so that
# what travels back in gpurun_out/ stays small.   usage: bash profiles/capture.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python profiles/drive_kernels.py > $OUT/drive_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/drive_${TAG}.log; exit 1; }
cap() {   # name, kernel regex, skip, count, extra ncu flags, drive args
    ncu --set full --clock-control none $5 -k "regex:$2" -s $3 -c $4 -o $OUT/prof_$1_${TAG} \
        python profiles/drive_kernels.py $6 > $OUT/ncu_$1_${TAG}.log 2>&1
    ncu -i $OUT/prof_$1_${TAG}.ncu-rep --page raw --csv > $OUT/raw_$1_${TAG}.csv 2>/dev/null
}
# the second lock-step pass of the step API (9 trace-generation + 9 warm-up launches of k_step skipped),
# with the SASS source page
cap step '^k_step$' 18 9 "--import-source on" "--only step"
ncu -i $OUT/prof_step_${TAG}.ncu-rep --page source --csv --print-source sass > $OUT/src_step_${TAG}.csv 2>/dev/null
gzip -f $OUT/src_step_${TAG}.csv
# the desynchronised batch: 9 trace + 4 mid-game + 31 mixing + 1 recording launches of k_step come first
cap desync '^k_step$' 45 2 "" "--only desync"
cap packed '^k_step_packed' 0 6 "" "--only packed"
cap io '^k_(observe|features|get_mask)' 0 8 "" "--only observe,features"
cap stepobs '^k_step_obs' 0 2 "" "--only stepobs"
cap qeval '^k_qeval' 0 4 "" "--only qeval"
cap play '^k_(rollout|sweep)' 0 6 "--import-source on" "--only rollout,sweep"
ncu -i $OUT/prof_play_${TAG}.ncu-rep --page source --csv --print-source sass > $OUT/src_play_${TAG}.csv 2>/dev/null
gzip -f $OUT/src_play_${TAG}.csv
cap mcts '^k_(mcts_run|env1)' 0 4 "" "--only mcts,env1"
# only the CSV exports travel back (gpurun returns at most 64 MiB): the reports themselves are dropped
rm -f $OUT/*.ncu-rep
tail -45 $OUT/drive_${TAG}.log
du -sh $OUT
