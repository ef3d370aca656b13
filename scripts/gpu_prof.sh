#!/bin/bash
# ncu evidence (GPU box): launch list of one epoch, then --set full captures of the step kernels.
mkdir -p gpurun_out
W=${1:-cfg2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv python scripts/prof_step.py $W 26 > gpurun_out/prof_launch.log 2>&1
# in-epoch launch order: fwd, then (hidden, update, backward+next-forward) per step -> skip 7 matches, take one step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_bwd_tc|k_hidden" -s 7 -c 3 -f -o gpurun_out/prof_$W python scripts/prof_step.py $W 26 > gpurun_out/prof_full.log 2>&1
# the standalone forward and the unfused backward (epoch boundary / validation path)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_fwd_tc" -s 1 -c 1 -f -o gpurun_out/prof_${W}_fwd python scripts/prof_step.py $W 26 > gpurun_out/prof_full2.log 2>&1
tail -3 gpurun_out/prof_launch.log gpurun_out/prof_full.log gpurun_out/prof_full2.log
ls -la gpurun_out/
