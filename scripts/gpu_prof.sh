#!/bin/bash
mkdir -p gpurun_out
W=${1:-cfg2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv python scripts/prof_step.py $W 6 > gpurun_out/prof_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_bwd_tc|k_hidden|k_l1_fwd_tc" -s 8 -c 4 -f -o gpurun_out/prof_$W python scripts/prof_step.py $W 6 > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_launch.log gpurun_out/prof_full.log
ls -la gpurun_out/
