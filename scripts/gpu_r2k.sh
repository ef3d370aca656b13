#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/timeline.py solo > gpurun_out/r2k_timeline_solo.txt 2> gpurun_out/r2k_timeline_solo.err
head -1 gpurun_out/r2k_timeline_solo.txt; tail -2 gpurun_out/r2k_timeline_solo.txt
LOC_NO_CHAIN=1 timeout 300 python scripts/timeline.py solo > gpurun_out/r2k_timeline_solo_nochain.txt 2>&1; head -1 gpurun_out/r2k_timeline_solo_nochain.txt;  tail -2 gpurun_out/r2k_timeline_solo_nochain.txt
