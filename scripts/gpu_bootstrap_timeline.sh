#!/bin/bash
# Where does the fixed cost of a 2-GPU bootstrap run go?  (LOC_TIMING milestones of parent and workers)
mkdir -p gpurun_out
LOC_TIMING=1 timeout 600 python scripts/bootstrap_bench.py --gpus 2 --nboots 16 --epochs 100 > gpurun_out/j_bootstrap_n2.log 2>&1
cat gpurun_out/j_bootstrap_n2.log | tail -45
