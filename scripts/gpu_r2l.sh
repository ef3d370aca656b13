#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/timeline.py solo > gpurun_out/r2l_timeline_solo.txt 2> gpurun_out/r2l_timeline_solo.err
head -1 gpurun_out/r2l_timeline_solo.txt; tail -2 gpurun_out/r2l_timeline_solo.txt
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-queue --no-cpu-baseline --group 0 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
cut -c1-260 gpurun_out/r2l_bench.json
