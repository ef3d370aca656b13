#!/bin/bash
# Round 2, call G: hidden stack with worker-warp layer chain + split cluster barrier: full GPU suite, clock trace,
# bench at the driver's flags, then racecheck / synccheck.
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
LOC_HID_TRACE=1 timeout 300 python scripts/hid_trace.py > gpurun_out/r2g_hid_trace.txt 2>&1
cat gpurun_out/r2g_hid_trace.txt | cut -c1-1200
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/r2g_pytest.log 2>&1
tail -8 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-queue > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
cut -c1-300 gpurun_out/r2g_bench.json; tail -2 gpurun_out/r2g_bench.err
bash scripts/gpu_sanitize.sh 2>&1 | tail -60
