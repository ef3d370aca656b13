#!/bin/bash
# Round 2, call H: hidden stack with 4-element finishes, full suite, trace, bench (incl. work queue), sanitizer triage.
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
LOC_HID_TRACE=1 timeout 300 python scripts/hid_trace.py > gpurun_out/r2h_hid_trace.txt 2>&1
head -3 gpurun_out/r2h_hid_trace.txt | cut -c1-900
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/r2h_pytest.log 2>&1
tail -6 gpurun_out/r2h_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
cut -c1-300 gpurun_out/r2h_bench.json; tail -2 gpurun_out/r2h_bench.err
bash scripts/gpu_sanitize.sh 2>&1 | tail -80
