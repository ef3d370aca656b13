#!/bin/bash
# Round 2, call B: all GPU tests (new: accuracy parity, wide inference, ingest compaction, TP vs oracle, tf32 oracle),
# host-overhead profile of e2e + work queue, wide-forward roofline.
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -120 ) > gpurun_out/r2b_pytest.log 2>&1
timeout 300 python scripts/wide_forward_bench.py > gpurun_out/r2b_wide.json 2> gpurun_out/r2b_wide.err
timeout 600 python scripts/host_overheads.py > gpurun_out/r2b_host.out 2> gpurun_out/r2b_host.err
tail -30 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_wide.json; tail -3 gpurun_out/r2b_wide.err
