#!/bin/bash
# Round 2, call D: the fixed tests, the fused-forward timing probe (where do its 9 us go), bench at the driver's flags.
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
timeout 300 python scripts/fuse_probe.py > gpurun_out/r2d_fuse_probe.jsonl 2> gpurun_out/r2d_fuse_probe.err
PROBE_CTAS=132 timeout 300 python scripts/fuse_probe.py > gpurun_out/r2d_fuse_probe132.jsonl 2>> gpurun_out/r2d_fuse_probe.err
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -120 ) > gpurun_out/r2d_pytest.log 2>&1
cat gpurun_out/r2d_fuse_probe.jsonl gpurun_out/r2d_fuse_probe132.jsonl; tail -30 gpurun_out/r2d_pytest.log
