#!/bin/bash
# Round 2, call A: old suite + the BASELINE-shape parity tests (no -x: collect every deviation), the bench at the
# driver's flags, and the L2 prefetch probe.
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/r2a_pytest.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit: $?" >> gpurun_out/r2a_bench.err
timeout 300 python scripts/l2_prefetch_probe.py > gpurun_out/r2a_prefetch.jsonl 2> gpurun_out/r2a_prefetch.err
PROBE_CTAS=132 timeout 300 python scripts/l2_prefetch_probe.py > gpurun_out/r2a_prefetch132.jsonl 2>> gpurun_out/r2a_prefetch.err
tail -15 gpurun_out/r2a_pytest.log; cut -c1-600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err; head -3 gpurun_out/r2a_prefetch.jsonl
