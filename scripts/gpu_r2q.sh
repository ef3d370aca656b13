#!/bin/bash
# Round 2, call Q: full suite + driver-flag bench + ring timeline at HEAD (update under the hidden stack).
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2q_pytest.log 2>&1
tail -6 gpurun_out/r2q_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
cut -c1-260 gpurun_out/r2q_bench.json; tail -2 gpurun_out/r2q_bench.err
timeout 300 python scripts/timeline.py ring > gpurun_out/r2q_timeline_ring.txt 2>&1; head -1 gpurun_out/r2q_timeline_ring.txt
python -c "import __graft_entry__ as g; g.smoke()"
