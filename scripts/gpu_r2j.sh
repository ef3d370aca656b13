#!/bin/bash
# Round 2, call J: chained step (PDL + hand-over flags).  Targeted tests first (short timeouts: a wrong flag is a
# 2 s time-out per kernel, not a hang), then timelines, then the whole suite and the bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "chained or group_training or fit_history or survive" 2>&1 | tail -15 > gpurun_out/r2j_pytest_chain.log
tail -6 gpurun_out/r2j_pytest_chain.log
for mode in solo ring; do
  timeout 300 python scripts/timeline.py $mode > gpurun_out/r2j_timeline_$mode.txt 2> gpurun_out/r2j_timeline_$mode.err
  head -1 gpurun_out/r2j_timeline_$mode.txt; tail -1 gpurun_out/r2j_timeline_$mode.txt
done
LOC_NO_CHAIN=1 timeout 300 python scripts/timeline.py solo > gpurun_out/r2j_timeline_solo_nochain.txt 2>&1; head -1 gpurun_out/r2j_timeline_solo_nochain.txt
LOC_L1_ALL_SMS=1 timeout 300 python scripts/timeline.py solo > gpurun_out/r2j_timeline_solo_allsms.txt 2>&1; head -1 gpurun_out/r2j_timeline_solo_allsms.txt
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/r2j_pytest.log 2>&1
tail -6 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
cut -c1-260 gpurun_out/r2j_bench.json; tail -2 gpurun_out/r2j_bench.err
timeout 600 python bench.py --steps 520 --warmup 52 --no-queue --no-cpu-baseline --group 0 > gpurun_out/r2j_bench_long.json 2>> gpurun_out/r2j_bench.err
cut -c1-260 gpurun_out/r2j_bench_long.json
