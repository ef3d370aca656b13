#!/bin/bash
# Round 2, call M: chained step on all SMs (skewed tile split), chained ring of two.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4
for mode in solo ring; do
  timeout 300 python scripts/timeline.py $mode > gpurun_out/r2m_timeline_$mode.txt 2> gpurun_out/r2m_timeline_$mode.err
  head -1 gpurun_out/r2m_timeline_$mode.txt; tail -2 gpurun_out/r2m_timeline_$mode.txt
done
LOC_NO_CHAIN=1 timeout 300 python scripts/timeline.py solo > gpurun_out/r2m_timeline_solo_nochain.txt 2>&1; head -1 gpurun_out/r2m_timeline_solo_nochain.txt
LOC_NO_CHAIN=1 timeout 300 python scripts/timeline.py ring > gpurun_out/r2m_timeline_ring_nochain.txt 2>&1; head -1 gpurun_out/r2m_timeline_ring_nochain.txt
for i in 1 2; do
timeout 600 python bench.py --steps 520 --warmup 52 --no-queue --no-cpu-baseline > gpurun_out/r2m_bench_$i.json 2> gpurun_out/r2m_bench.err
cut -c1-260 gpurun_out/r2m_bench_$i.json
LOC_NO_CHAIN=1 timeout 600 python bench.py --steps 520 --warmup 52 --no-queue --no-cpu-baseline --group 0 > gpurun_out/r2m_bench_nochain_$i.json 2>> gpurun_out/r2m_bench.err
cut -c1-260 gpurun_out/r2m_bench_nochain_$i.json
done
