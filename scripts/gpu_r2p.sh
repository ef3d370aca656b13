#!/bin/bash
# Round 2, call P: small-layer update under the hidden stack (chain order H -> U -> B) vs behind the backward (hbu).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_edges.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python scripts/timeline.py solo > gpurun_out/r2p_timeline_solo.txt 2> gpurun_out/r2p_timeline_solo.err
head -1 gpurun_out/r2p_timeline_solo.txt; tail -2 gpurun_out/r2p_timeline_solo.txt
LOC_CHAIN_ORDER=hbu timeout 300 python scripts/timeline.py solo > gpurun_out/r2p_timeline_solo_hbu.txt 2>&1; head -1 gpurun_out/r2p_timeline_solo_hbu.txt
for i in 1 2; do
timeout 600 python bench.py --steps 520 --warmup 52 --no-queue --no-cpu-baseline --group 0 > gpurun_out/r2p_bench_$i.json 2> gpurun_out/r2p_bench.err
cut -c1-230 gpurun_out/r2p_bench_$i.json
LOC_CHAIN_ORDER=hbu timeout 600 python bench.py --steps 520 --warmup 52 --no-queue --no-cpu-baseline --group 0 > gpurun_out/r2p_bench_hbu_$i.json 2>> gpurun_out/r2p_bench.err
cut -c1-230 gpurun_out/r2p_bench_hbu_$i.json
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-queue --no-cpu-baseline --group 0 > gpurun_out/r2p_bench_20.json 2>> gpurun_out/r2p_bench.err
cut -c1-230 gpurun_out/r2p_bench_20.json
