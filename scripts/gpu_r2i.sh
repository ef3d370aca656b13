#!/bin/bash
# Round 2, call I: kernel timelines of the solo / ring / lockstep schedules; quick parity of the branch-free elu; hid trace.
mkdir -p gpurun_out
for mode in solo ring lockstep; do
  timeout 300 python scripts/timeline.py $mode > gpurun_out/r2i_timeline_$mode.txt 2> gpurun_out/r2i_timeline_$mode.err
  head -1 gpurun_out/r2i_timeline_$mode.txt; tail -1 gpurun_out/r2i_timeline_$mode.txt
done
timeout 300 python scripts/timeline.py ring 2 > gpurun_out/r2i_timeline_ring2.txt 2>&1; head -1 gpurun_out/r2i_timeline_ring2.txt; tail -1 gpurun_out/r2i_timeline_ring2.txt
LOC_HID_TRACE=1 timeout 300 python scripts/hid_trace.py > gpurun_out/r2i_hid_trace.txt 2>&1
head -2 gpurun_out/r2i_hid_trace.txt | cut -c1-700
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ingest.py tests/test_gpu_edges.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-queue --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
cut -c1-260 gpurun_out/r2i_bench.json
