#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q 2>&1 | tail -30 > gpurun_out/pytest_hidtc.log
LOC_HIDDEN_IMPL=simt timeout 900 python -m pytest tests/test_gpu_model.py -q 2>&1 | tail -5 > gpurun_out/pytest_hidsimt.log
timeout 600 python bench.py --steps 260 --warmup 26 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err
cat gpurun_out/pytest_hidtc.log; cat gpurun_out/pytest_hidsimt.log; cut -c1-1500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
