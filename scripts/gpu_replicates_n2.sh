#!/bin/bash
# Two GPUs, heavier replicates (100 epochs each): windows and bootstrap through the real CLI, early worker start.
mkdir -p gpurun_out
timeout 700 python scripts/windows_bench.py --gpus 2 --windows 16 --epochs 100 > gpurun_out/i_windows_n2.log 2>&1
timeout 600 python scripts/bootstrap_bench.py --gpus 2 --nboots 16 --epochs 100 > gpurun_out/i_bootstrap_n2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -x -q -k "bootstrap or windows" 2>&1 | tail -3 > gpurun_out/i_cli.log
tail -3 gpurun_out/i_windows_n2.log; tail -3 gpurun_out/i_bootstrap_n2.log; cat gpurun_out/i_cli.log
