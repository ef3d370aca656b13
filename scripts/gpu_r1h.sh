#!/bin/bash
# Two GPUs: replicate drivers through the real CLI (windows with worker-side prefetch, bootstrap with the FULL model queued).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/h_gpus.txt
timeout 600 python scripts/windows_bench.py --gpus 2 --windows 24 > gpurun_out/h_windows_n2.log 2>&1
timeout 600 python scripts/bootstrap_bench.py --gpus 2 --nboots 16 > gpurun_out/h_bootstrap_n2.log 2>&1
timeout 300 python -m pytest tests/test_gpu_tp.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/h_tp.log
cat gpurun_out/h_gpus.txt; tail -4 gpurun_out/h_windows_n2.log; tail -4 gpurun_out/h_bootstrap_n2.log; cat gpurun_out/h_tp.log
