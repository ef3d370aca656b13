#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_edges.py -m gpu -q 2>&1 | tail -40 > gpurun_out/edges.log
cat gpurun_out/edges.log
