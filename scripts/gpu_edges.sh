#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_edges.py tests/test_gpu_cli.py -m gpu -q -k "edges or reference_code or shapes or degenerate or weights or max_snps or matrix or prediction or missing or batch_size" 2>&1 | tail -30 > gpurun_out/edges.log
cat gpurun_out/edges.log
