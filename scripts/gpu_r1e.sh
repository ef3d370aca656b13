#!/bin/bash
# Ingest kernels (integer-logic scan), jacknife sweep at config-5 shape, default bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/e_pytest.log
timeout 300 python scripts/ingest_bench.py > gpurun_out/e_ingest_bench.json 2> gpurun_out/e_ingest_bench.err
timeout 900 python scripts/jacknife_bench.py > gpurun_out/e_jacknife.log 2>&1
timeout 600 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -4 gpurun_out/e_pytest.log; cat gpurun_out/e_ingest_bench.json | tr -d '\n '; echo; tail -3 gpurun_out/e_jacknife.log; cut -c1-700 gpurun_out/e_bench.json; tail -2 gpurun_out/e_bench.err
