#!/bin/bash
# Round-1 (session 3) check: ingest kernels (vector loads), staged zarr upload, prefetching replicate drivers.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/c_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/c_pytest.log
timeout 300 python scripts/ingest_bench.py > gpurun_out/c_ingest_bench.json 2> gpurun_out/c_ingest_bench.err
timeout 600 python scripts/windows_bench.py --gpus 1 --windows 24 > gpurun_out/c_windows_n1.log 2>&1
LOC_WINDOWS_PARENT_INGEST=1 timeout 600 python scripts/windows_bench.py --gpus 1 --windows 24 --reuse > gpurun_out/c_windows_n1_serial.log 2>&1
tail -6 gpurun_out/c_pytest.log; cat gpurun_out/c_ingest_bench.json; tail -3 gpurun_out/c_ingest_bench.err; tail -3 gpurun_out/c_windows_n1.log; tail -3 gpurun_out/c_windows_n1_serial.log
