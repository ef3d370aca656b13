#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/last_pytest.log
cat gpurun_out/last_pytest.log
