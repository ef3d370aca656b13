#!/bin/bash
mkdir -p gpurun_out
timeout 180 python scripts/tc_probe.py 4136 32 > gpurun_out/probe1.log 2>&1; echo "exit $?" >> gpurun_out/probe1.log
timeout 180 python scripts/tc_probe.py 100000 21 > gpurun_out/probe2.log 2>&1; echo "exit $?" >> gpurun_out/probe2.log
timeout 180 python scripts/tc_probe.py 5830 32 > gpurun_out/probe3.log 2>&1; echo "exit $?" >> gpurun_out/probe3.log
cat gpurun_out/probe1.log gpurun_out/probe2.log gpurun_out/probe3.log
