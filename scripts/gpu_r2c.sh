#!/bin/bash
# Round 2, call C: profiling at HEAD.  Hidden-stack clock trace, ncu launch list of the bench command (extras off),
# ncu --set full of the in-epoch step kernels (fused backward, hidden stack) and of the wide inference forward.
mkdir -p gpurun_out
LOC_HID_TRACE=1 timeout 300 python scripts/hid_trace.py > gpurun_out/r2c_hid_trace.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_bench.csv \
  python bench.py --steps 20 --warmup 5 --no-queue --group 0 --no-cpu-baseline --e2e-epochs 1 > gpurun_out/r2c_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_bwd_tc|k_hidden_tc" -s 14 -c 2 -f -o gpurun_out/r2c_step \
  python scripts/prof_step.py cfg2 26 > gpurun_out/r2c_prof_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_fwd_wide" -s 1 -c 1 -f -o gpurun_out/r2c_wide \
  python scripts/prof_step.py cfg2 26 > gpurun_out/r2c_prof_wide.log 2>&1
tail -3 gpurun_out/r2c_bench_under_ncu.log gpurun_out/r2c_prof_full.log gpurun_out/r2c_prof_wide.log
cat gpurun_out/r2c_hid_trace.txt
ls -la gpurun_out/ | grep r2c
