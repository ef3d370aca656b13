#!/bin/bash
# Round 2, call N: full GPU suite at HEAD + bench at the driver's flags + launch list + ncu --set full of the step kernels.
mkdir -p gpurun_out
rm -f gpurun_out/parity_baseline_shapes.jsonl gpurun_out/parity_accuracy.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2n_pytest.log 2>&1
tail -6 gpurun_out/r2n_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
cut -c1-300 gpurun_out/r2n_bench.json; tail -2 gpurun_out/r2n_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2n_bench_ref.json 2>> gpurun_out/r2n_bench.err
cut -c1-300 gpurun_out/r2n_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2n_launches_bench.csv \
  python bench.py --steps 20 --warmup 5 --no-queue --group 0 --no-cpu-baseline --e2e-epochs 1 > gpurun_out/r2n_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_bwd_tc|k_hidden_tc|k_hidden_update" -s 21 -c 3 -f -o gpurun_out/r2n_step \
  python scripts/prof_step.py cfg2 26 > gpurun_out/r2n_prof_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_l1_fwd_wide|k_l1_fwd_tc" -c 2 -f -o gpurun_out/r2n_fwd \
  python scripts/prof_step.py cfg2 26 > gpurun_out/r2n_prof_fwd.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:"k_l1_bwd_tc" -s 30 -c 8 --csv --log-file gpurun_out/r2n_warm_l2.csv python scripts/prof_step.py cfg2 52 > gpurun_out/r2n_warm_l2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bb_l1_bwd" -s 1 -c 1 -f -o gpurun_out/r2n_bb64 \
  python scripts/bigbatch_profile.py 64 > gpurun_out/r2n_prof_bb.log 2>&1
tail -n 2 gpurun_out/r2n_prof_full.log gpurun_out/r2n_prof_fwd.log
ls -la gpurun_out | grep r2n
