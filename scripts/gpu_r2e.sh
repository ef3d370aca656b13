#!/bin/bash
# Round 2, call E: A/B of first-layer backward variants (store warp releases eagerly / fast gamma-beta Adam /
# forward warps on schedulers 2,3), steady-state DRAM traffic of the backward with a warm L2 (ncu --cache-control none),
# the work queue + e2e with the model handle pool.
mkdir -p gpurun_out
rm -f gpurun_out/r2e_probe.jsonl
V=locator_b200/lib/variants
for name in base eager eager_fast eager_fast_s23 s23; do
  lib=$PWD/$V/lib_$name.so; [ $name = base ] && lib=$PWD/locator_b200/lib/liblocator_b200.so
  for ctas in "" 132; do
    LOC_LIB_PATH=$lib PROBE_CTAS=$ctas PROBE_FLAGS="2:0,4:0,4:2,4:4" timeout 200 python scripts/fuse_probe.py 2>> gpurun_out/r2e_probe.err | sed "s/^{/{\"variant\": \"$name\", /" >> gpurun_out/r2e_probe.jsonl
  done
done
cat gpurun_out/r2e_probe.jsonl
for name in eager_fast eager_fast_s23; do
  LOC_LIB_PATH=$PWD/$V/lib_$name.so timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "not divergence" 2>&1 | tail -5 > gpurun_out/r2e_pytest_$name.log
  tail -3 gpurun_out/r2e_pytest_$name.log
done
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
  -k regex:"k_l1_bwd_tc" -s 30 -c 12 --csv --log-file gpurun_out/r2e_warm_l2.csv python scripts/prof_step.py cfg2 52 > gpurun_out/r2e_warm_l2.log 2>&1
tail -2 gpurun_out/r2e_warm_l2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
cut -c1-400 gpurun_out/r2e_bench.json; tail -2 gpurun_out/r2e_bench.err
