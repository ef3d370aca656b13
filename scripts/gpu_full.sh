#!/bin/bash
# Full round check on the GPU box: all GPU tests, smoke, bench (both arms), default run on the fixture.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "ref exit: $?" >> gpurun_out/bench_ref.log
mkdir -p /tmp/fx && (time timeout 900 python -m locator_b200 --vcf tests/golden/data/test_genotypes.vcf.gz --sample_data tests/golden/data/test_sample_data.txt --out /tmp/fx/run --seed 12345 --keras_verbose 0) > gpurun_out/fixture_default.log 2>&1
wc -l /tmp/fx/run_history.txt >> gpurun_out/fixture_default.log
cat gpurun_out/pytest_gpu.log | tail -4; cat gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-600; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_ref.log; tail -12 gpurun_out/fixture_default.log
