#!/bin/bash
# Final check of the session: all GPU tests, smoke, both bench arms, jacknife sweep with the library's binomial stream.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/k_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/k_smoke.log
timeout 400 python bench.py > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench exit: $?" >> gpurun_out/k_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/k_bench_ref.json 2>&1; echo "ref exit: $?" >> gpurun_out/k_bench_ref.json
timeout 600 python scripts/jacknife_bench.py > gpurun_out/k_jacknife.log 2>&1
tail -6 gpurun_out/k_pytest.log; cat gpurun_out/k_smoke.log | tail -2; cut -c1-300 gpurun_out/k_bench.json; tail -1 gpurun_out/k_bench.err; cut -c1-300 gpurun_out/k_bench_ref.json; tail -1 gpurun_out/k_jacknife.log
