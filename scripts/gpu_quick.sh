#!/bin/bash
# Quick GPU check: model parity tests, a short stress of both backward variants, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/q_pytest.log
STRESS_STAGES=24 timeout 120 python scripts/stress.py cfg2 20 > gpurun_out/q_stress.log 2>&1
timeout 300 python bench.py --steps 260 --warmup 26 --cpu-steps 2 > gpurun_out/q_bench.log 2>&1
tail -15 gpurun_out/q_pytest.log; tail -1 gpurun_out/q_stress.log
python - <<'PY'
import json
for l in open("gpurun_out/q_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["stage_ms"], "group", d["replicate_group"] and d["replicate_group"]["value"])
PY
