#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/h_pytest.log
timeout 100 python scripts/hid_trace.py 2>&1 | head -3 | cut -c1-900 > gpurun_out/h_trace.log
timeout 300 python bench.py --steps 260 --warmup 26 --cpu-steps 2 > gpurun_out/h_bench.log 2>&1
tail -15 gpurun_out/h_pytest.log; cat gpurun_out/h_trace.log
python - <<'PY'
import json
for l in open("gpurun_out/h_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "e2e", d["e2e"]["value"], d["roofline"]["stage_ms"], "group", d["replicate_group"] and d["replicate_group"]["value"])
PY
tail -3 gpurun_out/h_bench.log | cut -c1-300
