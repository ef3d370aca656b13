#!/bin/bash
# A/B on the GPU box: bench with and without an environment switch ($1), plus the model parity tests.
mkdir -p gpurun_out
V=${1:-LOC_NO_ALTERNATE}
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_pytest.log
timeout 300 python bench.py --steps 260 --warmup 26 --cpu-steps 2 > gpurun_out/ab_on.log 2>&1
env $V=1 timeout 300 python bench.py --steps 260 --warmup 26 --cpu-steps 2 > gpurun_out/ab_off.log 2>&1
timeout 300 python bench.py --steps 260 --warmup 26 --cpu-steps 2 > gpurun_out/ab_on2.log 2>&1
tail -3 gpurun_out/ab_pytest.log
for f in ab_on ab_off ab_on2; do python - <<PY
import json
for l in open("gpurun_out/$f.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"], d.get("stages_us"))
PY
done
