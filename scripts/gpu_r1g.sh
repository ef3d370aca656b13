#!/bin/bash
# Ring-of-pairs schedule: full GPU suite, default bench (group leg: ring vs lockstep), windows CLI on one GPU.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/g_pytest.log 2>&1
timeout 400 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
timeout 600 python scripts/windows_bench.py --gpus 1 --windows 24 > gpurun_out/g_windows_n1.log 2>&1
tail -6 gpurun_out/g_pytest.log
python - <<'PY'
import json
for l in open("gpurun_out/g_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "group", json.dumps(d["replicate_group"])[:700])
PY
tail -2 gpurun_out/g_bench.err; tail -2 gpurun_out/g_windows_n1.log
