#!/bin/bash
# Ring schedule of the replicate group (programmatic dependent launch): parity tests, then the bench's group leg.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "group or fused or survive" 2>&1 | tail -15 > gpurun_out/f_pytest.log
timeout 400 python bench.py --cpu-steps 2 > gpurun_out/f_bench_g4.json 2> gpurun_out/f_bench_g4.err
timeout 400 python bench.py --cpu-steps 2 --group 2 > gpurun_out/f_bench_g2.json 2> gpurun_out/f_bench_g2.err
tail -5 gpurun_out/f_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/f_bench_g4.json", "gpurun_out/f_bench_g2.json"):
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, "value", round(d["value"]), "group", json.dumps(d["replicate_group"])[:900])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/f_bench_g4.err
