#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/ctas.txt
for c in 148 132 116 100; do
  LOC_TC_CTAS=$c timeout 200 python bench.py --steps 130 --warmup 26 --no-cpu-baseline --group 0 > gpurun_out/c.log 2>&1
  python - $c <<'PY' >> gpurun_out/ctas.txt
import json, sys
for l in open("gpurun_out/c.log"):
    if l.startswith("{"):
        d = json.loads(l); s = d["roofline"]["stage_ms"]
        print(sys.argv[1], "CTAs | us/step", round(d["ms_per_step"]*1e3,1), "| bwd", round(s["l1_backward"]*1e3,1), "| bwd+fwd", round(s["l1_backward_with_fused_next_forward"]*1e3,1), "| fwd", round(s["l1_forward"]*1e3,1))
        break
else:
    print(sys.argv[1], "FAILED")
PY
done
cat gpurun_out/ctas.txt
