#!/bin/bash
# L2 policy sweep for the first-layer backward (LOC_RESIDENT_MB x LOC_STREAM_HINT x alternation).
mkdir -p gpurun_out
: > gpurun_out/sweep_l2.txt
run() {
  env "$@" timeout 200 python bench.py --steps 260 --warmup 26 --no-cpu-baseline --group 0 > gpurun_out/sw.log 2>&1
  python - "$*" <<'PY' >> gpurun_out/sweep_l2.txt
import json, sys
for l in open("gpurun_out/sw.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], "| us/step", round(d["ms_per_step"] * 1e3, 2), "| value", round(d["value"]), "| b1f us", round(d["roofline"]["stage_ms"]["l1_backward_with_fused_next_forward"] * 1e3, 1))
        break
else:
    print(sys.argv[1], "FAILED")
    import shutil, time
    shutil.copy("gpurun_out/sw.log", "gpurun_out/sw_fail_%d.log" % (time.time() * 10 % 100000))
PY
}
for i in 1 2 3 4 5 6; do run LOC_STREAM_HINT=3; done
cat gpurun_out/sweep_l2.txt
