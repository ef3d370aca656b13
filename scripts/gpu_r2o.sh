#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2o_probe.jsonl
for rep in 1 2; do
for name in main inter; do
  lib=$PWD/locator_b200/lib/liblocator_b200.so; [ $name = inter ] && lib=$PWD/locator_b200/lib/variants/lib_inter.so
  LOC_LIB_PATH=$lib PROBE_FLAGS="2:0,4:0" PROBE_REPS=40 timeout 200 python scripts/fuse_probe.py 2>> gpurun_out/r2o_probe.err | sed "s/^{/{\"variant\": \"$name\", /" >> gpurun_out/r2o_probe.jsonl
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r2o_probe.jsonl"):
    d = json.loads(l); print(d["variant"], "stage", d["stage"], round(d["us_mean"], 1), round(d["us_min"], 1))
PY
