#!/bin/bash
# Round 2, call F: 4 vs 2 forward warps in the fused backward; L2 tail-retention policies (LOC_TAIL=chunks,mode);
# parity tests on the 4-warp kernel.
mkdir -p gpurun_out
rm -f gpurun_out/r2f_probe.jsonl
V=locator_b200/lib/variants
probe() {  # name lib ctas tail
  LOC_LIB_PATH=$2 PROBE_CTAS=$3 LOC_TAIL=$4 PROBE_FLAGS="2:0,4:0" PROBE_REPS=30 timeout 200 python scripts/fuse_probe.py 2>> gpurun_out/r2f_probe.err \
    | sed "s/^{/{\"variant\": \"$1\", \"tail\": \"$4\", /" >> gpurun_out/r2f_probe.jsonl
}
MAIN=$PWD/locator_b200/lib/liblocator_b200.so
probe fw4 $MAIN "" ""
probe fw4 $MAIN 132 ""
probe fw2 $PWD/$V/lib_fw2.so "" ""
probe fw2 $PWD/$V/lib_fw2.so 132 ""
for ch in 12 24 36; do for mode in 0 1 3 7; do probe fw4 $MAIN "" "$ch,$mode"; done; done
probe fw4 $MAIN "" "48,3"
probe fw4 $MAIN "" "60,3"
probe fw4 $MAIN 132 "24,3"
python - <<'PY'
import json
for l in open("gpurun_out/r2f_probe.jsonl"):
    d = json.loads(l); print(d["variant"], d["l1_ctas"], "tail", d["tail"], "stage", d["stage"], round(d["us_mean"], 1), round(d["us_min"], 1))
PY
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "not divergence" 2>&1 | tail -5 > gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
