#!/bin/bash
# compute-sanitizer racecheck + synccheck over the step kernels (verdict item 7).  Logs -> gpurun_out/r2_sanitize_*.log
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_race.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "SUMMARY|done|hazard|Error" gpurun_out/r2_sanitize_$tool.log | sort | uniq -c | head -20
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_race.py skip_hidden > gpurun_out/r2_sanitize_${tool}_nohidden.log 2>&1
  echo "== $tool (first layer only) rc=$?"; grep -E "SUMMARY|done|hazard|Error" gpurun_out/r2_sanitize_${tool}_nohidden.log | sort | uniq -c | head -20
done
python scripts/race_triage.py gpurun_out/r2_sanitize_racecheck.log gpurun_out/r2_sanitize_racecheck_nohidden.log > gpurun_out/r2_race_triage.md 2>&1
head -60 gpurun_out/r2_race_triage.md
