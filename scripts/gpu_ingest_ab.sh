#!/bin/bash
# Ingest kernels: site scan with the dp4a fast path vs the byte-test version (lib_stat4.so, same box), tests, ncu --set full at HEAD.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/ingest_bench.py > gpurun_out/r2_ingest_bench.json 2> gpurun_out/r2_ingest_bench.err
LOC_LIB_PATH=$PWD/locator_b200/lib/variants/lib_stat4.so timeout 300 python scripts/ingest_bench.py > gpurun_out/r2_ingest_bench_bytetests.json 2>> gpurun_out/r2_ingest_bench.err
python - <<'PY'
import json
for f in ("r2_ingest_bench.json", "r2_ingest_bench_bytetests.json"):
    d = json.load(open("gpurun_out/" + f)); print(f, {k: (round(v["ms"], 3), round(v["GBps"])) for k, v in d.items()})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_site_stats|k_pack_sites" -c 4 -f -o gpurun_out/r2_ingest python scripts/ingest_bench.py > gpurun_out/r2_ingest_ncu.log 2>&1
tail -2 gpurun_out/r2_ingest_ncu.log
