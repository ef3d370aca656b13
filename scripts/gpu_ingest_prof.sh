#!/bin/bash
# Full GPU test suite (timed) + ingest bench + ncu capture of the ingest kernels.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 ) > gpurun_out/d_pytest.log 2>&1
timeout 300 python scripts/ingest_bench.py > gpurun_out/d_ingest_bench.json 2> gpurun_out/d_ingest_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_site_stats|k_pack_sites' -c 4 -o gpurun_out/d_ingest python scripts/ingest_bench.py > gpurun_out/d_ncu.log 2>&1
ncu -i gpurun_out/d_ingest.ncu-rep --page raw --csv > gpurun_out/d_ingest_raw.csv 2>/dev/null
tail -14 gpurun_out/d_pytest.log; cat gpurun_out/d_ingest_bench.json | tr -d '\n '; echo; tail -2 gpurun_out/d_ncu.log
