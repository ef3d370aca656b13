"""--jacknife through the real CLI at BASELINE config 5 shape (GPU box): a prediction-only sweep of R
replicates on a 2,500 x 200,000 model (250 prediction samples, 10,000 sites replaced per replicate).

The model is trained for 2 epochs only (the sweep does not care how good it is); the sweep time is the
wall-clock difference between a run with --nboots R and one with --nboots 0 (same ingest, same FULL model).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nboots", type=int, default=200)
    ap.add_argument("--snps", type=int, default=200000)
    ap.add_argument("--samples", type=int, default=2500)
    a = ap.parse_args()
    import bench
    from locator_b200 import io

    t0 = time.time()
    x, y = bench.synth(a.samples, a.snps // 4, 1003)   # a quarter of the sites, tiled: content is irrelevant here
    gt1 = np.stack([(x.T >= 1), (x.T >= 2)], axis=2).astype(np.int8)
    gt = np.concatenate([gt1] * 4)
    os.makedirs("/tmp/jk", exist_ok=True)
    z = "/tmp/jk/g.zarr"
    samples = [f"s{i}" for i in range(a.samples)]
    io.write_zarr(z, gt, samples, np.arange(len(gt)) * 10, chunk_variants=16384, compress=False)
    rng = np.random.default_rng(1)
    loc = rng.uniform(0, 50, size=(a.samples, 2))
    loc[rng.choice(a.samples, a.samples // 10, replace=False)] = np.nan
    with open("/tmp/jk/samples.txt", "w") as fh:
        fh.write("sampleID\tx\ty\n")
        for s, (u, v) in zip(samples, loc):
            fh.write(f"{s}\t{'NA' if np.isnan(u) else u}\t{'NA' if np.isnan(v) else v}\n")
    print(f"store: {gt.nbytes / 1e9:.2f} GB int8 built in {time.time() - t0:.1f} s", flush=True)
    secs = {}
    for r in (0, a.nboots):
        cmd = [sys.executable, "-m", "locator_b200", "--zarr", z, "--sample_data", "/tmp/jk/samples.txt", "--out",
               f"/tmp/jk/run{r}", "--seed", "12345", "--jacknife", "--nboots", str(r), "--jacknife_prop", "0.05",
               "--max_epochs", "2", "--patience", "1000", "--keras_verbose", "0"]
        t = time.time()
        p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
        secs[r] = time.time() - t
        if p.returncode != 0:
            print(p.stdout[-1500:], p.stderr[-3000:])
            raise SystemExit(1)
    done = len([f for f in os.listdir("/tmp/jk") if f.startswith(f"run{a.nboots}_boot") and f.endswith("predlocs.txt")])
    sweep = secs[a.nboots] - secs[0]
    print(json.dumps({"workload": f"jacknife sweep: {a.nboots} replicates x 250 prediction samples x {len(gt)} SNPs "
                                  f"(5 % of the sites redrawn per replicate)",
                      "seconds_with_sweep": secs[a.nboots], "seconds_without": secs[0], "sweep_seconds": sweep,
                      "ms_per_replicate": 1e3 * sweep / max(1, a.nboots), "predlocs_files": done,
                      "replicates_per_hour": 3600.0 * a.nboots / sweep}))


if __name__ == "__main__":
    main()
