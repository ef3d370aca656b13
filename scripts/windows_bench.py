"""--windows through the real CLI on a synthetic zarr genome (GPU box): models/hour at 1 and N GPUs.

BASELINE config 3 shape scaled to fit a quick run: W windows x S SNPs x N samples (10 % without location),
throughput schedule of SURVEY 8(d): --max_epochs 20 --patience 1000.  The zarr store is written raw
(no compressor) under /tmp; wall time of each CLI run is measured from process start to exit.
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
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--windows", type=int, default=16)
    ap.add_argument("--snps", type=int, default=50000)
    ap.add_argument("--samples", type=int, default=2500)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--reuse", action="store_true", help="use the store a previous call left under /tmp/wb")
    a = ap.parse_args()
    import bench
    from locator_b200 import io

    t0 = time.time()
    size = 1_000_000
    z = "/tmp/wb/genome.zarr"
    samples = [f"s{i}" for i in range(a.samples)]
    if a.reuse and os.path.exists(z):
        class gt:  # only the size is reported below
            nbytes = a.windows * a.snps * a.samples * 2
    else:
        x, y = bench.synth(a.samples, a.snps, 5)            # uint8 [n, K] alt counts with spatial structure
        gt1 = np.stack([(x.T >= 1), (x.T >= 2)], axis=2).astype(np.int8)  # [K, n, 2]
        gt = np.concatenate([gt1] * a.windows)               # the same block in every window (content is irrelevant here)
        pos = np.concatenate([w * size + np.sort(np.random.default_rng(w).choice(size, a.snps, replace=False))
                              for w in range(a.windows)])
        os.makedirs("/tmp/wb", exist_ok=True)
        io.write_zarr(z, gt, samples, pos, chunk_variants=16384, compress=False)
    rng = np.random.default_rng(1)
    loc = rng.uniform(0, 50, size=(a.samples, 2))
    loc[rng.choice(a.samples, a.samples // 10, replace=False)] = np.nan
    with open("/tmp/wb/samples.txt", "w") as fh:
        fh.write("sampleID\tx\ty\n")
        for s, (u, v) in zip(samples, loc):
            fh.write(f"{s}\t{'NA' if np.isnan(u) else u}\t{'NA' if np.isnan(v) else v}\n")
    print(f"store: {gt.nbytes / 1e9:.2f} GB int8, {a.windows} windows, built in {time.time() - t0:.1f} s", flush=True)
    res = {}
    for n in sorted({1, a.gpus}):
        out = f"/tmp/wb/run{n}"
        cmd = [sys.executable, "-m", "locator_b200", "--zarr", z, "--sample_data", "/tmp/wb/samples.txt", "--out", out,
               "--seed", "1", "--windows", "--window_size", str(size), "--window_stop", str(a.windows * size),
               "--max_epochs", str(a.epochs), "--patience", "1000", "--keras_verbose", "0", "--gpus", str(n)]
        t = time.time()
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
        dt = time.time() - t
        done = len([f for f in os.listdir("/tmp/wb") if f.startswith(f"run{n}_") and f.endswith("predlocs.txt")])
        res[n] = {"seconds": dt, "windows_done": done, "models_per_hour": 3600.0 * done / dt, "rc": r.returncode}
        if r.returncode != 0:
            print(r.stdout[-1500:], r.stderr[-3000:])
        if os.environ.get("LOC_TIMING"):
            print(f"process started at {t:.3f}")
            print("\n".join(l for l in r.stderr.splitlines() if "loc-timing" in l), flush=True)
        print(n, "GPU(s):", json.dumps(res[n]), flush=True)
    print(json.dumps({"workload": f"{a.windows} windows x {a.snps} SNPs x {a.samples} samples, {a.epochs} epochs each",
                      "results": res}))


if __name__ == "__main__":
    main()
