#!/usr/bin/env python3
"""Tensor-core roofline of the wide inference forward at BASELINE config 5's shape: one
jacknife replicate = predict(250 x 200,000) = 2 * 250 * 200,000 * 256 = 25.6 GFLOP of first-layer products
(200 replicates: 5.12 TFLOP) against one 204.8 MB stream of W1.  Times loc_predict with CUDA events (wide path and,
for comparison, the chunk-by-chunk path that streams W1 once per 32 rows) and the first-layer kernel alone.
Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locator_b200 import model, _cabi  # noqa: E402

K, n = int(os.environ.get("WIDE_K", "200000")), int(os.environ.get("WIDE_N", "250"))
rng = np.random.default_rng(3)
x = rng.binomial(2, rng.uniform(0.05, 0.95, K), size=(n, K)).astype(np.uint8)
m = model.LocatorModel(K, seed=1)
g = model.PackedGenotypes.from_counts(x)
out = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
lib = _cabi.lib


def predict():
    _cabi.check(lib.loc_predict(m._h, g.ptr, g.n, g.row_words, out.data_ptr(), stream), "loc_predict")


def timeit(reps=30):
    for _ in range(5):
        predict()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        predict()
        b.record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return float(np.mean(ts)), float(np.min(ts))


wide_ms, wide_min = timeit()
y_wide = out.cpu().numpy().copy()
os.environ["LOC_NO_WIDE"] = "1"
narrow_ms, narrow_min = timeit()
y_narrow = out.cpu().numpy().copy()
del os.environ["LOC_NO_WIDE"]
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
flop = 2.0 * n * K * 256
peak_tf = float(peaks.get("bf16_tflops_sustained", 1378.9))
hbm = float(peaks.get("hbm_gbs", 6458.1))
print(json.dumps({
    "workload": f"predict({n} x {K}), width 256, tf32 operands / fp32 accumulate",
    "flop_per_replicate": flop, "w1_bytes": 4.0 * K * 256,
    "wide": {"ms_mean": wide_ms, "ms_min": wide_min, "tflops": flop / (wide_ms * 1e-3) / 1e12,
             "frac_of_bf16_sustained_peak": flop / (wide_ms * 1e-3) / 1e12 / peak_tf,
             "w1_stream_gbs": 4.0 * K * 256 / (wide_ms * 1e-3) / 1e9, "frac_of_hbm_peak": 4.0 * K * 256 / (wide_ms * 1e-3) / 1e9 / hbm},
    "chunks_of_32": {"ms_mean": narrow_ms, "ms_min": narrow_min, "tflops": flop / (narrow_ms * 1e-3) / 1e12,
                     "w1_streams": -(-n // 32)},
    "speedup": narrow_ms / wide_ms, "max_abs_diff_wide_vs_chunks": float(np.abs(y_wide - y_narrow).max()),
    "peaks": {"bf16_tflops_sustained": peak_tf, "hbm_gbs": hbm,
              "note": "kind::tf32 runs at half the bf16 rate: the tf32 ceiling is 0.5 of the bf16 figure"},
    "sweep_of_200_replicates_ms_device_only": 200 * wide_ms}))
