#!/usr/bin/env python3
"""How much shorter is the first-layer backward + Adam kernel when the head of its walk is already in L2?

HBM idles while the hidden stack runs (43 us of a 167 us step); this probe prefetches chunks [skip, skip + n)
of every CTA's walk (loc_debug_stage 5, LOC_PREFETCH="skip,n") after a hidden-stack launch and times the
backward that follows, for a grid of (skip, n).  Output: one JSON line per setting.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model, _cabi  # noqa: E402

lib = _cabi.lib
K, H, B = int(os.environ.get("PROBE_K", "100000")), 256, 32
ctas = os.environ.get("PROBE_CTAS")
rng = np.random.default_rng(0)
n = 128
x = rng.binomial(2, rng.uniform(0.05, 0.95, K), size=(n, K)).astype(np.uint8)
y = rng.normal(size=(n, 2)).astype(np.float32)
m = model.LocatorModel(K, seed=1, l1_ctas=int(ctas) if ctas else None)
m.bind_train(x, y)
m.set_schedule(patience=100)
rows = torch.as_tensor(rng.permutation(n)[:B].astype(np.int32)).cuda()
stream = torch.cuda.current_stream().cuda_stream


def stage(s):
    _cabi.check(lib.loc_debug_stage(m._h, s, rows.data_ptr(), B, stream), "loc_debug_stage")


stage(0)
stage(1)
for _ in range(5):
    stage(2)
torch.cuda.synchronize()
settings = [(0, 0)] + [(s, c) for s in (0, 8, 16) for c in (8, 16, 24, 32)]
reps = int(os.environ.get("PROBE_REPS", "30"))
for bwd_stage in (2, 4):
    for skip, cnt in settings:
        os.environ["LOC_PREFETCH"] = f"{skip},{cnt}"
        times = []
        for r in range(reps + 3):
            stage(1)  # hidden stack: advances t, so the walk direction alternates as in training
            if cnt:
                stage(5)
            torch.cuda._sleep(200000)  # ~100 us: the prefetches land before the backward starts
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            stage(bwd_stage)
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                times.append(e0.elapsed_time(e1) * 1000.0)
        print(json.dumps({"K": K, "l1_ctas": ctas, "stage": bwd_stage, "skip": skip, "chunks": cnt,
                          "prefetched_MB": cnt * 24576 * (int(ctas) if ctas else 148) / 1e6,
                          "bwd_us_mean": float(np.mean(times)), "bwd_us_min": float(np.min(times))}), flush=True)
