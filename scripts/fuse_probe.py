#!/usr/bin/env python3
"""Where do the ~9 us go that the fused next-step forward adds to the first-layer backward + Adam kernel?
Times loc_debug_stage 4 (fused) with parts of the forward warps' work switched off (LOC_FUSE_DEBUG bits:
1 no forward MMA, 2 no gamma/beta Adam + stores, 4 no xhat build, 8 stage release does not wait for the forward
MMA, 32 no final partial-tile epilogue) against stage 2 (plain).  Results of runs with bits set are garbage:
timing only.  One JSON line per setting."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model, _cabi  # noqa: E402

lib = _cabi.lib
K, B = int(os.environ.get("PROBE_K", "100000")), 32
ctas = os.environ.get("PROBE_CTAS") or None
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
reps = int(os.environ.get("PROBE_REPS", "40"))
settings = [(2, 0), (4, 0), (4, 1), (4, 2), (4, 4), (4, 32), (4, 7), (4, 39), (4, 0), (2, 0)]
if os.environ.get("PROBE_FLAGS"):  # "stage:flags,stage:flags,..."
    settings = [tuple(int(v) for v in p.split(":")) for p in os.environ["PROBE_FLAGS"].split(",")]
for bwd_stage, flags in settings:
    os.environ["LOC_FUSE_DEBUG"] = str(flags)
    times = []
    for r in range(reps + 3):
        stage(1)  # hidden stack: advances t, so the walk direction alternates as in training
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stage(bwd_stage)
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            times.append(e0.elapsed_time(e1) * 1000.0)
    print(json.dumps({"K": K, "l1_ctas": ctas, "stage": bwd_stage, "flags": flags, "us_mean": float(np.mean(times)),
                      "us_min": float(np.min(times)), "us_median": float(np.median(times))}), flush=True)
os.environ.pop("LOC_FUSE_DEBUG", None)
