#!/usr/bin/env python3
"""Where does the host side spend its time?  cProfile of (1) the e2e call of bench.py (model creation + fit on
pageable arrays) and (2) the cfg4 work-queue leg (LOC_TIMING ticks + cProfile).  Diagnostic, not a benchmark."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from locator_b200 import model  # noqa: E402

K, H, L, B = 100_000, 256, 10, 32
x, y = bench.synth(900, K, 1002)
xtr, ytr, xva, yva = x[:810], y[:810], x[810:], y[810:]
mw = model.LocatorModel(K, max_epochs=2, seed=1)
mw.fit(xtr, ytr, epochs=1, validation_data=(xva, yva), patience=10 ** 6)
del mw
torch.cuda.synchronize()


def e2e():
    m2 = model.LocatorModel(K, max_epochs=20, seed=300)
    h = m2.fit(xtr, ytr, epochs=20, validation_data=(xva, yva), patience=10 ** 6)
    torch.cuda.synchronize()
    return h


for rep in range(2):
    t0 = time.perf_counter()
    e2e()
    print(f"e2e 20 epochs: {time.perf_counter() - t0:.4f} s", file=sys.stderr)
prof = cProfile.Profile()
prof.enable()
e2e()
prof.disable()
pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(35)


def timed(name, fn, n=3):
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del r
    print(f"{name}: {min(ts) * 1e3:.2f} ms (min of {n})", file=sys.stderr)


timed("H2D + pack of the 810 x 100k training matrix (pageable)", lambda: model.PackedGenotypes.from_counts(xtr))
timed("LocatorModel create + init", lambda: model.LocatorModel(K, max_epochs=20, seed=5))
mm = model.LocatorModel(K, max_epochs=20, seed=5)
timed("get_weights (download)", lambda: mm.get_weights(), 1)
del mm

os.environ["LOC_BENCH_PROFILE"] = "1"
os.environ["LOC_TIMING"] = "1"
out = bench.work_queue_leg(0, 1, None, torch.cuda.synchronize, 1000, K, 4)
print(out, file=sys.stderr)
