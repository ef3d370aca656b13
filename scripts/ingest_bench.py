"""Ingest kernels at BASELINE config-3 size (GPU box): device time and achieved bytes/s."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import genotypes as G

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))

nvar, N = 400_000, 2500   # ~200k SNPs survive the filter
gen = torch.Generator(device="cuda").manual_seed(1)
p = torch.rand(nvar, 1, 1, device="cuda", generator=gen) ** 2
gt = (torch.rand(nvar, N, 2, device="cuda", generator=gen) < p).to(torch.int8)
out = {}
g, na, alt, miss, keep = G.site_stats(gt, 2)
t = timed(lambda: G.site_stats(gt, 2))
out["site_stats"] = {"ms": t, "GBps": gt.numel() / t / 1e6}
idx = torch.nonzero(keep).flatten()
K = int(idx.numel())
t = timed(lambda: G.pack_sites(gt, idx))
out["pack_sites"] = {"ms": t, "K": K, "GBps": (2 * K * N + K * N / 4) / t / 1e6}
packed = G.pack_sites(gt, idx)
rows = torch.randperm(N, device="cuda")[:2025]
t = timed(lambda: packed.take_rows(rows))
out["gather_rows(train split)"] = {"ms": t, "GBps": 2 * 2025 * packed.row_words * 4 / t / 1e6}
cols = torch.randint(0, K, (K,), device="cuda")
tr = packed.take_rows(rows)
t = timed(lambda: tr.take_cols(cols))
out["gather_cols(bootstrap)"] = {"ms": t, "GBps": (2 * 2025 * packed.row_words * 4 + K * 8) / t / 1e6}
print(json.dumps(out, indent=1))
