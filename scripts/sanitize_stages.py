"""compute-sanitizer driver (GPU box): every kernel except the hidden-stack cluster kernels, whose
shared::cta -> shared::cluster bulk copies memcheck reports as invalid writes (tool limitation)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model, genotypes as G

rng = np.random.default_rng(0)
for K, impl in ((4136, None), (1000, "simt")):
    if impl:
        os.environ["LOC_L1_IMPL"] = impl
    x = rng.integers(0, 3, size=(70, K), dtype=np.uint8); y = rng.normal(size=(70, 2)).astype(np.float32)
    m = model.LocatorModel(K, seed=1); m.bind_train(x, y); m.set_schedule()
    for nb in (32, 21):
        rows = rng.permutation(70)[:nb]
        for st in ((0, 2, 4, 3) if m.impl == "tcgen05" else (0, 2, 3)):
            m.debug_stage(st, rows)
    w = m.get_weights(); m.set_weights(w); m.snapshot()
    print(K, m.impl, "W1 finite:", bool(np.isfinite(m.get_weights()[4]).all()))
gt = (rng.uniform(size=(500, 77, 2)) < 0.3).astype(np.int8); gt[rng.uniform(size=(500, 77)) < 0.05] = -1
g, na, alt, miss, keep = G.site_stats(gt, 2)
idx = np.flatnonzero(keep.cpu().numpy()); p = G.pack_sites(g, idx)
p.take_rows(rng.permutation(77)[:50]).take_cols(rng.integers(0, len(idx), len(idx))).to_counts()
p.replace_cols(np.arange(5), rng.integers(0, 3, size=(5, 77), dtype=np.uint8)); p.patch([1, 2], [3, 4], [1, 2])
print("ingest ok")
