"""compute-sanitizer driver (GPU box) for the large-batch step (csrc/bigbatch.cu): steps of 40..250 rows on the tcgen05
path (width 256: wide forward, grouped hidden stack, mma.sync backward on the tiled W1 layout), and on the CUDA-core
path with both column-group widths of the backward (width 64 -> 64 columns, width 96 -> 32), ragged chunks included.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck [--kernel-name kns=k_bb] python scripts/sanitize_bigbatch.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model  # noqa: E402
import torch  # noqa: E402

rng = np.random.default_rng(0)
for K, H, L, B, sizes in ((2100, 256, 10, 250, (250, 70)), (1000, 64, 4, 80, (80, 33)), (523, 96, 3, 40, (40, 40))):
    n = 260
    x = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
    y = rng.normal(size=(n, 2)).astype(np.float32)
    m = model.LocatorModel(K, width=H, nlayers=L, batch_size=B, seed=1, max_epochs=2)
    m.bind_train(x, y)
    m.bind_val(x[:40], y[:40])
    m.set_schedule()
    for nb in sizes:
        m.train_step(rng.permutation(n)[:nb])
    torch.cuda.synchronize()
    st = m.state()
    print(f"K {K} width {H} batch {B}: impl {m.impl} t {st.t} loss {st.last_loss:.5f} finite {bool(np.isfinite(m.get_weights()[4]).all())}",
          flush=True)
    del m
print("done", flush=True)
