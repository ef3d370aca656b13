"""Clock trace of k_hidden (GPU box): LOC_HID_TRACE=1 python scripts/hid_trace.py"""
import os, sys
import numpy as np
os.environ["LOC_HID_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model, _cabi
K = 100000
rng = np.random.default_rng(0)
x = rng.integers(0, 3, size=(64, K), dtype=np.uint8); y = rng.normal(size=(64, 2)).astype(np.float32)
m = model.LocatorModel(K, seed=1); m.bind_train(x, y); m.set_schedule()
for s in range(3):
    m.train_step(rng.permutation(64)[:32])
a = np.empty(16 * 256 * 2, np.float32)
_cabi.lib.loc_debug_read(m._h, 3, a.ctypes.data, a.size, 0)
t = a.view(np.int64).reshape(16, 256)
for r in (0, 7, 15):
    row = t[r]; n = int((row != 0).sum()); d = np.diff(row[:n])
    print(f"cta {r}: {n} marks, total {row[n-1]-row[0]} cycles")
    print("  ", " ".join(str(int(v)) for v in d))
