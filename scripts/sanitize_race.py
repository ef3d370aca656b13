"""compute-sanitizer racecheck / synccheck driver (GPU box): every step kernel of the tcgen05 path with several
64-SNP tiles per CTA (K = 4,096 on 16 first-layer CTAs = 4 tiles each; full and ragged batch): first-layer forward,
hidden stack, plain backward + Adam, fused backward + next forward, small-layer update, the wide inference
forward + grouped hidden stack, a two-model ring group.  usage: sanitize_race.py [skip_hidden]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locator_b200 import model, _cabi  # noqa: E402
import torch  # noqa: E402

skip_hidden = len(sys.argv) > 1 and sys.argv[1] == "skip_hidden"
rng = np.random.default_rng(0)
K, n = 4096, 96
x = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
y = rng.normal(size=(n, 2)).astype(np.float32)
m = model.LocatorModel(K, seed=1, l1_ctas=16, max_epochs=4)
m.bind_train(x, y)
m.bind_val(x[:70], y[:70])
m.set_schedule()
stages = (0, 2, 4, 3) if skip_hidden else (0, 1, 2, 4, 3)
for nb in (32, 21):
    rows = rng.permutation(n)[:nb]
    for st in stages:
        m.debug_stage(st, rows)
    torch.cuda.synchronize()
print("stages done; W1 finite:", bool(np.isfinite(m.get_weights()[4]).all()), flush=True)
if not skip_hidden:
    m.train_epochs(np.stack([rng.permutation(n)]).astype(np.int32))  # 3 steps, fused forwards, wide validation pass
    torch.cuda.synchronize()
    print("epoch done: loss", m.state().last_loss, "val", m.state().last_val_loss, flush=True)
    os.environ["LOC_GROUP_SCHEDULE"] = "ring"
    ms = [model.LocatorModel(K, seed=s, l1_ctas=16, max_epochs=4) for s in (2, 3)]
    for q in ms:
        q.bind_train(x, y)
        q.bind_val(x[:70], y[:70])
        q.set_schedule()
    perms = [torch.as_tensor(np.stack([rng.permutation(n)]).astype(np.int32)).cuda() for _ in ms]
    handles = (ctypes.c_void_p * 2)(*[q._h for q in ms])
    pp = (ctypes.c_void_p * 2)(*[p.data_ptr() for p in perms])
    _cabi.check(_cabi.lib.loc_group_train_epochs(handles, 2, pp, 1, torch.cuda.current_stream().cuda_stream), "group")
    torch.cuda.synchronize()
    print("ring group done: loss", ms[0].state().last_loss, flush=True)
