"""Where does LocatorModel.fit spend host time? (GPU box)  Repeats bench.py's e2e leg with phase timers."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from locator_b200 import model

n_total, K = bench.WORKLOADS["cfg2"]
ntr, nva = bench.split_sizes(n_total)
x, y = bench.synth(ntr + nva, K, 1002)
xtr_p = torch.from_numpy(x[:ntr]).pin_memory(); xva_p = torch.from_numpy(x[ntr:]).pin_memory()
for rep in range(8):
    t = [time.perf_counter()]
    m = model.LocatorModel(K, max_epochs=11, seed=300 + rep)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    g = m.bind_train(xtr_p, y[:ntr]); m.bind_val(xva_p, y[ntr:]); m.set_schedule(patience=10 ** 6)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    rng = np.random.default_rng(rep)
    p = np.stack([rng.permutation(ntr) for _ in range(10)]).astype(np.int32)
    t.append(time.perf_counter())
    m.train_epochs(p); t.append(time.perf_counter())
    st = m.state(); t.append(time.perf_counter())
    rows = m.history_rows(st.epoch); t.append(time.perf_counter())
    del m; torch.cuda.synchronize(); t.append(time.perf_counter())
    names = ["create+init", "bind(H2D+pack)", "perms", "enqueue", "sync(state)", "history", "destroy"]
    print(rep, " ".join(f"{n}={1e3*(b-a):.1f}ms" for n, a, b in zip(names, t, t[1:])), flush=True)
