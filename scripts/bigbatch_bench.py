"""GPU box: epochs of model.fit's schedule at BASELINE config[1] (1,000 x 100,000) for --batch_size 32 .. 256.

Batches above 32 rows take the chunked path (csrc/bigbatch.cu).  Prints one JSON line per batch size:
samples/s over whole epochs (validation pass and callbacks included), ms per optimizer step.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from locator_b200 import model, _cabi  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n_total, K = bench.WORKLOADS[workload]
    ntr, nva = bench.split_sizes(n_total)
    x, y = bench.synth(ntr + nva, K, 1002)
    rng = np.random.default_rng(0)
    epochs = 6
    for B in (32, 64, 128, 256):
        m = model.LocatorModel(K, batch_size=B, seed=1, max_epochs=64)
        m.bind_train(x[:ntr], y[:ntr])
        m.bind_val(x[ntr:], y[ntr:])
        m.set_schedule(patience=1000)
        perms = lambda e: np.stack([rng.permutation(ntr) for _ in range(e)]).astype(np.int32)  # noqa: E731
        m.train_epochs(perms(2))  # warm-up
        torch.cuda.synchronize()
        l0 = _cabi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p = perms(epochs)
        e0.record()
        m.train_epochs(p)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        spe = -(-ntr // B)
        st = m.state()
        print(json.dumps({"workload": workload, "batch_size": B, "epochs": epochs, "steps_per_epoch": spe,
                          "ms_per_epoch": round(ms / epochs, 4), "ms_per_step": round(ms / epochs / spe, 4),
                          "samples_per_s": round(ntr * epochs / (ms * 1e-3), 1), "launches": _cabi.launch_count() - l0,
                          "last_loss": round(float(st.last_loss), 5), "last_val_loss": round(float(st.last_val_loss), 5)}),
              flush=True)
        del m


if __name__ == "__main__":
    main()
