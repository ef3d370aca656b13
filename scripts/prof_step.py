"""Profiling driver (GPU box): one training epoch + one unfused step at a BASELINE shape."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from locator_b200 import model  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n_total, K = bench.WORKLOADS[workload]
    ntr, nva = bench.split_sizes(n_total)
    x, y = bench.synth(ntr + nva, K, 1002)
    m = model.LocatorModel(K, seed=1, max_epochs=4)
    m.bind_train(x[:ntr], y[:ntr])
    m.bind_val(x[ntr:], y[ntr:])
    m.set_schedule(patience=100)
    rng = np.random.default_rng(0)
    # one epoch as the CLI runs it (loc_train_epochs): first-layer forward once, then per step the hidden
    # stack, the small-layer update and the backward kernel that also runs the next step's forward;
    # the epoch's last step is unfused and is followed by the validation pass
    for _ in range(max(1, nsteps // 26)):
        m.train_epochs(np.stack([rng.permutation(ntr)]).astype(np.int32))
    m.train_step(rng.permutation(ntr)[:32])  # one unfused step: forward, hidden, update, plain backward
    print("loss", m.state().last_loss, "val", m.evaluate(x[ntr:ntr + 32], y[ntr:ntr + 32]))


if __name__ == "__main__":
    main()
