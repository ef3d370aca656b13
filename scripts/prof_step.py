"""Profiling driver (GPU box): a few optimizer steps + one validation chunk at a BASELINE shape."""
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
    for s in range(nsteps):
        m.train_step(rng.permutation(ntr)[:32])
    m.debug_stage(4, rng.permutation(ntr)[:32])  # the fused backward + next forward, as train_epochs runs it
    print("loss", m.state().last_loss, "val", m.evaluate(x[ntr:ntr + 32], y[ntr:ntr + 32]))


if __name__ == "__main__":
    main()
