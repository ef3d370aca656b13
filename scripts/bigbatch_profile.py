"""GPU box, under ncu: two epochs of model.fit at BASELINE config[1] with --batch_size argv[1] (launch lists / captures of the
large-batch kernels, csrc/bigbatch.cu)."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from locator_b200 import model
B = int(sys.argv[1])
n_total, K = bench.WORKLOADS["cfg2"]
ntr, nva = bench.split_sizes(n_total)
x, y = bench.synth(ntr + nva, K, 1002)
rng = np.random.default_rng(0)
m = model.LocatorModel(K, batch_size=B, seed=1, max_epochs=8)
m.bind_train(x[:ntr], y[:ntr]); m.bind_val(x[ntr:], y[ntr:]); m.set_schedule(patience=1000)
m.train_epochs(np.stack([rng.permutation(ntr) for _ in range(2)]).astype(np.int32))
print(m.state().last_loss)
