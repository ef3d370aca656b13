"""One model sharded over SNP columns (tensor parallelism): two shards exchanging the
first-layer tile through an all_reduce must reproduce the unsharded model -- identical initial weights,
identical replicas of the hidden stack on every shard, losses / predictions within the tf32 tolerance (the
only difference is the fp32 summation order of the split-K partial sums).  Two exchanges are covered: the host
hook (torch.distributed all_reduce) and the peer-memory kernels (loc_tp_*: push over cudaIpc-mapped buffers)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["hook", "peer"])
def test_two_shards_match_the_unsharded_model(tmp_path, mode):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = tmp_path / "tp.json"
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_tp_runner.py"), str(out), backend, mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    if res["impl"] != "tcgen05":
        pytest.skip("sharded models need the tcgen05 kernels")
    assert res["bounds"] == [0, 3008]
    assert res["init_equal"], "shards must initialise exactly as slices of the unsharded layer"
    assert res["replicas_identical"], "every shard must hold bitwise identical hidden stacks / losses / predictions"
    ht, hf = np.array(res["hist_tp"]), np.array(res["hist_full"])
    np.testing.assert_allclose(ht[:, :2], hf[:, :2], rtol=5e-3)
    assert np.array_equal(ht[:, 2], hf[:, 2])  # learning-rate column
    assert res["pred_maxdiff"] < 5e-3
    assert res["w1_update_rel"] < 0.05
    # ... and the CPU oracle (oracle/model_ref.py) directly, not only the unsharded CUDA model: 3 epochs of 3 steps
    for numerics, o in res["oracle"].items():
        assert o["init_equal"], "shards must initialise exactly as slices of the oracle's Philox layer"
        np.testing.assert_allclose(ht[:, 0], o["hist"][0], rtol=1e-2, err_msg=f"loss vs {numerics} oracle")
        np.testing.assert_allclose(ht[:, 1], o["hist"][1], rtol=3e-2, err_msg=f"val_loss vs {numerics} oracle")
        assert o["pred_maxdiff"] < 5e-2, (numerics, o)
        assert o["w1_update_rel"] < 0.15, (numerics, o)
    print("TP vs oracle:", json.dumps(res["oracle"]), file=sys.stderr)
