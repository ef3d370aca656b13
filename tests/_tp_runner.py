"""Two-shard tensor-parallel run (launched by tests/test_gpu_tp.py through torch.distributed.run).

Both ranks use the visible GPU(s) round-robin and exchange through the process group's all_reduce
(gloo stages CUDA tensors through the host, so two ranks can share one GPU in the test; on a multi-GPU
box the same code runs over NCCL).  Rank 0 also trains the unsharded model and writes the comparison.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    out = sys.argv[1]
    backend = sys.argv[2] if len(sys.argv) > 2 else "gloo"
    mode = sys.argv[3] if len(sys.argv) > 3 else "hook"  # hook: all_reduce through the process group; peer: NVLink peer memory
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend)
    from locator_b200 import model as M

    rng = np.random.default_rng(17)
    K, ntr, nva, epochs = 6000, 96, 40, 3
    p = rng.uniform(0.05, 0.95, size=K)
    x = rng.binomial(2, p, size=(ntr, K)).astype(np.uint8)
    xv = rng.binomial(2, p, size=(nva, K)).astype(np.uint8)
    y = rng.normal(size=(ntr, 2)).astype(np.float32)
    yv = rng.normal(size=(nva, 2)).astype(np.float32)
    perms = np.stack([rng.permutation(ntr) for _ in range(epochs)]).astype(np.int32)

    k0, k1 = M.shard_bounds(K, rank, world)
    m = M.LocatorModel(k1 - k0, seed=5, max_epochs=epochs, shard=(k0, K),
                       exchange="peer" if mode == "peer" else M.all_reduce_exchange())
    res = {"impl": m.impl, "bounds": [k0, k1]}
    w_init = m.get_weights()[4]
    m.bind_train(x[:, k0:k1], y)
    m.bind_val(xv[:, k0:k1], yv)
    m.set_schedule(patience=100)
    m.train_epochs(perms)
    torch.cuda.synchronize()
    m.check_peers()
    hist = np.asarray(m.history_rows(epochs), dtype=np.float64)
    pred = m.predict(xv[:, k0:k1])
    w1 = m.get_weights()[4]
    gathered = [None] * world
    dist.all_gather_object(gathered, {"w_init": w_init, "w1": w1, "hist": hist, "pred": pred,
                                      "small": m.get_weights()[6]})
    if rank == 0:
        full = M.LocatorModel(K, seed=5, max_epochs=epochs)
        f_init = full.get_weights()[4]
        full.bind_train(x, y)
        full.bind_val(xv, yv)
        full.set_schedule(patience=100)
        full.train_epochs(perms)
        torch.cuda.synchronize()
        f_hist = np.asarray(full.history_rows(epochs), dtype=np.float64)
        f_pred = full.predict(xv)
        f_w1 = full.get_weights()[4]
        s_init = np.concatenate([g["w_init"] for g in gathered])
        s_w1 = np.concatenate([g["w1"] for g in gathered])
        # the CPU oracle from the same start (Philox dropout masks keyed by the model seed), both numerics
        from oracle import model_ref

        oracle = {}
        for numerics in ("fp32", "tf32"):
            ws = model_ref.init_weights(K, 256, 10, seed=5)
            ref = model_ref.RefLocator(K, 256, 10, dropout=0.25, weights=ws, numerics=numerics)
            hr = model_ref.fit(ref, x, y, xv, yv, epochs, batch_size=32, patience=100, perms=perms, seed=5)
            r_w1 = ref.get_weights()[4]
            oracle[numerics] = {
                "hist": [hr["loss"], hr["val_loss"]],
                "pred_maxdiff": float(np.abs(gathered[0]["pred"] - ref.predict(xv)).max()),
                "w1_update_rel": float(np.linalg.norm(s_w1 - r_w1) / np.linalg.norm(r_w1 - ws[4])),
                "init_equal": bool(np.array_equal(s_init, ws[4])),
            }
        res["oracle"] = oracle
        res.update({
            "init_equal": bool(np.array_equal(s_init, f_init)),
            "replicas_identical": bool(all(np.array_equal(g["hist"], gathered[0]["hist"]) and
                                           np.array_equal(g["pred"], gathered[0]["pred"]) and
                                           np.array_equal(g["small"], gathered[0]["small"]) for g in gathered)),
            "hist_tp": gathered[0]["hist"].tolist(), "hist_full": f_hist.tolist(),
            "pred_maxdiff": float(np.abs(gathered[0]["pred"] - f_pred).max()),
            "w1_update_rel": float(np.linalg.norm(s_w1 - f_w1) / np.linalg.norm(f_w1 - f_init)),
        })
        with open(out, "w") as fh:
            json.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
