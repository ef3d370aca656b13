"""Trained-accuracy parity (north_star's second correctness check): the CUDA path and the CPU oracle trained
with the reference's DEFAULT schedule (max_epochs 5000, patience 100 -> early stopping, LR halving, reload of
the best epoch; /root/reference/locator/locator.py:330-394) on the reference's own example data
(data/test_genotypes.vcf.gz, the seed-12345 split), from identical initial weights, batch orders and
dropout masks.  Compared: the median validation error in map units (the number the reference prints,
locator.py:437-467; README.md:147-155 reports 3.30 for its own unseeded run).

Margin.  Individual weights cannot agree after thousands of Adam steps (tests/test_gpu_baseline_shapes.py
explains why); the trained models are compared as predictors: |median_cuda - median_oracle| <= MARGIN map
units on the 50 x 50 landscape, the same bound between the oracle's fp32 and tf32 numerics.
"""
import json
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
VCF = os.path.join(HERE, "golden", "data", "test_genotypes.vcf.gz")
SAMPLES = os.path.join(HERE, "golden", "data", "test_sample_data.txt")
MARGIN = 1.0  # map units


def _report(**kw):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_accuracy.jsonl"), "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass
    print("ACCURACY", json.dumps(kw), file=sys.stderr)


def _median_error(pred, testlocs, norm):
    meanlong, sdlong, meanlat, sdlat = norm
    p = np.stack([pred[:, 0] * sdlong + meanlong, pred[:, 1] * sdlat + meanlat], axis=1)
    t = np.stack([testlocs[:, 0] * sdlong + meanlong, testlocs[:, 1] * sdlat + meanlat], axis=1)
    d = np.sqrt(((p - t) ** 2).sum(axis=1))
    return float(np.median(d)), float(np.mean(d))


def test_trained_median_validation_error_matches_oracle(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from locator_b200 import locator as L, model as M
    from oracle import model_ref

    L.set_args(L.build_parser().parse_args(["--vcf", VCF, "--sample_data", SAMPLES, "--out", str(tmp_path / "acc"),
                                            "--seed", "12345"]))
    np.random.seed(12345)
    genotypes, samples = L.load_genotypes()
    sample_data, locs = L.sort_samples(samples, genotypes)
    meanlong, sdlong, meanlat, sdlat, nlocs = L.normalize_locs(locs)
    norm = (meanlong, sdlong, meanlat, sdlat)
    ac = L.filter_snps(genotypes)
    train, test, traingen, testgen, trainlocs, testlocs, pred, predgen = L.split_train_test(ac, nlocs)
    xt, xv = traingen.to_numpy(), testgen.to_numpy()
    yt, yv = trainlocs.astype(np.float32), testlocs.astype(np.float32)
    K, ntr = xt.shape[1], xt.shape[0]
    assert (K, ntr, len(xv)) == (5830, 405, 45)
    max_epochs, patience = 5000, 100  # the reference's defaults
    results = []
    for seed, oracle_modes in ((1, ("fp32", "tf32")), (2, ("fp32",))):
        prng = np.random.default_rng(seed)
        perms = np.stack([prng.permutation(ntr) for _ in range(max_epochs)]).astype(np.int32)
        m = M.LocatorModel(K, seed=seed, max_epochs=max_epochs)
        w0 = m.get_weights()
        t0 = time.time()
        h = m.fit(xt, yt, epochs=max_epochs, validation_data=(xv, yv), patience=patience, perms=perms)
        m.restore_best()
        t_cuda = time.time() - t0
        med_c, mean_c = _median_error(m.predict(xv), testlocs, norm)
        row = {"seed": seed, "cuda": {"epochs": len(h.history["loss"]), "median": med_c, "mean": mean_c,
                                      "best_val_loss": min(h.history["val_loss"]), "seconds": t_cuda}}
        for mode in oracle_modes:
            ref = model_ref.RefLocator(K, 256, 10, dropout=0.25, weights=w0, numerics=mode)
            t0 = time.time()
            hr = model_ref.fit(ref, xt, yt, xv, yv, max_epochs, batch_size=32, patience=patience, perms=perms, seed=seed)
            med_r, mean_r = _median_error(ref.predict(xv), testlocs, norm)
            row["oracle_" + mode] = {"epochs": len(hr["loss"]), "median": med_r, "mean": mean_r,
                                     "best_val_loss": min(hr["val_loss"]), "seconds": time.time() - t0}
        _report(test="trained_accuracy_fixture", readme_median=3.30, margin=MARGIN, **row)
        results.append(row)
        del m
    for row in results:
        for k, v in row.items():
            if k.startswith("oracle_"):
                assert abs(row["cuda"]["median"] - v["median"]) <= MARGIN, row
                # both stop early, well before max_epochs, within a factor of two of each other
                assert v["epochs"] < max_epochs and row["cuda"]["epochs"] < max_epochs, row
                assert 0.5 <= row["cuda"]["epochs"] / v["epochs"] <= 2.0, row
        # and the README's magnitude: the reference's own run reports 3.30
        assert row["cuda"]["median"] <= 3.30 + 1.5, row
