"""GPU end-to-end edge cases of the `locator` command on small synthetic inputs: ragged batches, training sets
smaller than a batch, no prediction samples, minimum depth, non-default widths / batch sizes, missing calls,
the --matrix input -- the shapes the reference's loops accept (locator.py:295-308, :367-376, :414-435)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from locator_b200 import locator

    return locator


def _write_inputs(tmp_path, n, nvar, n_na, seed=0, missing=0.0, matrix=False):
    """A small VCF (or --matrix table) with spatial structure + its sample_data table."""
    rng = np.random.default_rng(seed)
    loc = rng.uniform(0, 50, size=(n, 2))
    z = (loc - 25.0) / 14.0
    c, a, b = rng.normal(0, 1.0, nvar), rng.normal(0, 0.8, nvar), rng.normal(0, 0.8, nvar)
    p = 1.0 / (1.0 + np.exp(-(c[:, None] + a[:, None] * z[None, :, 0] + b[:, None] * z[None, :, 1])))  # [nvar, n]
    gt = (rng.uniform(size=(nvar, n, 2)) < p[:, :, None]).astype(np.int8)
    if missing:
        gt[rng.uniform(size=(nvar, n)) < missing] = -1
    names = [f"ind{i}" for i in range(n)]
    if matrix:
        path = str(tmp_path / "g.txt")
        counts = (gt == 1).sum(axis=2)  # [nvar, n]
        with open(path, "w") as fh:
            fh.write("sampleID\t" + "\t".join(f"s{k}" for k in range(nvar)) + "\n")
            for i, nm in enumerate(names):
                fh.write(nm + "\t" + "\t".join(str(int(v)) for v in counts[:, i]) + "\n")
    else:
        path = str(tmp_path / "g.vcf")
        with open(path, "w") as fh:
            fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n")
            for k in range(nvar):
                calls = "\t".join("./." if gt[k, i, 0] < 0 else f"{gt[k, i, 0]}|{gt[k, i, 1]}" for i in range(n))
                fh.write(f"1\t{100 + 10 * k}\t.\tA\tT\t.\tPASS\t.\tGT\t{calls}\n")
    sd = str(tmp_path / "samples.txt")
    na = set(rng.choice(n, n_na, replace=False).tolist()) if n_na else set()
    with open(sd, "w") as fh:
        fh.write("sampleID\tx\ty\n")
        for i, nm in enumerate(names):
            fh.write(f"{nm}\tNA\tNA\n" if i in na else f"{nm}\t{loc[i, 0]}\t{loc[i, 1]}\n")
    return path, sd, sorted(na)


def _rows(path):
    lines = open(path).read().strip().split("\n")
    return lines[0], lines[1:]


@pytest.mark.parametrize("flags,n,n_na", [
    ([], 70, 7),                                            # ragged last batch (57 train / 6 val)
    ([], 30, 4),                                            # training set smaller than one batch of 32
    (["--batch_size", "8"], 45, 5),                         # small batches, ragged
    (["--batch_size", "64"], 200, 10),                      # steps of more than 32 rows (csrc/bigbatch.cu), ragged
    (["--batch_size", "48", "--width", "64", "--nlayers", "4"], 120, 6),  # ... on the CUDA-core kernels
    (["--batch_size", "256"], 100, 5),                      # batch larger than the training set
    (["--nlayers", "2"], 40, 4),                            # minimum depth: no hidden Dense before / after the dropout
    (["--width", "64", "--nlayers", "4"], 40, 4),           # CUDA-core kernels (width != 256)
    (["--dropout_prop", "0"], 40, 4),
    (["--train_split", "0.5"], 40, 4),
])
def test_cli_shapes(L, tmp_path, flags, n, n_na):
    vcf, sd, na = _write_inputs(tmp_path, n, 300, n_na)
    out = str(tmp_path / "o")
    assert L.main(["--vcf", vcf, "--sample_data", sd, "--out", out, "--seed", "3", "--max_epochs", "6", "--patience", "3",
                   "--keras_verbose", "0"] + flags) == 0
    header, rows = _rows(out + "_predlocs.txt")
    assert header == "x,y,sampleID" and [r.split(",")[2] for r in rows] == [f"ind{i}" for i in na]
    xy = np.array([[float(v) for v in r.split(",")[:2]] for r in rows])
    assert np.all(np.isfinite(xy))
    hh, hrows = _rows(out + "_history.txt")
    assert hh.split("\t")[:2] == ["loss", "val_loss"] and 1 <= len(hrows) <= 6
    assert all(np.isfinite(float(v)) for r in hrows for v in r.split("\t"))


def test_no_prediction_samples(L, tmp_path):
    """Every sample has a location: the reference's pred index is empty (it crashes there, locator.py:298,307);
    here the run trains, validates and writes a header-only predlocs file."""
    vcf, sd, na = _write_inputs(tmp_path, 40, 200, 0)
    out = str(tmp_path / "o")
    assert L.main(["--vcf", vcf, "--sample_data", sd, "--out", out, "--seed", "3", "--max_epochs", "3",
                   "--keras_verbose", "0"]) == 0
    header, rows = _rows(out + "_predlocs.txt")
    assert header == "x,y,sampleID" and rows == []


def test_missing_calls_with_and_without_imputation(L, tmp_path):
    vcf, sd, na = _write_inputs(tmp_path, 50, 300, 5, missing=0.1)
    outs = []
    for extra in ([], ["--impute_missing"]):
        out = str(tmp_path / ("o" + str(len(extra))))
        assert L.main(["--vcf", vcf, "--sample_data", sd, "--out", out, "--seed", "11", "--max_epochs", "4",
                       "--keras_verbose", "0"] + extra) == 0
        outs.append(open(out + "_predlocs.txt").read())
        assert len(outs[-1].strip().split("\n")) == 6
    assert outs[0] != outs[1]  # imputed calls change the allele counts


def test_matrix_input_matches_vcf_input(L, tmp_path):
    """--matrix (counts 0/1/2 per sample and site, locator.py:200-227) gives the same run as the VCF it was
    derived from: same filter, same split, same model."""
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    vcf, sd, na = _write_inputs(tmp_path / "a", 40, 250, 4, seed=2)
    mat, sd2, _ = _write_inputs(tmp_path / "b", 40, 250, 4, seed=2, matrix=True)
    o1, o2 = str(tmp_path / "v"), str(tmp_path / "m")
    common = ["--seed", "5", "--max_epochs", "4", "--keras_verbose", "0"]
    assert L.main(["--vcf", vcf, "--sample_data", sd, "--out", o1] + common) == 0
    assert L.main(["--matrix", mat, "--sample_data", sd2, "--out", o2] + common) == 0
    assert open(o1 + "_predlocs.txt").read() == open(o2 + "_predlocs.txt").read()


def test_batch_size_above_256_is_refused_loudly(L, tmp_path):
    vcf, sd, na = _write_inputs(tmp_path, 40, 100, 4)
    with pytest.raises(SystemExit, match="batch_size"):  # validate_args: before any data is read
        L.main(["--vcf", vcf, "--sample_data", sd, "--out", str(tmp_path / "o"), "--seed", "1", "--max_epochs", "2",
                "--batch_size", "512", "--keras_verbose", "0"])


def test_replicate_drivers_on_degenerate_counts(L, tmp_path):
    """One bootstrap replicate with groups of four (a group of one), a jacknife that replaces no site, and
    --min_mac 1 (the count filter is skipped entirely, locator.py:270)."""
    vcf, sd, na = _write_inputs(tmp_path, 45, 260, 5, seed=4)
    common = ["--vcf", vcf, "--sample_data", sd, "--seed", "8", "--max_epochs", "3", "--keras_verbose", "0"]
    out = str(tmp_path / "b")
    assert L.main(common + ["--out", out, "--bootstrap", "--nboots", "1", "--replicates_per_gpu", "4"]) == 0
    assert len(_rows(out + "_bootFULL_predlocs.txt")[1]) == 5 and len(_rows(out + "_boot0_predlocs.txt")[1]) == 5
    out = str(tmp_path / "j")
    assert L.main(common + ["--out", out, "--jacknife", "--nboots", "2", "--jacknife_prop", "0"]) == 0
    full = open(out + "_bootFULL_predlocs.txt").read()
    assert open(out + "_boot0_predlocs.txt").read() == full and open(out + "_boot1_predlocs.txt").read() == full
    out1, out2 = str(tmp_path / "m1"), str(tmp_path / "m2")
    assert L.main(common + ["--out", out1, "--min_mac", "1"]) == 0
    assert L.main(common + ["--out", out2, "--min_mac", "2"]) == 0
    assert len(_rows(out1 + "_predlocs.txt")[1]) == 5


def test_weights_file_round_trip_and_load_params(L, tmp_path):
    """--keep_weights writes the best weights (Keras order); a fresh model that loads them predicts the same
    bytes.  --load_params re-runs with the stored arguments (locator.py:177-184)."""
    from locator_b200 import model as M
    import json

    vcf, sd, na = _write_inputs(tmp_path, 45, 260, 5, seed=6)
    out = str(tmp_path / "w")
    assert L.main(["--vcf", vcf, "--sample_data", sd, "--out", out, "--seed", "21", "--max_epochs", "4",
                   "--keras_verbose", "0", "--keep_weights"]) == 0
    first = open(out + "_predlocs.txt").read()
    with np.load(out + ".weights.npz") as z:
        ws = [z[k] for k in sorted(z.files)]
    K = ws[0].shape[0]
    assert ws[4].shape == (K, 256) and len(ws) == 4 + 2 * (10 + 2)
    m = M.LocatorModel(K, seed=1)
    m.load_weights(out + ".weights.npz")
    for a, b in zip(m.get_weights(), ws):
        assert np.array_equal(a, b)
    params = json.load(open(out + "_params.json"))
    params["out"] = str(tmp_path / "again")
    with open(tmp_path / "p.json", "w") as fh:
        json.dump(params, fh)
    assert L.main(["--load_params", str(tmp_path / "p.json"), "--out", "ignored"]) == 0
    assert open(str(tmp_path / "again") + "_predlocs.txt").read() == first
    # --load_weights: prediction from the stored weights, no training; same predictions, empty history
    out2 = str(tmp_path / "frozen")
    assert L.main(["--vcf", vcf, "--sample_data", sd, "--out", out2, "--seed", "21", "--keras_verbose", "0",
                   "--load_weights", out + ".weights.npz"]) == 0
    assert open(out2 + "_predlocs.txt").read() == first
    assert len(open(out2 + "_history.txt").read().splitlines()) == 1


def test_too_many_max_snps_fails_like_numpy(L, tmp_path):
    vcf, sd, na = _write_inputs(tmp_path, 40, 120, 4)
    with pytest.raises(ValueError):  # np.random.choice(range(K), max_SNPs, replace=False) with max_SNPs > K
        L.main(["--vcf", vcf, "--sample_data", sd, "--out", str(tmp_path / "o"), "--seed", "1", "--max_epochs", "2",
                "--max_SNPs", "100000", "--keras_verbose", "0"])
