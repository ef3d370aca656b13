#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE'S OWN CODE, executed in the build container.

`/root/reference/locator/locator.py` cannot be imported here (TensorFlow, scikit-allel, zarr are absent),
but its ingest / index functions only need numpy and pandas.  This script parses the reference source with
`ast`, compiles the function definitions it needs *from the source where it lies* (nothing is copied into
this repository) and runs them:

  sort_samples, normalize_locs, split_train_test      -- verbatim, on the reference's fixture data
  filter_snps, replace_md                              -- verbatim control flow (filter order, min-MAC rule,
        the (site, sample) order of the imputation draws, the max_SNPs draw) over `GA`, a minimal stand-in for
        the four scikit-allel calls they make (count_alleles / is_biallelic / to_allele_counts / is_missing,
        written from scikit-allel's documentation: plain counting)
  the bootstrap and jacknife loops of main()           -- the statements that draw from numpy's stream
        (reseed + site resample, locator.py:637-653; frequencies, site choice and binomial replacement,
        :713-727), located in main()'s AST and executed with the arrays the real run would hold

Outputs (committed): tests/golden/reference_vectors.json and reference_vectors.npz.  tests/test_oracle.py
checks the oracle against them; tests/test_gpu_cli.py checks the CUDA path against the same files.

Run in the build container only:   python tests/golden/make_reference_vectors.py

One deviation, stated: the jacknife frequency loop does `sum(ac[i, :])` on uint8 rows, which promotes to a
wide integer under the numpy the reference pins (< 1.25) and wraps at 255 under numpy >= 2 (this container).
The allele-count matrix is therefore handed to that loop as int64 -- the pinned-numpy behaviour.
"""
import ast
import contextlib
import copy
import hashlib
import io
import json
import os
import sys
import types

import numpy as np
import pandas as pd

REF_SRC = "/root/reference/locator/locator.py"
REF_DATA = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------
# stand-in for the scikit-allel objects the reference's functions touch
# ---------------------------------------------------------------------------------------------
class Counts(np.ndarray):
    """allel.AlleleCountsArray: [n_variants, n_alleles] counts of called alleles."""

    def is_biallelic(self):
        return np.asarray((np.asarray(self) > 0).sum(axis=1) == 2)


class GA:
    """allel.GenotypeArray over int8 [n_variants, n_samples, 2]; a negative allele is a missing allele."""

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.int8)

    @property
    def shape(self):
        return self.a.shape

    def __len__(self):
        return self.a.shape[0]

    def __getitem__(self, key):
        if isinstance(key, tuple) and isinstance(key[0], list):  # genotypes[ac_filter, :, :] with a list of bools
            key = (np.asarray(key[0], dtype=bool),) + key[1:]
        return GA(self.a[key])

    def count_alleles(self):
        m = max(int(self.a.max()) if self.a.size else 0, 0)
        flat = self.a.reshape(self.a.shape[0], -1)
        out = np.stack([(flat == k).sum(axis=1) for k in range(m + 1)], axis=1).astype(np.int32)
        return out.view(Counts)

    def to_allele_counts(self):
        m = max(int(self.a.max()) if self.a.size else 0, 0)
        return np.stack([(self.a == k).sum(axis=2) for k in range(m + 1)], axis=2).astype(np.uint8)

    def is_missing(self):
        return (self.a < 0).any(axis=2)


# ---------------------------------------------------------------------------------------------
# the reference's code, compiled from its source file
# ---------------------------------------------------------------------------------------------
def reference_namespace(args):
    tree = ast.parse(open(REF_SRC).read(), filename=REF_SRC)
    wanted = {"sort_samples", "normalize_locs", "split_train_test", "filter_snps", "replace_md", "predict_locs",
              "load_genotypes", "load_network", "load_callbacks", "train_network"}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {f.name for f in fns} == wanted
    from scipy import spatial

    ns = {"np": np, "pd": pd, "sys": sys, "copy": copy, "args": args, "tqdm": lambda it, *a, **k: it, "spatial": spatial}
    exec(compile(ast.Module(body=fns, type_ignores=[]), REF_SRC, "exec"), ns)
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    return ns, main


def _uses_numpy_stream(node):
    return "np.random" in ast.unparse(node)


def loop_statements(main, marker):
    """Statements of the `for boot in ...` loop of main() whose source contains `marker`, restricted to the
    ones that draw from numpy's stream or prepare what those draws write to (pg = deepcopy(predgen))."""
    loops = [n for n in ast.walk(main) if isinstance(n, ast.For) and isinstance(n.target, ast.Name)
             and n.target.id == "boot" and marker in ast.unparse(n)]
    assert len(loops) == 1, (marker, len(loops))
    keep = [s for s in loops[0].body if _uses_numpy_stream(s) or ast.unparse(s).startswith("pg = ")]
    return keep


def run(stmts, ns):
    exec(compile(ast.Module(body=stmts, type_ignores=[]), REF_SRC, "exec"), ns)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def fixture_genotypes():
    """GT int8 [nvar, N, 2] + sample names of the reference's example VCF, through pandas (an independent parser)."""
    df = pd.read_csv(os.path.join(REF_DATA, "test_genotypes.vcf.gz"), sep="\t", comment=None, skiprows=5, dtype=str,
                     compression="gzip")
    samples = [c for c in df.columns[9:] if not c.startswith("Unnamed")]
    body = df[samples].to_numpy(dtype=str)
    parts = np.char.partition(body, "|")
    return np.stack([parts[:, :, 0].astype(np.int8), parts[:, :, 2].astype(np.int8)], axis=2), np.array(samples)


def main():
    args = types.SimpleNamespace(sample_data=os.path.join(REF_DATA, "test_sample_data.txt"), train_split=0.9, min_mac=2,
                                 impute_missing=False, max_SNPs=None, nboots=2, jacknife_prop=0.05)
    ns, main_fn = reference_namespace(args)
    gt, samples = fixture_genotypes()
    vec, arrays = {"generated_by": "tests/golden/make_reference_vectors.py (reference functions executed from "
                                   "/root/reference/locator/locator.py)", "seed": 12345}, {}

    # ---- config 1 flow: seed -> sort_samples -> normalize_locs -> filter_snps -> split_train_test ----------
    np.random.seed(12345)
    sample_data, locs = quiet(ns["sort_samples"], samples, gt)
    meanlong, sdlong, meanlat, sdlat, nlocs = ns["normalize_locs"](locs)
    ac = quiet(ns["filter_snps"], GA(gt))
    train, test, traingen, testgen, trainlocs, testlocs, pred, predgen = ns["split_train_test"](ac, nlocs)
    vec["fixture"] = {
        "locs_sha256": sha(locs.astype(np.float64)), "norm": [float(meanlong), float(sdlong), float(meanlat), float(sdlat)],
        "normalized_locs_sha256": sha(nlocs.astype(np.float64)),
        "ac_shape": list(ac.shape), "ac_dtype": str(ac.dtype), "ac_sha256": sha(ac),
        "train": [int(v) for v in train], "test": [int(v) for v in test], "pred": [int(v) for v in pred],
        "traingen_sha256": sha(traingen.astype(np.uint8)), "testgen_sha256": sha(testgen.astype(np.uint8)),
        "predgen_sha256": sha(predgen.astype(np.uint8)),
        "trainlocs_sha256": sha(trainlocs.astype(np.float64)), "testlocs_sha256": sha(testlocs.astype(np.float64)),
    }
    state_after_split = np.random.get_state()

    # ---- bootstrap loop (two replicates), continuing the stream of the run above ------------------------------
    boot_stmts = loop_statements(main_fn, "starting bootstrap")
    src = [ast.unparse(s) for s in boot_stmts]
    assert any("np.random.seed(np.random.choice(range(int(1000000.0)), 1))" in s for s in src) and \
        any(s.startswith("site_order = np.random.choice") for s in src), src
    boots = []
    for boot in range(2):
        env = dict(ns, traingen2=traingen, boot=boot)
        before = np.random.get_state()
        # the reseed value itself: replay the draw the first statement is about to make
        reseed = int(np.random.choice(range(int(1e6)), 1)[0])
        np.random.set_state(before)
        run(boot_stmts, env)
        boots.append({"reseed": reseed, "site_order_prefix": [int(v) for v in env["site_order"][:16]],
                      "site_order_sha256": sha(env["site_order"].astype(np.int64))})
    vec["bootstrap"] = boots

    # ---- jacknife loop (two replicates) from the same point of the stream ---------------------------------------
    np.random.set_state(state_after_split)
    fl = [n for n in ast.walk(main_fn) if isinstance(n, ast.For) and "af.append(sum(ac[i, :])" in ast.unparse(n)]
    assert len(fl) == 1
    env = dict(ns, ac=ac.astype(np.int64), af=[])  # int64: see the module docstring
    run([fl[0]], env)
    af = np.array(env["af"])
    jk_stmts = loop_statements(main_fn, "sites_to_remove")
    src = [ast.unparse(s) for s in jk_stmts]
    assert src[0].startswith("pg = copy.deepcopy(predgen)") and src[1].startswith("sites_to_remove = np.random.choice") \
        and "np.random.binomial(2, af[i], pg.shape[0])" in src[2], src
    jks = []
    for boot in range(2):
        env = dict(ns, predgen=predgen, af=af, boot=boot)
        run(jk_stmts, env)
        jks.append({"sites_prefix": [int(v) for v in env["sites_to_remove"][:16]], "nsites": int(len(env["sites_to_remove"])),
                    "sites_sha256": sha(env["sites_to_remove"].astype(np.int64)), "pg_sha256": sha(env["pg"].astype(np.uint8))})
    vec["jacknife"] = {"af_sha256": sha(af.astype(np.float64)), "replicates": jks,
                       "next_uniform": float(np.random.random())}

    # ---- imputation + SNP subsample on a small cube with missing calls ---------------------------------------
    rng = np.random.default_rng(2024)
    nvar, N = 300, 40
    p = rng.uniform(0.05, 0.95, size=(nvar, 1, 1))
    small = (rng.uniform(size=(nvar, N, 2)) < p).astype(np.int8)
    small[rng.uniform(size=(nvar, N)) < 0.08] = -1          # whole calls missing
    small[rng.uniform(size=(nvar, N, 2)) < 0.01] = -1       # half-missing calls
    small[rng.uniform(size=(nvar, N, 2)) < 0.004] = 2       # a few third alleles
    small[0] = 0
    small[1] = -1
    args.impute_missing, args.max_SNPs, args.min_mac = True, 120, 3
    np.random.seed(777)
    ac_small = quiet(ns["filter_snps"], GA(small))
    vec["impute_subsample"] = {"seed": 777, "min_mac": 3, "max_SNPs": 120, "ac_shape": list(ac_small.shape),
                               "next_uniform": float(np.random.random())}
    arrays["small_gt"] = small
    arrays["small_ac"] = np.asarray(ac_small, dtype=np.uint8)
    args.impute_missing, args.max_SNPs, args.min_mac = False, None, 1
    ac_mac1 = quiet(ns["filter_snps"], GA(small))
    arrays["small_ac_min_mac_1"] = np.asarray(ac_mac1, dtype=np.uint8)

    # ---- predict_locs: output files and printed summary, from a stand-in model with fixed predictions -----------
    class StubModel:
        def __init__(self, table):
            self.table = table

        def predict(self, x):
            return self.table[id(x)]

    class StubHistory:
        history = {"loss": [1.25, 0.75, 0.6180339887498949], "val_loss": [1.5, 0.875, 0.7071067811865476],
                   "learning_rate": [0.0010000000474974513, 0.0010000000474974513, 0.0005000000237487257]}

    prng = np.random.default_rng(31)
    n_pred, n_val = 7, 9
    pg, tg = object(), object()
    p_pred = prng.normal(size=(n_pred, 2)).astype(np.float32)
    p_val = prng.normal(size=(n_val, 2)).astype(np.float32)
    vlocs = prng.normal(size=(n_val, 2))
    names = np.array([f"msp_{i:02d}" for i in range(30)])
    pidx = np.array([1, 4, 5, 11, 17, 23, 29])
    arrays.update(out_p_pred=p_pred, out_p_val=p_val, out_testlocs=vlocs, out_pred_idx=pidx)
    outdir = os.path.join(HERE, "ref_out")
    import shutil

    shutil.rmtree(outdir, ignore_errors=True)  # only what predict_locs writes below lives here
    os.makedirs(outdir)
    # meanlong, sdlong, meanlat, sdlat: numpy float64 scalars, as normalize_locs returns them (the de-normalisation
    # float32 * float64 + float64 is then float64 under every numpy version)
    norm = tuple(np.float64(v) for v in (25.07687431, 14.2187, 24.9312, 14.09656562))
    cases = {"plain": dict(bootstrap=False, jacknife=False, windows=False, boot=0),
             "boot": dict(bootstrap=True, jacknife=False, windows=False, boot=3),
             "bootfull": dict(bootstrap=False, jacknife=True, windows=False, boot="FULL"),
             "window": dict(bootstrap=False, jacknife=False, windows=True, boot=0)}
    printed = {}
    cwd = os.getcwd()
    os.chdir(outdir)
    try:
        for name, c in cases.items():
            args.out = name
            args.bootstrap, args.jacknife, args.windows = c["bootstrap"], c["jacknife"], c["windows"]
            args.window_start, args.window_size = "0", "250000"   # as they arrive from the command line
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                dists = ns["predict_locs"](StubModel({id(pg): p_pred, id(tg): p_val}), pg, norm[1], norm[0], norm[3], norm[2],
                                           vlocs, pidx, names, tg, StubHistory(), c["boot"])
            printed[name] = {"stdout": buf.getvalue(), "dists": [float(d) for d in dists]}
    finally:
        os.chdir(cwd)
    vec["predict_locs"] = {"norm": [float(v) for v in norm], "history": StubHistory.history, "cases": cases, "printed": printed,
                           "files": sorted(os.listdir(outdir))}

    # ---- the --windows loop of main() (:531-583): bounds, per-window ingest and split, output naming ---------------
    # Everything that is not ingest / index work is a recording stub; the loop body itself is the reference's.
    vdf = pd.read_csv(os.path.join(REF_DATA, "test_genotypes.vcf.gz"), sep="\t", skiprows=5, usecols=[1], dtype=np.int64,
                      compression="gzip")
    positions = vdf["POS"].to_numpy()
    wloops = [n for n in ast.walk(main_fn) if isinstance(n, ast.For) and "Processing window" in ast.unparse(n)]
    assert len(wloops) == 1
    records = []

    def record_predict(model, predgen, sdlong, meanlong, sdlat, meanlat, testlocs, pred, samples_, testgen, history):
        records.append({"i": int(wenv["i"]), "a": int(wenv["a"]), "b": int(wenv["b"]), "out": args.out,
                        "K": int(wenv["ac"].shape[0]), "ac_sha256": sha(np.asarray(wenv["ac"], dtype=np.uint8)),
                        "test": [int(v) for v in wenv["test"]], "train_sha256": sha(np.asarray(wenv["train"], dtype=np.int64)),
                        "traingen_sha256": sha(np.asarray(wenv["traingen"], dtype=np.uint8)),
                        "predgen_sha256": sha(np.asarray(predgen, dtype=np.uint8)),
                        "norm": [float(meanlong), float(sdlong), float(meanlat), float(sdlat)]})
        return []

    args.impute_missing, args.max_SNPs, args.min_mac, args.train_split = False, None, 2, 0.9
    args.out, args.plot_history, args.keep_weights, args.dropout_prop = "win", False, True, 0.25
    import time as _time

    wenv = dict(ns, allel=types.SimpleNamespace(GenotypeArray=GA), gt=gt, samples=samples, positions=positions,
                start=0, stop=int(np.max(positions)), size=625000, time=_time,
                load_network=lambda *a: None, load_callbacks=lambda *a: None, train_network=lambda *a: (None, None),
                predict_locs=record_predict, plot_history=lambda *a: None,
                subprocess=types.SimpleNamespace(run=lambda *a, **k: None))
    np.random.seed(777)
    with contextlib.redirect_stdout(io.StringIO()):
        ns["split_train_test"](ac, nlocs)   # main() splits the whole genome before the window loop (:514-516)
        run(wloops, wenv)
    vec["windows"] = {"seed": 777, "window_size": 625000, "stop": int(np.max(positions)), "records": records,
                      "next_uniform": float(np.random.random())}

    # ---- load_genotypes, --matrix branch (:200-227): count table -> haplotype pairs -> genotype cube ---------------
    mrng = np.random.default_rng(12)
    counts = mrng.integers(0, 3, size=(9, 25))
    mpath = os.path.join(HERE, "matrix_input.txt")
    with open(mpath, "w") as fh:
        fh.write("sampleID\t" + "\t".join(f"site{j}" for j in range(counts.shape[1])) + "\n")
        for i in range(counts.shape[0]):
            fh.write(f"ind_{i}\t" + "\t".join(str(int(c)) for c in counts[i]) + "\n")

    class Haps:  # allel.HaplotypeArray [n_variants, n_haplotypes]: consecutive haplotype pairs form a sample
        def __init__(self, h):
            self.h = np.asarray(h)

        def to_genotypes(self, ploidy):
            return GA(self.h.reshape(self.h.shape[0], self.h.shape[1] // ploidy, ploidy))

    args.zarr, args.vcf, args.matrix = None, None, mpath
    ns["allel"] = types.SimpleNamespace(HaplotypeArray=Haps, GenotypeArray=GA)  # the functions' globals are `ns`
    m_genotypes, m_samples = ns["load_genotypes"]()
    arrays["matrix_gt"] = m_genotypes.a
    vec["matrix"] = {"samples": [str(v) for v in m_samples], "shape": list(m_genotypes.a.shape)}

    # ---- post-hoc summary (locator_py/plot_locator.py:26-60): kernel-density peak, centroid, distances ----------
    from math import atan2, cos, radians, sin, sqrt
    from sklearn.neighbors import KernelDensity

    ptree = ast.parse(open("/root/reference/locator_py/plot_locator.py").read())
    pf = [n for n in ptree.body if isinstance(n, ast.FunctionDef) and n.name in ("kdepred", "centroid", "distance", "distance_km")]
    assert len(pf) == 4
    sns = {"np": np, "KernelDensity": KernelDensity, "sin": sin, "cos": cos, "sqrt": sqrt, "atan2": atan2, "radians": radians}
    exec(compile(ast.Module(body=pf, type_ignores=[]), "plot_locator.py", "exec"), sns)
    srng = np.random.default_rng(44)
    summ = []
    for case in range(6):
        n = int(srng.integers(5, 60))
        cx, cy = srng.uniform(-50, 50, 2)
        xs = cx + 0.3 * srng.normal(size=n)
        ys = cy + 0.3 * srng.normal(size=n)
        if case % 2:  # a second, smaller cluster far away pulls the centroid, not the density peak
            k = n // 4
            xs[:k] += 7.0
            ys[:k] -= 5.0
        tx, ty = cx + 0.1, cy - 0.2
        kd = sns["kdepred"](xs, ys)        # positional arrays: the behaviour the function documents
        gc = sns["centroid"](xs, ys)
        summ.append({"x": xs.tolist(), "y": ys.tolist(), "truth": [float(tx), float(ty)],
                     "kd": [float(kd[0]), float(kd[1])], "gc": [float(gc[0]), float(gc[1])],
                     "kd_dist": float(sns["distance"](kd[0], kd[1], tx, ty)), "gc_dist": float(sns["distance"](gc[0], gc[1], tx, ty)),
                     "kd_dist_km": float(sns["distance_km"](kd[0], kd[1], tx, ty))})
    vec["summarize"] = summ

    # ---- load_network / load_callbacks / train_network (:311-394) around RECORDING stand-ins for Keras ---------
    # No arithmetic is pinned here (Keras is absent): what is pinned is what the reference ASKS Keras for -- layer
    # sequence and arguments per nlayers / width / dropout, optimizer, the loss expression (evaluated with numpy
    # as the backend), callback settings and file names per driver, fit() arguments, the weights reload.
    class Rec:
        def __init__(self, kind, *a, **k):
            self.kind, self.a, self.k = kind, a, k

        def desc(self):
            d = {"kind": self.kind}
            if self.a:
                d["args"] = [x if isinstance(x, (int, float, str, bool, type(None))) else list(x) for x in self.a]
            d.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in self.k.items()})
            return d

    class Seq:
        def __init__(self):
            self.layers, self.compiled, self.fit_kwargs, self.loaded = [], None, None, []

        def add(self, layer):
            self.layers.append(layer.desc())

        def compile(self, optimizer, loss):
            self.compiled, self.loss = {"optimizer": optimizer}, loss

        def fit(self, x, y, **kw):
            self.fit_kwargs = {k: (v if isinstance(v, (int, float, str, bool, type(None))) else type(v).__name__)
                               for k, v in kw.items()}
            self.fit_xy = (x, y, kw.get("validation_data"), kw.get("callbacks"))
            return "HISTORY"

        def load_weights(self, path):
            self.loaded.append(path)

    def factory(kind):
        return lambda *a, **k: Rec(kind, *a, **k)

    keras = types.SimpleNamespace(
        Sequential=Seq,
        layers=types.SimpleNamespace(BatchNormalization=factory("BatchNormalization"), Dense=factory("Dense"),
                                     Dropout=factory("Dropout")),
        callbacks=types.SimpleNamespace(ModelCheckpoint=factory("ModelCheckpoint"), EarlyStopping=factory("EarlyStopping"),
                                        ReduceLROnPlateau=factory("ReduceLROnPlateau")))
    backend = types.ModuleType("tensorflow.keras.backend")
    backend.sqrt, backend.sum, backend.square = np.sqrt, np.sum, np.square
    fake_tf, fake_keras = types.ModuleType("tensorflow"), types.ModuleType("tensorflow.keras")
    fake_tf.keras, fake_keras.backend = fake_keras, backend
    sys.modules.update({"tensorflow": fake_tf, "tensorflow.keras": fake_keras, "tensorflow.keras.backend": backend})
    shell = []
    ns.update(tf=types.SimpleNamespace(keras=keras), time=_time,
              subprocess=types.SimpleNamespace(run=lambda cmd, **k: shell.append(cmd)))
    lrng = np.random.default_rng(8)
    yt, yp = lrng.normal(size=(6, 2)).astype(np.float32), lrng.normal(size=(6, 2)).astype(np.float32)
    nets = []
    for nlayers, width, drop in [(10, 256, 0.25), (2, 256, 0.25), (3, 64, 0.0), (7, 128, 0.5), (8, 32, 0.1)]:
        args.nlayers, args.width, args.dropout_prop = nlayers, width, drop
        m = ns["load_network"](np.zeros((4, 321), np.uint8), 0.99)   # the argument is ignored: args.dropout_prop is used
        nets.append({"nlayers": nlayers, "width": width, "dropout_prop": drop, "layers": m.layers,
                     "compile": m.compiled, "loss_of_fixed_arrays": [float(v) for v in m.loss(yt, yp)]})
    arrays.update(loss_y_true=yt, loss_y_pred=yp)
    cbs, trains = [], []
    for bootstrap, jacknife, boot, keep in [(False, False, 0, False), (True, False, 5, False), (False, True, "FULL", True)]:
        args.bootstrap, args.jacknife, args.out, args.keep_weights = bootstrap, jacknife, "o/run", keep
        args.patience, args.keras_verbose, args.max_epochs, args.batch_size = 100, 1, 5000, 32
        got = ns["load_callbacks"](boot)
        cbs.append({"bootstrap": bootstrap, "jacknife": jacknife, "boot": boot, "callbacks": [c.desc() for c in got]})
        m = Seq()
        del shell[:]
        with contextlib.redirect_stdout(io.StringIO()):
            hist, back = ns["train_network"](m, "TRAINGEN", "TRAINLOCS", "TESTLOCS_PLACEHOLDER", "TL", got, boot)
        trains.append({"bootstrap": bootstrap, "jacknife": jacknife, "boot": boot, "keep_weights": keep,
                       "fit_kwargs": m.fit_kwargs, "loaded": list(m.loaded), "shell": list(shell),
                       "returns_history_and_model": bool(hist == "HISTORY" and back is m)})
    args.patience = 37
    cbs.append({"bootstrap": False, "jacknife": True, "boot": "FULL", "patience": 37,
                "callbacks": [c.desc() for c in ns["load_callbacks"]("FULL")]})
    vec["keras_requests"] = {"networks": nets, "callbacks": cbs, "train_network": trains}
    for k in ("tensorflow", "tensorflow.keras", "tensorflow.keras.backend"):
        sys.modules.pop(k, None)

    # ---- the command line: the reference's own argparse definition and its params.json dump (:12-184) ----------
    import argparse

    tree = ast.parse(open(REF_SRC).read(), filename=REF_SRC)
    parser_stmts = [n for n in tree.body
                    if (isinstance(n, ast.Assign) and ast.unparse(n).startswith("parser = argparse.ArgumentParser"))
                    or (isinstance(n, ast.Expr) and ast.unparse(n).startswith("parser.add_argument("))]
    pns = {"argparse": argparse}
    run(parser_stmts, pns)
    argvs = [
        ["--vcf", "a.vcf.gz", "--sample_data", "s.txt", "--out", "o/run"],
        ["--zarr", "g.zarr", "--sample_data", "s.txt", "--out", "w", "--windows", "--window_size", "250000",
         "--window_start", "1000", "--window_stop", "9000000", "--seed", "12345"],
        ["--matrix", "m.txt", "--sample_data", "s.txt", "--out", "b", "--bootstrap", "--nboots", "20", "--batch_size", "16",
         "--max_epochs", "100", "--patience", "10", "--min_mac", "1", "--max_SNPs", "5000", "--impute_missing",
         "--dropout_prop", "0.5", "--nlayers", "8", "--width", "128", "--train_split", "0.8", "--gpu_number", "1",
         "--plot_history", "False", "--keep_weights", "--keras_verbose", "2"],
        ["--vcf", "a.vcf", "--sample_data", "s.txt", "--out", "j", "--jacknife", "--jacknife_prop", "0.1", "--nboots", "7",
         "--load_params", "old_params.json"],
    ]
    vec["cli"] = [{"argv": av, "params_json": json.dumps(pns["parser"].parse_args(av).__dict__, indent=2)} for av in argvs]
    vec["cli_help_flags"] = sorted(a.option_strings[0] for a in pns["parser"]._actions if a.option_strings)

    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **arrays)
    print("reference vectors written:", vec["fixture"]["ac_shape"], "fixture matrix;", len(vec["fixture"]["test"]),
          "validation samples; bootstrap reseeds", [b["reseed"] for b in boots])


if __name__ == "__main__":
    main()
