"""Weights files: this build's ``.weights.npz``  <->  the reference's Keras ``.weights.h5``.

The reference checkpoints with ``keras.callbacks.ModelCheckpoint(..., save_weights_only=True)`` into
``{out}[_boot{b}].weights.h5`` and reloads it with ``model.load_weights`` (/root/reference/locator/locator.py:332-348,
:380,386).  HDF5 needs h5py, which this build does not depend on (it is not installable on the build image), so
``--keep_weights`` writes ``.weights.npz``: the arrays of ``model.get_weights()`` in Keras order

    [gamma, beta, moving_mean, moving_var, W1, b1, ..., W_L, b_L, Wo1, bo1, Wo2, bo2]     (kernels [in, out])

named ``w000, w001, ...``.  On a machine that has Keras (any backend) this module converts either way THROUGH Keras
itself -- it builds the reference's network (``load_network``, locator.py:311-327) and calls ``set_weights`` /
``save_weights`` or ``load_weights`` / ``get_weights`` -- so the ``.h5`` layout is whatever the installed Keras
writes and reads; nothing about the HDF5 layout is restated here.

    python -m locator_b200.keras_weights to-h5  run.weights.npz run.weights.h5
    python -m locator_b200.keras_weights to-npz run.weights.h5  run.weights.npz --nsnps 100000 [--nlayers 10 --width 256]

Without Keras the converter says so and exits with status 2; ``read_npz`` / ``write_npz`` / ``describe`` work everywhere.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def expected_shapes(nsnps, nlayers=10, width=256):
    """Shapes of model.get_weights() for the reference network, in Keras order."""
    K, H, L = int(nsnps), int(width), int(nlayers)
    shapes = [(K,)] * 4
    dims = [K] + [H] * L + [2, 2]
    for i in range(len(dims) - 1):
        shapes += [(dims[i], dims[i + 1]), (dims[i + 1],)]
    return shapes


def describe(weights):
    """(nsnps, nlayers, width) of a Keras-order weight list; raises ValueError when it is not the reference network."""
    ws = [np.asarray(w) for w in weights]
    if len(ws) < 4 + 2 * 4 or len(ws) % 2:
        raise ValueError(f"{len(ws)} arrays: not a Locator weight list")
    K = ws[0].shape[0]
    L = (len(ws) - 4) // 2 - 2
    H = ws[4].shape[1]
    want = expected_shapes(K, L, H)
    got = [tuple(w.shape) for w in ws]
    if got != want:
        bad = next(i for i, (g, w) in enumerate(zip(got, want)) if g != w)
        raise ValueError(f"array {bad} has shape {got[bad]}, the reference network has {want[bad]} there")
    return K, L, H


def read_npz(path):
    with np.load(path) as z:
        return [np.asarray(z[k], dtype=np.float32) for k in sorted(z.files)]


def write_npz(path, weights):
    with open(path, "wb") as fh:
        np.savez(fh, **{f"w{i:03d}": np.asarray(w, dtype=np.float32) for i, w in enumerate(weights)})


def _keras():
    try:
        import keras  # Keras 3
        return keras
    except Exception:
        try:
            from tensorflow import keras  # TF-bundled Keras 2
            return keras
        except Exception:
            return None


def build_reference_network(keras, nsnps, nlayers=10, width=256, dropout_prop=0.25):
    """The layer stack of load_network (locator.py:317-326), for weight (de)serialisation only."""
    layers = keras.layers
    model = keras.Sequential()
    model.add(keras.Input(shape=(int(nsnps),)))
    model.add(layers.BatchNormalization())
    for _ in range(int(np.floor(nlayers / 2))):
        model.add(layers.Dense(width, activation="elu"))
    model.add(layers.Dropout(dropout_prop))
    for _ in range(int(np.ceil(nlayers / 2))):
        model.add(layers.Dense(width, activation="elu"))
    model.add(layers.Dense(2))
    model.add(layers.Dense(2))
    return model


def npz_to_h5(npz_path, h5_path):
    keras = _keras()
    if keras is None:
        raise RuntimeError("Keras is not importable here: run the conversion where the reference itself runs")
    ws = read_npz(npz_path)
    K, L, H = describe(ws)
    model = build_reference_network(keras, K, L, H)
    model.set_weights(ws)
    model.save_weights(h5_path)


def h5_to_npz(h5_path, npz_path, nsnps, nlayers=10, width=256):
    keras = _keras()
    if keras is None:
        raise RuntimeError("Keras is not importable here: run the conversion where the reference itself runs")
    model = build_reference_network(keras, nsnps, nlayers, width)
    model.load_weights(h5_path)
    ws = model.get_weights()
    describe(ws)
    write_npz(npz_path, ws)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("to-h5")
    a.add_argument("npz")
    a.add_argument("h5")
    b = sub.add_parser("to-npz")
    b.add_argument("h5")
    b.add_argument("npz")
    b.add_argument("--nsnps", type=int, required=True)
    b.add_argument("--nlayers", type=int, default=10)
    b.add_argument("--width", type=int, default=256)
    c = sub.add_parser("describe")
    c.add_argument("npz")
    ns = ap.parse_args(argv)
    try:
        if ns.cmd == "to-h5":
            npz_to_h5(ns.npz, ns.h5)
        elif ns.cmd == "to-npz":
            h5_to_npz(ns.h5, ns.npz, ns.nsnps, ns.nlayers, ns.width)
        else:
            K, L, H = describe(read_npz(ns.npz))
            print(f"nsnps {K} nlayers {L} width {H}")
    except RuntimeError as exc:
        print(f"keras_weights: {exc}", file=sys.stderr)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main())
