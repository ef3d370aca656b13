"""Post-hoc summary of a set of replicate predictions (bootstrap / jacknife / windows runs).

Reference behaviour: /root/reference/locator_py/plot_locator.py:26-134 (and scripts/plot_locator.R:57-132):
collect every ``*predlocs*`` file of a folder, and per sample report the geographic centroid of its
predicted locations and the prediction with the highest Gaussian kernel density (bandwidth 0.2, the
density is evaluated at the sample's own predictions); with ``--error`` the distances of both summaries
to the known location (hypotenuse, or great-circle km with ``--longlat``) are printed.  Output:
``{out}_centroids.txt`` (tab-separated: sampleID x y kd_x kd_y gc_x gc_y).  Plotting is not part of this
build (no matplotlib in the image).  CPU-only host tool: the work is a few thousand points.
The four numeric functions are checked against the reference's own (run with scikit-learn's KernelDensity on
positional arrays, tests/golden/make_reference_vectors.py).  One deliberate difference: the reference's script
hands pandas Series to kdepred, whose ``xcoords[max_index]`` is then a *label* lookup that raises for most
samples and silently falls back to the mean; here the density peak is always the peak.

usage: python -m locator_b200.summarize --infile DIR --sample_data FILE --out STEM [--error] [--longlat]
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import pandas as pd

BANDWIDTH = 0.2


def kde_peak(x, y, bandwidth=BANDWIDTH):
    """The prediction with the highest kernel density among a sample's predictions (first one on ties);
    falls back to the mean when the density cannot be evaluated (plot_locator.py:26-37)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.size == 0 or not (np.all(np.isfinite(x)) and np.all(np.isfinite(y))):
        return float(np.mean(x)) if x.size else np.nan, float(np.mean(y)) if y.size else np.nan
    d2 = (x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2
    e = -d2 / (2.0 * bandwidth * bandwidth)
    mx = e.max(axis=1, keepdims=True)
    score = mx[:, 0] + np.log(np.exp(e - mx).sum(axis=1))  # log-density up to a constant
    i = int(np.argmax(score))
    return float(x[i]), float(y[i])


def centroid(x, y):
    return float(np.sum(x) / len(x)), float(np.sum(y) / len(y))


def distance_km(xpred, ypred, x, y):
    """The reference's haversine (degrees are passed straight to sin/cos there, too: plot_locator.py:46-50)."""
    dlon, dlat = xpred - x, ypred - y
    a = np.sin(dlat / 2) ** 2 + np.cos(y) * np.cos(ypred) * np.sin(dlon / 2) ** 2
    return 6373.0 * 2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a))


def load_predictions(infile):
    """One predlocs file, or every file of a folder whose name contains 'predlocs'."""
    if os.path.isdir(infile):
        files = sorted(os.path.join(infile, f) for f in os.listdir(infile) if "predlocs" in f)
    else:
        files = [infile]
    if not files:
        raise SystemExit(f"no predlocs files under {infile}")
    frames = [pd.read_csv(f) for f in files]
    return pd.concat(frames, ignore_index=True).rename(columns={"x": "xpred", "y": "ypred"})


def summarize(preds, sample_data, longlat=False):
    locs = pd.read_csv(sample_data, sep="\t") if isinstance(sample_data, str) else sample_data
    aeg = pd.merge(preds, locs, on="sampleID")
    rows, kd_d, gc_d = [], [], []
    for sid, grp in aeg.groupby("sampleID", sort=False):
        xs, ys = grp["xpred"].to_numpy(), grp["ypred"].to_numpy()
        x0, y0 = float(grp["x"].iloc[0]), float(grp["y"].iloc[0])
        kx, ky = kde_peak(xs, ys)
        gx, gy = centroid(xs, ys)
        rows.append((sid, x0, y0, kx, ky, gx, gy))
        dist = distance_km if longlat else (lambda a, b, c, d: float(np.hypot(a - c, b - d)))
        kd_d.append(dist(kx, ky, x0, y0))
        gc_d.append(dist(gx, gy, x0, y0))
    table = pd.DataFrame(rows, columns=["sampleID", "x", "y", "kd_x", "kd_y", "gc_x", "gc_y"])
    return table, np.asarray(kd_d, dtype=np.float64), np.asarray(gc_d, dtype=np.float64)


def build_parser():
    p = argparse.ArgumentParser(description="Summarise a set of locator predictions (centroids and kernel-density peaks).")
    p.add_argument("--infile", required=True, help="folder with predlocs files, or one predlocs file")
    p.add_argument("--sample_data", required=True, help="tab-separated sampleID / x / y table")
    p.add_argument("--out", required=True, help="output stem (writes {out}_centroids.txt)")
    p.add_argument("--error", default=False, action="store_true", help="print error summaries (needs known locations)")
    p.add_argument("--longlat", default=False, action="store_true", help="report errors in kilometres")
    p.add_argument("--silence", default=False, action="store_true")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    table, kd, gc = summarize(load_predictions(args.infile), args.sample_data, args.longlat)
    table.to_csv(args.out + "_centroids.txt", index=False, sep="\t")
    if args.error and not args.silence:
        ok = np.isfinite(kd)
        for name, d in (("kernel peak", kd[ok]), ("centroid", gc[ok])):
            print(f"mean {name} error = {np.mean(d)}")
            print(f"median {name} error = {np.median(d)}")
            print(f"90% CI for {name} error = {np.quantile(d, 0.05)} {np.quantile(d, 0.95)}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
