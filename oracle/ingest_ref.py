"""Oracle: genotype ingest exactly as the reference does it (CPU, numpy).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows
``/root/reference/locator/locator.py``:

  load_genotypes   :187-228   (VCF text / --matrix; zarr is handled by the product reader)
  sort_samples     :231-247
  replace_md       :250-262
  filter_snps      :265-281
  normalize_locs   :284-292
  split_train_test :295-308
  bootstrap draw   :637-653
  jacknife draw    :714-727

scikit-allel is not installable here; its semantics are restated from its
published behaviour: ``count_alleles`` counts non-missing
allele indices, ``is_biallelic`` means exactly two alleles observed,
``to_allele_counts()[:, :, 1]`` counts index-1 alleles per call, ``is_missing``
means any allele of the call is negative.

All random draws go through numpy's *legacy global* stream (``np.random.*``),
the same calls in the same order as the reference, so indices are bit-exact for
a given ``--seed``.
"""
from __future__ import annotations

import gzip
import io

import numpy as np

# --------------------------------------------------------------------------
# VCF text -> GT int8 [nvar, nsamples, 2]   (allel.read_vcf, locator.py:197-199)
# --------------------------------------------------------------------------


def _open_text(path):
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return io.TextIOWrapper(gzip.open(path, "rb"), encoding="ascii", errors="replace")
    return open(path, "r", encoding="ascii", errors="replace")


def _parse_allele(tok):
    return -1 if tok in (".", "") else int(tok)


def read_vcf(path):
    """Slow, obviously-correct VCF reader: one python loop per call.

    Returns dict with 'calldata/GT' int8 [nvar, N, 2] (missing = -1, haploid
    second allele = -1), 'samples' (str array), 'variants/POS' (int64).
    """
    samples = None
    rows = []
    pos = []
    with _open_text(path) as fh:
        for line in fh:
            if line.startswith("##"):
                continue
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith("#CHROM"):
                samples = np.array(line.rstrip("\t").split("\t")[9:], dtype=str)
                continue
            if not line:
                continue
            f = line.rstrip("\t").split("\t")
            fmt = f[8].split(":")
            gi = fmt.index("GT")
            g = np.full((len(samples), 2), -1, dtype=np.int8)
            for s, field in enumerate(f[9 : 9 + len(samples)]):
                gt = field.split(":")[gi]
                alle = gt.replace("|", "/").split("/")
                g[s, 0] = _parse_allele(alle[0])
                if len(alle) > 1:
                    g[s, 1] = _parse_allele(alle[1])
            rows.append(g)
            pos.append(int(f[1]))
    gt = np.stack(rows) if rows else np.zeros((0, len(samples), 2), np.int8)
    return {"calldata/GT": gt, "samples": samples, "variants/POS": np.array(pos, dtype=np.int64)}


def matrix_to_gt(counts):
    """--matrix kludge, locator.py:204-227: allele count c -> haplotypes (c>=1, c>=2)."""
    counts = np.asarray(counts, dtype=np.int8)  # [nsamples, nsites]
    h1 = (counts >= 1).astype(np.int8)
    h2 = (counts >= 2).astype(np.int8)
    # counts outside 0/1/2 append nothing in the reference (ragged -> crash); reject.
    if np.any((counts < 0) | (counts > 2)):
        raise ValueError("matrix entries must be 0, 1 or 2")
    return np.stack([h1.T, h2.T], axis=2)  # [nsites, nsamples, 2]


# --------------------------------------------------------------------------
# scikit-allel semantics
# --------------------------------------------------------------------------


def count_alleles(gt):
    """allel GenotypeArray.count_alleles(): [nvar, max_allele+1] int32."""
    gt = np.asarray(gt)
    m = int(gt.max()) if gt.size else 0
    m = max(m, 0)
    out = np.zeros((gt.shape[0], m + 1), dtype=np.int32)
    flat = gt.reshape(gt.shape[0], int(np.prod(gt.shape[1:])))  # (-1 cannot be inferred for zero sites)
    for a in range(m + 1):
        out[:, a] = (flat == a).sum(axis=1)
    return out


def is_biallelic(ac):
    return (ac > 0).sum(axis=1) == 2


def alt_allele_counts(gt):
    """GenotypeArray.to_allele_counts()[:, :, 1] -> uint8 [nvar, N]."""
    return (np.asarray(gt) == 1).sum(axis=2).astype(np.uint8)


def is_missing(gt):
    return (np.asarray(gt) < 0).any(axis=2)


# --------------------------------------------------------------------------
# reference functions
# --------------------------------------------------------------------------


def replace_md(gt):
    """locator.py:250-262.  Scalar binomial draws in row-major (site, sample) order."""
    dc = count_alleles(gt)[:, 1]
    ac = alt_allele_counts(gt)
    missing = is_missing(gt)
    ninds = (~missing).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        af = dc / (2 * ninds)
    for i in range(ac.shape[0]):
        for j in range(ac.shape[1]):
            if missing[i, j]:
                ac[i, j] = np.random.binomial(2, af[i])
    return ac


def filter_snps(gt, min_mac=2, impute_missing=False, max_SNPs=None, return_index=False):
    """locator.py:265-281.  Returns ac uint8 [K, N] (and the kept site indices)."""
    gt = np.asarray(gt)
    idx = np.arange(gt.shape[0])
    tmp = count_alleles(gt)
    biallel = is_biallelic(tmp)
    gt = gt[biallel]
    idx = idx[biallel]
    if not min_mac == 1:
        derived = count_alleles(gt)
        derived = derived[:, 1] if derived.shape[1] > 1 else np.zeros(len(gt), np.int32)
        keep = derived >= min_mac
        gt = gt[keep]
        idx = idx[keep]
    if impute_missing:
        ac = replace_md(gt)
    else:
        ac = alt_allele_counts(gt)
    if max_SNPs is not None:
        sel = np.random.choice(range(ac.shape[0]), max_SNPs, replace=False)
        ac = ac[sel, :]
        idx = idx[sel]
    if return_index:
        return ac, idx
    return ac


def sort_samples(sample_ids, x, y, genotype_samples):
    """locator.py:231-247 without pandas: reindex metadata rows to genotype order.

    sample_ids/x/y are the columns of --sample_data; returns locs float64 [N, 2]
    in genotype-sample order, raising SystemExit like the reference when an ID
    is absent.
    """
    lookup = {}
    for i, s in enumerate(sample_ids):
        lookup[str(s)] = i  # duplicates: pandas would raise on reindex; last wins here
    locs = np.full((len(genotype_samples), 2), np.nan)
    for k, s in enumerate(genotype_samples):
        i = lookup.get(str(s))
        if i is None:
            print("sample ordering failed! Check that sample IDs match the VCF.")
            raise SystemExit
        locs[k, 0] = x[i]
        locs[k, 1] = y[i]
    return locs


def read_sample_data(path):
    """Tab-delimited table with header containing sampleID, x, y (NA = unknown)."""
    with open(path, "r") as fh:
        header = [h.strip().strip('"') for h in fh.readline().rstrip("\n").split("\t")]
        ci, cx, cy = header.index("sampleID"), header.index("x"), header.index("y")
        ids, xs, ys = [], [], []
        for line in fh:
            if not line.strip():
                continue
            f = [t.strip().strip('"') for t in line.rstrip("\n").split("\t")]
            ids.append(f[ci])
            xs.append(np.nan if f[cx] in ("NA", "", "NaN", "nan") else float(f[cx]))
            ys.append(np.nan if f[cy] in ("NA", "", "NaN", "nan") else float(f[cy]))
    return np.array(ids, dtype=str), np.array(xs), np.array(ys)


def normalize_locs(locs):
    """locator.py:284-292."""
    meanlong = np.nanmean(locs[:, 0])
    sdlong = np.nanstd(locs[:, 0])
    meanlat = np.nanmean(locs[:, 1])
    sdlat = np.nanstd(locs[:, 1])
    out = np.array([[(x[0] - meanlong) / sdlong, (x[1] - meanlat) / sdlat] for x in locs])
    return meanlong, sdlong, meanlat, sdlat, out


def split_train_test(ac, locs, train_split=0.9):
    """locator.py:295-308."""
    train = np.argwhere(~np.isnan(locs[:, 0]))
    train = np.array([x[0] for x in train])
    pred = np.array([x for x in range(len(locs)) if x not in set(train.tolist())])
    test = np.random.choice(train, round((1 - train_split) * len(train)), replace=False)
    tset = set(test.tolist())
    train = np.array([x for x in train if x not in tset])
    traingen = np.transpose(ac[:, train])
    trainlocs = locs[train]
    testgen = np.transpose(ac[:, test])
    testlocs = locs[test]
    predgen = np.transpose(ac[:, pred.astype(int)]) if len(pred) else np.zeros((0, ac.shape[0]), ac.dtype)
    return train, test, traingen, testgen, trainlocs, testlocs, pred, predgen


def bootstrap_site_order(nsites):
    """locator.py:637 + :648-650 -- reseed from the stream, then resample sites."""
    np.random.seed(np.random.choice(range(int(1e6)), 1))
    return np.random.choice(nsites, nsites, replace=True)


def jacknife_af(ac):
    """locator.py:714-717 with a wide integer sum (numpy>=2 would wrap uint8)."""
    return ac.astype(np.int64).sum(axis=1) / (ac.shape[1] * 2)


def jacknife_replace(predgen, af, prop):
    """locator.py:721-727 -- returns (copy with replaced columns, sites)."""
    pg = predgen.copy()
    sites = np.random.choice(pg.shape[1], int(pg.shape[1] * prop), replace=False)
    for i in sites:
        pg[:, i] = np.random.binomial(2, af[i], pg.shape[0])
    return pg, sites


def window_bounds(positions, start, stop, size):
    """locator.py:531-538: yields (i, a, b); the slice is gt[a:b] (b excluded)."""
    positions = np.asarray(positions)
    for i in np.arange(start, stop, size):
        mask = np.logical_and(positions >= i, positions < i + size)
        w = np.argwhere(mask)
        a = np.min(w)
        b = np.max(w)
        yield int(i), int(a), int(b)
