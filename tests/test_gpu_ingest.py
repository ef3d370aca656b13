"""GPU parity: ingest kernels (through the C ABI) vs the CPU oracle -- bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from locator_b200 import genotypes

    return genotypes


def _rand_gt(rng, nvar, N, miss=0.05, multi=0.02):
    p = rng.uniform(0, 1, size=(nvar, 1, 1)) ** 2
    gt = (rng.uniform(size=(nvar, N, 2)) < p).astype(np.int8)
    gt[rng.uniform(size=gt.shape) < multi] = 2
    gt[rng.uniform(size=gt.shape) < miss] = -1
    gt[0] = 0            # monomorphic
    gt[1] = -1           # all missing
    gt[2, :, :] = 1      # fixed alt
    return gt


def test_site_stats_and_pack_fixture(G, fixture_gt, golden_dir):
    import json
    from oracle import ingest_ref

    gt = fixture_gt["calldata/GT"]
    facts = json.load(open(os.path.join(golden_dir, "fixture_facts.json")))
    g, na, alt, miss, keep = G.site_stats(gt, min_mac=2)
    ac_ref, idx_ref = ingest_ref.filter_snps(gt, min_mac=2, return_index=True)
    cnt = ingest_ref.count_alleles(gt)
    assert np.array_equal(na.cpu().numpy(), (cnt > 0).sum(1))
    assert np.array_equal(alt.cpu().numpy(), cnt[:, 1])
    assert int(miss.sum()) == facts["n_missing_calls"]
    idx = np.flatnonzero(keep.cpu().numpy())
    assert len(idx) == facts["n_kept_min_mac_2"]
    assert np.array_equal(idx, idx_ref)
    packed = G.pack_sites(g, idx)
    counts = packed.to_counts().cpu().numpy()  # [N, K]
    assert np.array_equal(counts, ac_ref.T)


@pytest.mark.parametrize("nvar,N,min_mac", [(300, 37, 2), (1000, 129, 1), (64, 500, 3), (17, 5, 2), (700, 1000, 2),
                                             (333, 2500, 2), (50, 1026, 2), (40, 4099, 1), (9, 1, 1), (600, 264, 2)])
def test_site_stats_random(G, nvar, N, min_mac):
    from oracle import ingest_ref

    rng = np.random.default_rng(nvar * 7 + N)
    gt = _rand_gt(rng, nvar, N, multi=0.02 if N < 200 else 0.2 / N)  # keep some sites biallelic at large N
    g, na, alt, miss, keep = G.site_stats(gt, min_mac=min_mac)
    cnt = ingest_ref.count_alleles(gt)
    assert np.array_equal(na.cpu().numpy(), (cnt > 0).sum(1))
    assert np.array_equal(alt.cpu().numpy(), cnt[:, 1] if cnt.shape[1] > 1 else 0 * cnt[:, 0])
    assert np.array_equal(miss.cpu().numpy(), ingest_ref.is_missing(gt).sum(1))
    ac_ref, idx_ref = ingest_ref.filter_snps(gt, min_mac=min_mac, return_index=True)
    idx = np.flatnonzero(keep.cpu().numpy())
    assert np.array_equal(idx, idx_ref)
    packed = G.pack_sites(g, idx)
    assert np.array_equal(packed.to_counts().cpu().numpy(), ac_ref.T)
    # pad bits are zero
    w = packed.words.cpu().numpy().view(np.uint32)
    K = len(idx)
    if K % 16:
        assert np.all(w[:, K // 16] >> np.uint32(2 * (K % 16)) == 0)
    assert np.all(w[:, (K + 15) // 16:] == 0)


@pytest.mark.parametrize("offset", [0, 1, 2, 6, 8, 14])
def test_site_stats_and_pack_at_any_pointer_offset(G, offset):
    """The C ABI takes raw pointers: rows that start at any byte offset (a window slice, an odd address)
    go through the scalar head / tail of the scan and the narrower loads of the pack kernel."""
    from locator_b200._cabi import lib, check
    from oracle import ingest_ref

    rng = np.random.default_rng(offset)
    nvar, N = 257, 264
    gt = _rand_gt(rng, nvar, N, multi=0.05)
    gt[5, 7, 0] = 100  # a far-away allele index
    buf = torch.zeros(gt.size + 64, dtype=torch.int8, device="cuda")
    buf[offset:offset + gt.size] = torch.as_tensor(gt.reshape(-1)).cuda()
    ptr = buf.data_ptr() + offset
    na = torch.empty(nvar, dtype=torch.int32, device="cuda")
    alt, miss = torch.empty_like(na), torch.empty_like(na)
    keep = torch.empty(nvar, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    check(lib.loc_site_stats(ptr, nvar, N, 2, na.data_ptr(), alt.data_ptr(), miss.data_ptr(), keep.data_ptr(), st))
    cnt = ingest_ref.count_alleles(gt)
    assert np.array_equal(na.cpu().numpy(), (cnt > 0).sum(1))
    assert np.array_equal(alt.cpu().numpy(), cnt[:, 1])
    assert np.array_equal(miss.cpu().numpy(), ingest_ref.is_missing(gt).sum(1))
    ac_ref, idx_ref = ingest_ref.filter_snps(gt, min_mac=2, return_index=True)
    assert np.array_equal(np.flatnonzero(keep.cpu().numpy()), idx_ref)
    out = G.PackedGenotypes.empty(N, len(idx_ref))
    idx = torch.as_tensor(idx_ref.astype(np.int64)).cuda()
    rc = lib.loc_pack_sites(ptr, nvar, N, idx.data_ptr(), len(idx_ref), out.ptr, out.row_words, st)
    if offset % 2:
        assert rc != 0 and b"aligned" in lib.loc_last_error()  # calls are byte pairs: an odd address is refused
    else:
        check(rc)
        assert np.array_equal(out.to_counts().cpu().numpy(), ac_ref.T)


@pytest.mark.parametrize("n,K", [(1, 1), (3, 16), (45, 5830), (90, 100003), (7, 63)])
def test_pack_unpack_gather(G, n, K):
    rng = np.random.default_rng(n + K)
    counts = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
    p = G.PackedGenotypes.from_counts(counts)
    assert p.row_words % 4 == 0 and p.row_words * 16 >= K
    assert np.array_equal(p.to_counts().cpu().numpy(), counts)
    rows = rng.integers(0, n, size=2 * n + 1)
    assert np.array_equal(p.take_rows(rows).to_counts().cpu().numpy(), counts[rows])
    cols = rng.integers(0, K, size=K)  # bootstrap-style resample with replacement
    assert np.array_equal(p.take_cols(cols).to_counts().cpu().numpy(), counts[:, cols])
    sub = rng.permutation(K)[: max(1, K // 3)]  # max_SNPs-style subsample
    assert np.array_equal(p.take_cols(sub).to_counts().cpu().numpy(), counts[:, sub])


def test_replace_cols_matches_jacknife_oracle(G):
    from oracle import ingest_ref

    rng = np.random.default_rng(5)
    n, K = 50, 4000
    predgen = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
    af = rng.uniform(0.01, 0.99, size=K)
    np.random.seed(777)
    ref, sites = ingest_ref.jacknife_replace(predgen, af, 0.05)
    # same draws, vectorised per site in the returned order
    np.random.seed(777)
    sites2 = np.random.choice(K, int(K * 0.05), replace=False)
    vals = np.stack([np.random.binomial(2, af[i], n) for i in sites2]).astype(np.uint8)
    assert np.array_equal(sites, sites2)
    p = G.PackedGenotypes.from_counts(predgen)
    p.replace_cols(sites2, vals)
    assert np.array_equal(p.to_counts().cpu().numpy(), ref)


def test_patch_calls(G):
    rng = np.random.default_rng(9)
    n, K = 33, 777
    counts = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
    p = G.PackedGenotypes.from_counts(counts)
    m = 500
    flat = rng.permutation(n * K)[:m]
    samp, ks = flat // K, flat % K
    vals = rng.integers(0, 3, size=m, dtype=np.uint8)
    p.patch(ks, samp, vals)
    counts[samp, ks] = vals
    assert np.array_equal(p.to_counts().cpu().numpy(), counts)


def test_full_size_roundtrip_properties(G):
    """BASELINE config 3 shape (2,500 x 200k): size-independent properties on the device."""
    n, K = 2500, 200_000
    gen = torch.Generator(device="cuda").manual_seed(3)
    counts = torch.randint(0, 3, (n, K), dtype=torch.uint8, device="cuda", generator=gen)
    p = G.PackedGenotypes.from_counts(counts)
    assert torch.equal(p.to_counts(), counts)
    perm = torch.randperm(K, device="cuda", generator=gen)
    q = p.take_cols(perm)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(K, device="cuda")
    assert torch.equal(q.take_cols(inv).to_counts(), counts)  # gather by a permutation, then its inverse
    rows = torch.randperm(n, device="cuda", generator=gen)
    assert torch.equal(p.take_rows(rows).to_counts(), counts[rows])
    assert int(p.to_counts().sum(dtype=torch.int64)) == int(counts.sum(dtype=torch.int64))


def test_full_size_filter_and_pack_properties(G):
    """A cfg3-sized window of raw calls (100k sites x 2,500 samples, 0.5 GB): the filter / pack kernels against
    torch reductions of the same cube on the device."""
    nvar, N = 100_000, 2500
    gen = torch.Generator(device="cuda").manual_seed(11)
    p = torch.rand((nvar, 1, 1), device="cuda", generator=gen) ** 2
    gt = (torch.rand((nvar, N, 2), device="cuda", generator=gen) < p).to(torch.int8)
    gt[torch.rand((nvar, N, 2), device="cuda", generator=gen) < 0.01] = -1
    gt[torch.rand((nvar, N, 2), device="cuda", generator=gen) < 0.002] = 2
    g, na, alt, miss, keep = G.site_stats(gt, min_mac=2)
    assert torch.equal(alt.long(), (gt == 1).sum(dim=(1, 2)))
    assert torch.equal(miss.long(), (gt < 0).any(dim=2).sum(dim=1))
    seen = torch.stack([(gt == a).any(dim=2).any(dim=1) for a in (0, 1, 2)]).sum(dim=0)
    assert torch.equal(na.long(), seen)
    assert torch.equal(keep.bool(), (seen == 2) & (alt >= 2))
    idx = torch.nonzero(keep).flatten()
    packed = G.pack_sites(g, idx)
    assert torch.equal(packed.to_counts(), (gt[idx] == 1).sum(dim=2).to(torch.uint8).T.contiguous())


@pytest.mark.parametrize("nvar", [0, 1, 5, 2047, 2048, 2049, 70001, 600_000])
def test_compact_sites_matches_flatnonzero(G, nvar):
    """Device-side ordered compaction of the keep mask (prefix sum + scatter; filter_snps, locator.py:269,273)."""
    rng = np.random.default_rng(nvar + 1)
    keep = (rng.uniform(size=nvar) < 0.57).astype(np.uint8)
    if nvar > 10:
        keep[:3] = 1
        keep[-2:] = 1
    idx = G.compact_sites(torch.as_tensor(keep).cuda())
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), np.flatnonzero(keep))
    # all kept / none kept
    if nvar:
        assert np.array_equal(G.compact_sites(torch.ones(nvar, dtype=torch.uint8, device="cuda")).cpu().numpy(), np.arange(nvar))
        assert G.compact_sites(torch.zeros(nvar, dtype=torch.uint8, device="cuda")).numel() == 0


@pytest.mark.parametrize("nvar,N,miss", [(300, 37, 0.05), (64, 500, 0.3), (17, 5, 0.0), (5000, 129, 0.01), (40, 2500, 0.02)])
def test_missing_calls_in_reference_order(G, nvar, N, miss):
    """(site, sample) of the missing calls of the kept sites, row-major: np.nonzero(is_missing) of the filtered
    cube -- the order replace_md draws in (locator.py:255-261)."""
    from oracle import ingest_ref

    rng = np.random.default_rng(nvar * 3 + N)
    gt = _rand_gt(rng, nvar, N, miss=miss, multi=0.0)
    g, na, alt, nmiss, keep = G.site_stats(gt, min_mac=1)
    idx = G.compact_sites(keep)
    ks, samps = G.missing_calls(g, idx, nmiss)
    sub = gt[idx.cpu().numpy()]
    ek, es = np.nonzero(ingest_ref.is_missing(sub))
    assert np.array_equal(ks.cpu().numpy(), ek)
    assert np.array_equal(samps.cpu().numpy(), es)


@pytest.mark.parametrize("n,K", [(1, 1), (7, 15), (33, 16), (250, 1000), (1000, 5830), (2500, 40_001)])
def test_site_sums_match_numpy(G, n, K):
    """Per-SNP sums over all samples with wide integers (jacknife allele frequencies, locator.py:714-717)."""
    rng = np.random.default_rng(n + K)
    x = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
    p = G.PackedGenotypes.from_counts(x)
    sums = p.site_sums().cpu().numpy()
    assert sums.dtype == np.int64
    assert np.array_equal(sums, x.astype(np.int64).sum(axis=0))


@pytest.mark.parametrize("n,K", [(1, 1), (7, 33), (300, 5830), (700, 30001), (40, 900_001)])
def test_host_upload_pack_equals_device_pack(G, n, K):
    """from_counts on a HOST uint8 matrix (loc_upload_pack_counts: pinned double-buffered row blocks, packed
    while the next block is staged) gives the same words as packing a device copy (loc_pack_counts), across the
    staging-block boundary (16 MB) and for rows longer than a block's share."""
    import torch

    rng = np.random.default_rng(n + K)
    x = rng.integers(0, 3, size=(n, K), dtype=np.uint8)
    a = G.PackedGenotypes.from_counts(x)
    b = G.PackedGenotypes.from_counts(torch.as_tensor(x).cuda())
    assert a.row_words == b.row_words and torch.equal(a.words, b.words)
    assert np.array_equal(a.to_numpy(), x)
    # a non-contiguous view (the reference slices ac[:, idx] before fit)
    c = G.PackedGenotypes.from_counts(x[:, ::-1][:, ::-1][::2])
    assert np.array_equal(c.to_numpy(), x[::2])
