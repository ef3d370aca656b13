"""Blosc-LZ4 / vlen-utf8 zarr decoding (load_genotypes' zarr branch, /root/reference/locator/locator.py:187-194).

Pinning: tests/golden/zarr_blosc/{Albania,Guyana} are two chunks (data files, byte-for-byte) of the
Blosc-LZ4 + byte-shuffle zarr store the reference ships (locator_py/map.zarr, country outlines as
float64 [2, n] lon/lat).  No zarr / numcodecs / blosc package exists in this image, so the decoded
values are checked through what the data must be (size-exact LZ4 streams, finite coordinates inside
the country's bounding box, contiguous outline) -- "parity unpinned" for the byte values themselves.
The multi-block / split / stored-frame layouts are exercised by frames built in this file.
"""
import json
import os
import struct

import numpy as np
import pytest

from locator_b200 import io

HERE = os.path.dirname(os.path.abspath(__file__))
ZB = os.path.join(HERE, "golden", "zarr_blosc")


def _lz4_encode(data: bytes) -> bytes:
    """Greedy LZ4 block encoder (test-side only)."""
    n, out, i, anchor, table = len(data), bytearray(), 0, 0, {}

    def emit(lit, mlen, off):
        token_l = min(len(lit), 15)
        token_m = 0 if mlen is None else min(mlen - 4, 15)
        out.append((token_l << 4) | token_m)
        if token_l == 15:
            r = len(lit) - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(lit)
        if mlen is not None:
            out.extend(struct.pack("<H", off))
            if token_m == 15:
                r = mlen - 4 - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    while i + 4 <= n - 5:
        key = data[i:i + 4]
        j = table.get(key)
        table[key] = i
        if j is not None and i - j <= 65535:
            m = 4
            while i + m < n - 5 and data[j + m] == data[i + m]:
                m += 1
            emit(data[anchor:i], m, i - j)
            i += m
            anchor = i
        else:
            i += 1
    emit(data[anchor:], None, 0)
    return bytes(out)


def _blosc_frame(raw: bytes, typesize: int, blocksize: int, shuffle=True, dont_split=False, memcpy=False, codec="lz4"):
    """codec "zstd": the streams are Zstandard frames (written by pyarrow's codec); Blosc never splits those."""
    nbytes = len(raw)
    if codec == "zstd":
        import pyarrow as pa

        dont_split = True
        encode = lambda part: pa.Codec("zstd").compress(part, asbytes=True)  # noqa: E731
    else:
        encode = _lz4_encode
    flags = (1 if shuffle else 0) | (0x10 if dont_split else 0) | ((4 if codec == "zstd" else 1) << 5) | (0x2 if memcpy else 0)
    if memcpy:
        body = raw
        return bytes([2, 1, flags, typesize]) + struct.pack("<III", nbytes, blocksize, 16 + nbytes) + body
    nblocks = -(-nbytes // blocksize)
    streams, bstarts, pos = [], [], 16 + 4 * nblocks
    for b in range(nblocks):
        blk = raw[b * blocksize:(b + 1) * blocksize]
        leftover = len(blk) != blocksize
        if shuffle:
            ne = len(blk) // typesize
            a = np.frombuffer(blk[:ne * typesize], np.uint8).reshape(ne, typesize).T.tobytes()
            blk = a + blk[ne * typesize:]
        nsplits = typesize if (not dont_split and not leftover and typesize <= 16 and blocksize // typesize >= 128) else 1
        ne = len(blk) // nsplits
        bstarts.append(pos)
        for s in range(nsplits):
            part = blk[s * ne:(s + 1) * ne]
            enc = encode(part)
            if len(enc) >= len(part):
                enc = part  # stored split: cbytes == neblock
            streams.append(struct.pack("<I", len(enc)) + enc)
            pos += 4 + len(enc)
    body = b"".join(struct.pack("<I", x) for x in bstarts) + b"".join(streams)
    return bytes([2, 1, flags, typesize]) + struct.pack("<III", nbytes, blocksize, 16 + len(body)) + body


@pytest.mark.parametrize("country,lon,lat", [("Albania", (19.0, 21.2), (39.5, 42.8)), ("Guyana", (-61.5, -56.4), (1.1, 8.6))])
def test_reference_blosc_chunks_decode(country, lon, lat):
    a = io._zarr_array(ZB, country)
    meta = json.load(open(os.path.join(ZB, country, ".zarray")))
    assert meta["compressor"]["id"] == "blosc" and meta["compressor"]["cname"] == "lz4" and meta["compressor"]["shuffle"] == 1
    assert a.shape == tuple(meta["shape"]) and a.dtype == np.float64
    assert np.isfinite(a).all()
    assert lon[0] < a[0].min() and a[0].max() < lon[1]
    assert lat[0] < a[1].min() and a[1].max() < lat[1]
    # a border outline: consecutive vertices are close to each other
    assert np.median(np.hypot(np.diff(a[0]), np.diff(a[1]))) < 0.1


@pytest.mark.parametrize("typesize,blocksize,n,shuffle,dont_split", [
    (1, 4096, 10000, True, False),     # int8 GT: shuffle is a no-op for typesize 1, 3 blocks with a leftover
    (8, 2048, 2048 * 3 + 40, True, False),   # split into 8 streams, leftover block unsplit
    (4, 1024, 5000, True, True),       # dont_split flag
    (8, 4096, 4096, False, False),     # no shuffle, one block
    (2, 100, 1000, True, False),       # blocks too small to split
])
def test_blosc_layouts_round_trip(typesize, blocksize, n, shuffle, dont_split):
    rng = np.random.default_rng(typesize * 1000 + n)
    raw = (rng.integers(0, 3, n // 2).astype(np.uint8).tobytes() + bytes(n - n // 2))  # compressible + run of zeros
    frame = _blosc_frame(raw, typesize, blocksize, shuffle, dont_split)
    assert io._blosc_decompress(frame) == raw
    frame = _blosc_frame(raw, typesize, blocksize, memcpy=True)
    assert io._blosc_decompress(frame) == raw


def test_blosc_rejects_corrupt_and_unsupported():
    raw = bytes(range(256)) * 8
    frame = bytearray(_blosc_frame(raw, 1, 4096))
    bad = bytearray(frame)
    bad[2] = (bad[2] & 0x1F) | (3 << 5)  # zlib-inside-Blosc codec id: not supported
    with pytest.raises(RuntimeError, match="LZ4"):
        io._blosc_decompress(bytes(bad))
    bad[2] = (bad[2] & 0x1F) | (4 << 5)  # zstd codec id over LZ4 streams: a corrupt Zstandard stream
    with pytest.raises(RuntimeError, match="Zstandard"):
        io._blosc_decompress(bytes(bad))
    with pytest.raises(RuntimeError):
        io._blosc_decompress(bytes(frame[:40]))


def test_zarr_store_with_blosc_gt_and_vlen_samples(tmp_path, fixture_gt):
    """A store laid out like allel.vcf_to_zarr's: Blosc int8 calldata/GT, vlen-utf8 object samples, Blosc POS."""
    gt = fixture_gt["calldata/GT"][:700]
    nvar, N, _ = gt.shape
    samples = [f"msp_{i}" if i % 3 else f"sample-é{i}" for i in range(N)]
    pos = np.arange(nvar, dtype=np.int32) * 17 + 5
    root = tmp_path / "b.zarr"

    def put(name, meta, chunks):
        d = root / name
        d.mkdir(parents=True)
        (d / ".zarray").write_text(json.dumps(meta))
        for key, blob in chunks.items():
            (d / key).write_bytes(blob)

    comp = {"blocksize": 0, "clevel": 5, "cname": "lz4", "id": "blosc", "shuffle": 1}
    cv = 256
    chunks = {}
    for c in range(-(-nvar // cv)):
        blk = np.full((cv, N, 2), -1, np.int8)
        part = gt[c * cv:(c + 1) * cv]
        blk[:len(part)] = part
        chunks[f"{c}.0.0"] = _blosc_frame(blk.tobytes(), 1, 32768)
    put("calldata/GT", {"zarr_format": 2, "shape": [nvar, N, 2], "chunks": [cv, N, 2], "dtype": "|i1", "order": "C",
                        "compressor": comp, "fill_value": -1, "filters": None}, chunks)
    enc = struct.pack("<I", N) + b"".join(struct.pack("<I", len(s.encode())) + s.encode() for s in samples)
    put("samples", {"zarr_format": 2, "shape": [N], "chunks": [N], "dtype": "|O", "order": "C", "compressor": comp,
                    "fill_value": "", "filters": [{"id": "vlen-utf8"}]}, {"0": _blosc_frame(enc, 1, 1 << 16, shuffle=False)})
    put("variants/POS", {"zarr_format": 2, "shape": [nvar], "chunks": [cv], "dtype": "<i4", "order": "C",
                         "compressor": comp, "fill_value": 0, "filters": None},
        {str(c): _blosc_frame(np.pad(pos[c * cv:(c + 1) * cv], (0, cv - len(pos[c * cv:(c + 1) * cv]))).astype("<i4").tobytes(), 4, 1024)
         for c in range(-(-nvar // cv))})
    back = io.read_zarr(str(root))
    np.testing.assert_array_equal(back["calldata/GT"], gt)
    assert list(back["samples"]) == samples
    np.testing.assert_array_equal(back["variants/POS"], pos)


@pytest.mark.parametrize("codec", ["raw", "zlib", "blosc", "zstd", "blosc-zstd"])
def test_threaded_reads_of_a_chunk_grid_over_variants_and_samples(tmp_path, monkeypatch, codec):
    """allel.vcf_to_zarr chunks calldata/GT over variants AND samples ((65536, 64, 2) by default); the reader
    decodes the chunks of a row range from a thread pool.  Row ranges inside / across chunks, ragged edge
    chunks, a never-written chunk (fill value) and the single-thread path give the same array."""
    import zlib

    rng = np.random.default_rng(3)
    nvar, N, cv, cn = 5000, 300, 512, 64
    gt = rng.integers(-1, 3, size=(nvar, N, 2)).astype(np.int8)
    d = tmp_path / "g.zarr" / "calldata" / "GT"
    d.mkdir(parents=True)
    if "zstd" in codec:
        pa = pytest.importorskip("pyarrow")
    comp = {"raw": None, "zlib": {"id": "zlib", "level": 1}, "zstd": {"id": "zstd", "level": 1},
            "blosc": {"blocksize": 0, "clevel": 5, "cname": "lz4", "id": "blosc", "shuffle": 1},
            "blosc-zstd": {"blocksize": 0, "clevel": 1, "cname": "zstd", "id": "blosc", "shuffle": 1}}[codec]
    (d / ".zarray").write_text(json.dumps({"zarr_format": 2, "shape": [nvar, N, 2], "chunks": [cv, cn, 2], "dtype": "|i1",
                                           "order": "C", "compressor": comp, "fill_value": -1, "filters": None}))
    missing = (3, 2)
    for i in range(-(-nvar // cv)):
        for j in range(-(-N // cn)):
            if (i, j) == missing:
                continue
            blk = np.full((cv, cn, 2), -1, np.int8)
            part = gt[i * cv:(i + 1) * cv, j * cn:(j + 1) * cn]
            blk[:part.shape[0], :part.shape[1]] = part
            raw = blk.tobytes()
            blob = {"raw": lambda: raw, "zlib": lambda: zlib.compress(raw, 1), "blosc": lambda: _blosc_frame(raw, 1, 16384),
                    "zstd": lambda: pa.Codec("zstd").compress(raw, asbytes=True),
                    "blosc-zstd": lambda: _blosc_frame(raw, 1, 16384, codec="zstd")}[codec]()
            (d / f"{i}.{j}.0").write_bytes(blob)
    want = gt.copy()
    want[missing[0] * cv:(missing[0] + 1) * cv, missing[1] * cn:(missing[1] + 1) * cn] = -1
    rows = io.ZarrRows(str(tmp_path / "g.zarr"), "calldata/GT")
    assert rows.shape == (nvar, N, 2)
    monkeypatch.setenv("LOC_IO_THREADS", "4")
    monkeypatch.setattr(io, "_PARALLEL_MIN_BYTES", 0)
    for a, b in [(0, nvar), (100, 130), (500, 2100), (1536, 2048), (4990, nvar), (7, 7)]:
        np.testing.assert_array_equal(rows.rows(a, b).read(), want[a:b])
    monkeypatch.setenv("LOC_IO_THREADS", "1")
    np.testing.assert_array_equal(rows.rows(500, 2100).read(), want[500:2100])


def test_raw_whole_row_chunks_are_read_in_place_and_truncation_is_an_error(tmp_path):
    gt = np.random.default_rng(1).integers(-1, 2, size=(1000, 33, 2)).astype(np.int8)
    z = str(tmp_path / "r.zarr")
    io.write_zarr(z, gt, [f"s{i}" for i in range(33)], np.arange(1000), chunk_variants=300, compress=False)
    np.testing.assert_array_equal(io.ZarrRows(z, "calldata/GT", 250, 950).read(), gt[250:950])
    np.testing.assert_array_equal(io.read_zarr(z)["calldata/GT"], gt)
    fn = os.path.join(z, "calldata", "GT", "1.0.0")
    with open(fn, "rb") as fh:
        blob = fh.read()
    with open(fn, "wb") as fh:
        fh.write(blob[:len(blob) // 2])
    with pytest.raises(ValueError, match="truncated"):
        io.ZarrRows(z, "calldata/GT", 250, 950).read()


def test_lz4_decoder_fast_paths_at_buffer_edges():
    """The decoder copies short literal runs and matches with fixed-size moves when there is room and falls back
    to exact copies near the ends of the buffers; overlapping matches (runs, short periods) are replicated by
    doubling.  Periodic data of every small period, with and without a random tail, and low-entropy data of many
    lengths around the move sizes must round-trip for several block sizes."""
    rng = np.random.default_rng(5)
    for period in (1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 64):
        pat = bytes(rng.integers(0, 256, period, dtype=np.uint8))
        for tail in (0, 1, 5, 100):
            data = (pat * 1200)[:9000] + bytes(rng.integers(0, 256, tail, dtype=np.uint8))
            assert io._blosc_decompress(_blosc_frame(data, 1, 1 << 16, shuffle=False)) == data, (period, tail)
    for n in list(range(1, 40)) + [255, 256, 257, 4095, 4096, 4097, 20000]:
        data = bytes(rng.integers(0, 4, n, dtype=np.uint8))
        for bs in (64, 4096, 1 << 16):
            assert io._blosc_decompress(_blosc_frame(data, 1, bs, shuffle=False)) == data, (n, bs)
    # the shuffle flag on 1-byte items (what zarr writes for int8 calldata/GT) is the identity
    data = bytes(rng.integers(0, 3, 5000, dtype=np.uint8))
    assert io._blosc_decompress(_blosc_frame(data, 1, 2048, shuffle=True)) == data


def test_blosc_zstd_frames(tmp_path):
    """Stores written with zarr.Blosc(cname="zstd"): the library binds the system's libzstd at first use.  Frames
    with Zstandard streams (made by pyarrow's codec) of several item sizes / block sizes, with and without byte
    shuffle, and a zarr store whose calldata/GT chunks are such frames."""
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("zstd"):
        pytest.skip("pyarrow without zstd")
    rng = np.random.default_rng(9)
    for typesize, blocksize, n, shuffle in [(1, 4096, 20000, True), (1, 1 << 16, 70001, False), (4, 2048, 5000, True),
                                            (8, 4096, 1234, True), (2, 512, 513, True)]:
        raw = rng.integers(0, 3, n * typesize, dtype=np.uint8).tobytes()
        assert io._blosc_decompress(_blosc_frame(raw, typesize, blocksize, shuffle=shuffle, codec="zstd")) == raw
    gt = rng.integers(-1, 2, size=(900, 70, 2)).astype(np.int8)
    d = tmp_path / "z.zarr" / "calldata" / "GT"
    d.mkdir(parents=True)
    comp = {"blocksize": 0, "clevel": 1, "cname": "zstd", "id": "blosc", "shuffle": 1}
    (d / ".zarray").write_text(json.dumps({"zarr_format": 2, "shape": [900, 70, 2], "chunks": [256, 64, 2], "dtype": "|i1",
                                           "order": "C", "compressor": comp, "fill_value": 0, "filters": None}))
    for i in range(4):
        for j in range(2):
            blk = np.zeros((256, 64, 2), np.int8)
            part = gt[i * 256:(i + 1) * 256, j * 64:(j + 1) * 64]
            blk[:part.shape[0], :part.shape[1]] = part
            (d / f"{i}.{j}.0").write_bytes(_blosc_frame(blk.tobytes(), 1, 8192, codec="zstd"))
    np.testing.assert_array_equal(io.ZarrRows(str(tmp_path / "z.zarr"), "calldata/GT").read(), gt)
