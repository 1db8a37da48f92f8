"""The zarr-v2 embedding-store reader (vod_b200/zarr_io.py) against layouts written independently of it.

The reference writes `[N, D]` float32 arrays with chunks `[100, D]` and fill value NaN through TensorStore
(src/vod_tools/ts_factory/ts_factory.py:54-90). The fixtures here are assembled with plain numpy / json / zlib and,
for the blosc container TensorStore uses by default, with a frame encoder written from the blosc-1 format
description (lz4 streams from pyarrow) — neither TensorStore nor blosc exists in this image, so the blosc decoder is
pinned to the format description only (said so in the module docstring)."""
import json
import pickle
import struct
import zlib

import numpy as np
import pytest

from vod_b200 import zarr_io


def _write_raw_zarr(path, a, chunk_rows, compressor=None, encode=lambda b: b, skip=()):
    """Independent writer: the zarr-v2 layout spelled out by hand (does not use zarr_io.write_zarr_v2)."""
    path.mkdir(parents=True)
    n, d = a.shape
    (path / ".zarray").write_text(json.dumps({
        "zarr_format": 2, "shape": [n, d], "chunks": [chunk_rows, d], "dtype": a.dtype.str, "fill_value": "NaN",
        "order": "C", "filters": None, "compressor": compressor}))
    for ci in range(-(-n // chunk_rows)):
        if ci in skip:
            continue
        block = np.full((chunk_rows, d), np.nan, a.dtype)
        rows = a[ci * chunk_rows:(ci + 1) * chunk_rows]
        block[:len(rows)] = rows
        (path / f"{ci}.0").write_bytes(encode(block.tobytes()))


def _blosc_frame(data: bytes, typesize: int, blocksize: int, split: bool, cname: str = "lz4") -> bytes:
    """blosc-1 frame: header, block offsets, per block 1 or `typesize` streams (int32 size + lz4 block), byte shuffle."""
    import pyarrow as pa

    codec_id, pa_name = {"lz4": (1, "lz4_raw"), "zstd": (4, "zstd")}[cname]
    nbytes = len(data)
    nblocks = -(-nbytes // blocksize)
    body, bstarts = b"", []
    base = 16 + 4 * nblocks
    for b in range(nblocks):
        blk = data[b * blocksize:(b + 1) * blocksize]
        n = len(blk) // typesize
        shuf = np.frombuffer(blk, np.uint8, count=n * typesize).reshape(n, typesize).T.tobytes() + blk[n * typesize:]
        nsplits = typesize if (split and len(blk) % typesize == 0 and len(blk) == blocksize) else 1
        bstarts.append(base + len(body))
        part = len(shuf) // nsplits
        for s in range(nsplits):
            piece = shuf[s * part:(s + 1) * part]
            comp = pa.compress(piece, codec=pa_name, asbytes=True)
            if len(comp) >= len(piece):
                comp = piece  # stored raw: csize == uncompressed size
            body += struct.pack("<i", len(comp)) + comp
    flags = 0x1 | (codec_id << 5) | (0 if split else 0x10)
    header = struct.pack("<BBBBIII", 2, 1, flags, typesize, nbytes, blocksize, base + len(body))
    return header + struct.pack(f"<{nblocks}i", *bstarts) + body


@pytest.fixture()
def vectors():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(1234, 48)).astype(np.float32)
    a[::7] = np.round(a[::7] * 4) / 4  # compressible rows too
    return a


@pytest.mark.parametrize("codec", ["none", "zlib", "gzip", "blosc-lz4-split", "blosc-lz4-nosplit", "blosc-zstd"])
def test_reads_what_an_independent_writer_wrote(tmp_path, vectors, codec):
    enc = {
        "none": (None, lambda b: b),
        "zlib": ({"id": "zlib", "level": 1}, lambda b: zlib.compress(b, 1)),
        "gzip": ({"id": "gzip", "level": 1}, lambda b: __import__("gzip").compress(b, 1)),
        "blosc-lz4-split": ({"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": -1, "blocksize": 0},
                            lambda b: _blosc_frame(b, 4, 4096, True)),
        "blosc-lz4-nosplit": ({"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1, "blocksize": 0},
                              lambda b: _blosc_frame(b, 4, 8000, False)),
        "blosc-zstd": ({"id": "blosc", "cname": "zstd", "clevel": 3, "shuffle": 1, "blocksize": 0},
                       lambda b: _blosc_frame(b, 4, 4096, True, "zstd")),
    }[codec]
    _write_raw_zarr(tmp_path / "store", vectors, 100, compressor=enc[0], encode=enc[1])
    arr = zarr_io.ZarrV2Array(tmp_path / "store")
    assert len(arr) == 1234 and arr.shape == (1234, 48) and arr.dtype == np.float32
    assert np.array_equal(arr[0:1234], vectors)
    assert np.array_equal(arr[95:305], vectors[95:305])         # crosses chunk boundaries
    assert np.array_equal(arr[1200:5000], vectors[1200:])       # clipped like lazy_array._slice_arr
    assert np.array_equal(arr[17], vectors[17]) and arr[17].ndim == 1
    assert np.array_equal(arr[-1], vectors[-1])
    assert np.array_equal(arr[[5, 6, 7]], vectors[5:8])
    assert np.array_equal(arr[[900, 3, 450]], vectors[[900, 3, 450]])
    again = pickle.loads(pickle.dumps(arr))                      # handed to workers like the reference's lazy array
    assert np.array_equal(again[100:200], vectors[100:200])


def test_missing_chunks_read_as_the_fill_value(tmp_path, vectors):
    _write_raw_zarr(tmp_path / "s", vectors, 100, skip={3})
    arr = zarr_io.ZarrV2Array(tmp_path / "s")
    got = arr[250:450]
    assert np.array_equal(got[:50], vectors[250:300]) and np.isnan(got[50:150]).all()
    assert np.array_equal(got[150:], vectors[400:450])


def test_factory_json_round_trip_and_float16(tmp_path, vectors):
    """`open_vectors` follows <path>/factory.json like TensorStoreFactory.from_path (ts_factory.py:92-101)."""
    half = vectors.astype(np.float16)
    p = zarr_io.write_zarr_v2(tmp_path / "emb", half, chunk_size=100)
    spec = json.loads((p / "factory.json").read_text())
    assert spec["driver"] == "zarr" and spec["kvstore"]["driver"] == "file"
    assert spec["metadata"] == {"dtype": "<f2", "shape": [1234, 48], "chunks": [100, 48], "fill_value": "NaN"}
    arr = zarr_io.open_vectors(p)
    assert arr.dtype == np.float16 and np.array_equal(arr[:], half)
    z = zarr_io.write_zarr_v2(tmp_path / "embz", vectors, chunk_size=64, compressor="zlib")
    assert np.array_equal(zarr_io.open_vectors(z)[:], vectors)


def test_unsupported_layouts_fail_loudly(tmp_path, vectors):
    _write_raw_zarr(tmp_path / "s", vectors, 100, compressor={"id": "lzma"})
    with pytest.raises(zarr_io.UnsupportedCodecError):
        zarr_io.ZarrV2Array(tmp_path / "s")
    with pytest.raises(FileNotFoundError):
        zarr_io.ZarrV2Array(tmp_path / "nothing")
    with pytest.raises(ValueError):
        zarr_io.blosc_decode(b"\x02\x01\x21\x04" + struct.pack("<III", 100, 100, 999))  # truncated frame


def test_ingest_streams_blocks_in_order(tmp_path, vectors):
    """`ingest` feeds a store-like sink block by block (the HBM store is exercised by the GPU tests)."""
    p = zarr_io.write_zarr_v2(tmp_path / "emb", vectors, chunk_size=100)

    class Sink:
        def __init__(self):
            self.blocks = []

        def add(self, rows, row0):
            self.blocks.append((row0, np.array(rows)))

    sink = Sink()
    assert zarr_io.ingest(sink, zarr_io.open_vectors(p), batch_rows=300) == 1234
    assert [r for r, _ in sink.blocks] == [0, 300, 600, 900, 1200]
    assert np.array_equal(np.concatenate([b for _, b in sink.blocks]), vectors)


def test_blosc_frames_of_random_shapes_round_trip():
    """Property check of the blosc-1 decoder against the test-side frame encoder: random payload sizes (including
    sizes that leave a short last block and bytes that do not fill an element), type sizes, block sizes, split /
    unsplit blocks, compressible and incompressible data."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(n=st.integers(1, 20_000), typesize=st.sampled_from([1, 2, 4, 8]), blocksize=st.sampled_from([512, 4096, 6000]),
               split=st.booleans(), cname=st.sampled_from(["lz4", "zstd"]), compressible=st.booleans(), seed=st.integers(0, 2**31))
    def check(n, typesize, blocksize, split, cname, compressible, seed):
        rng = np.random.default_rng(seed)
        data = (rng.integers(0, 4, size=n) if compressible else rng.integers(0, 256, size=n)).astype(np.uint8).tobytes()
        frame = _blosc_frame(data, typesize, blocksize, split and typesize > 1, cname)
        assert zarr_io.blosc_decode(frame) == data

    check()


def test_row_slicing_matches_numpy_for_random_requests(tmp_path):
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    rng = np.random.default_rng(3)
    a = rng.normal(size=(357, 20)).astype(np.float32)
    _write_raw_zarr(tmp_path / "s", a, 50, compressor={"id": "zlib", "level": 1}, encode=lambda b: zlib.compress(b, 1))
    arr = zarr_io.ZarrV2Array(tmp_path / "s", threads=2)

    @hyp.settings(max_examples=80, deadline=None)
    @hyp.given(start=st.integers(-400, 400), stop=st.integers(-400, 800), step=st.sampled_from([1, 2, 7, -1, -3]))
    def check(start, stop, step):
        assert np.array_equal(arr[start:stop:step], a[start:stop:step])

    check()
