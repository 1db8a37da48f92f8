"""Reader (and minimal writer) for the embedding stores the reference's predict step leaves on disk.

The reference writes passage embeddings into a zarr-v2 array through TensorStore and reads them back lazily when
it builds an index:

    TensorStoreFactory.instantiate(path, shape, chunk_size=100, driver="zarr", dtype="float32")
        -> <path>/factory.json + zarr metadata {"dtype": "<f4", "shape": [N, D], "chunks": [100, D], "fill_value": "NaN"}
                                                         src/vod_tools/ts_factory/ts_factory.py:54-90
    _write_vectors_to_store(store, vectors, idx)          src/vod_ops/workflows/predict/compute.py:118-137
    TensorStoreFactoryLazyArray / slice_arrays_sequence   src/vod_types/lazy_array.py:100-172
    build_faiss_index(vectors=<lazy array>)               src/vod_search/faiss_search/build.py:51-73

TensorStore is not needed to read that layout: a zarr-v2 array is a `.zarray` JSON document plus one file per chunk
(`"i.j"`, C order, little endian, optionally compressed). `ZarrV2Array` exposes it with the slicing interface
`build_b200_index` expects (`len`, `[i]`, `[a:b]`), so a store written by the reference's predict step can be
ingested into the HBM corpus store directly:

    vectors = open_vectors("/path/to/store")              # the directory TensorStoreFactory.from_path reads
    master = B200SearchMaster(vectors, dtype="bfloat16")  # or ingest(store, vectors)

Chunk codecs: none, zlib / gzip (stdlib), and the blosc container TensorStore's zarr driver writes by default
(`{"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": -1}`) with the lz4 / zstd / zlib inner codecs and byte
shuffle, decoded here from the published blosc-1 frame layout with pyarrow's raw lz4 / zstd decompressors. No
TensorStore- or blosc-written file exists in this image to pin the blosc decoder against (neither library is
installed), so that part is checked against frames assembled by the tests from the same format description;
uncompressed and zlib chunks are plain numpy / stdlib.

`write_zarr_v2` writes the same layout uncompressed (fixtures, and an encoder-side sink that needs no TensorStore).
"""
from __future__ import annotations

import concurrent.futures
import json
import pathlib
import struct
import typing as typ
import zlib

import numpy as np

_BLOSC_CODECS = {0: "blosclz", 1: "lz4", 2: "snappy", 3: "zlib", 4: "zstd"}


class UnsupportedCodecError(RuntimeError):
    """The chunk compressor named in `.zarray` cannot be decoded in this environment."""


def _raw_decompress(codec: str, data: bytes, out_size: int) -> bytes:
    if codec == "zlib":
        return zlib.decompress(data)
    if codec in ("lz4", "zstd", "snappy"):
        import pyarrow as pa

        name = {"lz4": "lz4_raw", "zstd": "zstd", "snappy": "snappy"}[codec]
        if not pa.Codec.is_available(name):
            raise UnsupportedCodecError(f"pyarrow was built without the `{name}` codec")
        return pa.decompress(data, decompressed_size=out_size, codec=name, asbytes=True)
    raise UnsupportedCodecError(f"blosc inner codec `{codec}` is not supported (lz4, zstd, zlib, snappy are)")


def _unshuffle(block: bytes, typesize: int) -> bytes:
    """Inverse of blosc's byte shuffle: the block holds `typesize` planes of n bytes (byte t of every element), then
    the bytes that did not fill a whole element."""
    n = len(block) // typesize
    if typesize <= 1 or n == 0:
        return block
    body = np.frombuffer(block, np.uint8, count=n * typesize).reshape(typesize, n).T.tobytes()
    return body + block[n * typesize:]


def blosc_decode(buf: bytes) -> bytes:
    """Decode one blosc-1 frame.

    Layout: 16-byte header `version, versionlz, flags, typesize, nbytes u32, blocksize u32, cbytes u32`; then, unless
    the MEMCPYED flag (0x2) is set, `nblocks` int32 block offsets and the blocks. A block is one stream, or `typesize`
    streams of equal uncompressed size when the block was split; every stream is `int32 csize` + data, stored raw when
    csize equals its uncompressed size. flags: 0x1 byte shuffle, 0x4 bit shuffle, 0x10 "do not split", bits 5-7 the
    inner codec. Whether a block was split depends on the blosc version that wrote it, so both readings are tried
    and the one that consumes exactly the block's bytes wins.
    """
    if len(buf) < 16:
        raise ValueError("blosc frame shorter than its header")
    _version, _versionlz, flags, typesize = buf[0], buf[1], buf[2], buf[3]
    nbytes, blocksize, cbytes = struct.unpack_from("<III", buf, 4)
    if cbytes > len(buf):
        raise ValueError(f"blosc frame truncated: header says {cbytes} bytes, got {len(buf)}")
    if flags & 0x2:
        return bytes(buf[16:16 + nbytes])
    if flags & 0x4:
        raise UnsupportedCodecError("blosc bit-shuffle is not supported (TensorStore uses it for 1-byte types only)")
    codec = _BLOSC_CODECS.get((flags >> 5) & 0x7, "?")
    if nbytes == 0:
        return b""
    nblocks = -(-nbytes // blocksize)
    bstarts = struct.unpack_from(f"<{nblocks}i", buf, 16)
    out = bytearray()
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        end = bstarts[b + 1] if b + 1 < nblocks else cbytes
        candidates = [1] if (flags & 0x10) or typesize <= 1 or bsize % typesize else [typesize, 1]
        block = None
        for nsplits in candidates:
            pos, parts, ok = bstarts[b], [], True
            part_size = bsize // nsplits
            for _ in range(nsplits):
                if pos + 4 > end:
                    ok = False
                    break
                (csize,) = struct.unpack_from("<i", buf, pos)
                pos += 4
                if csize <= 0 or pos + csize > end:
                    ok = False
                    break
                raw = bytes(buf[pos:pos + csize])
                pos += csize
                try:
                    parts.append(raw if csize == part_size else _raw_decompress(codec, raw, part_size))
                except UnsupportedCodecError:
                    raise
                except Exception:  # noqa: BLE001 - wrong split hypothesis: the stream does not decode
                    ok = False
                    break
                if len(parts[-1]) != part_size:
                    ok = False
                    break
            # blocks are written back to back (in any order when blosc ran multi-threaded): the right reading ends
            # on a stream boundary that is another block's start or the end of the frame
            if ok and (pos == end or pos in bstarts or pos == cbytes):
                block = b"".join(parts)
                break
        if block is None:
            raise ValueError(f"blosc block {b} does not parse as 1 or {typesize} streams")
        out += _unshuffle(block, typesize) if flags & 0x1 else block
    if len(out) != nbytes:
        raise ValueError(f"blosc frame decoded to {len(out)} bytes, header says {nbytes}")
    return bytes(out)


def _decode_chunk(raw: bytes, compressor: dict | None, out_size: int) -> bytes:
    if compressor is None:
        return raw
    cid = compressor.get("id")
    if cid in ("zlib", "gzip"):
        return zlib.decompress(raw, wbits=47)  # auto-detects the zlib / gzip header
    if cid == "blosc":
        return blosc_decode(raw)
    if cid == "zstd":
        return _raw_decompress("zstd", raw, out_size)
    raise UnsupportedCodecError(f"zarr compressor `{cid}` is not supported (none, zlib, gzip, zstd, blosc are)")


class ZarrV2Array:
    """Read-only view of a 2-D zarr-v2 array on the local file system, sliceable by rows.

    `len(a)`, `a.shape`, `a.dtype`, `a[i]` (1-D row), `a[i:j]` and `a[[i, j, ...]]` (2-D blocks) — the access pattern of
    `vt.slice_arrays_sequence` / `build_faiss_index` (lazy_array.py:165-172, build.py:62-73). Chunks that were never
    written read as the fill value (NaN for the reference's stores), like TensorStore.
    """

    def __init__(self, path: str | pathlib.Path, *, threads: int = 8):
        self.path = pathlib.Path(path)
        meta_path = self.path / ".zarray"
        if not meta_path.exists():
            raise FileNotFoundError(f"no zarr-v2 array at `{self.path}` (`.zarray` is missing)")
        meta = json.loads(meta_path.read_text())
        if meta.get("zarr_format") != 2:
            raise ValueError(f"zarr_format {meta.get('zarr_format')} is not supported (2 is)")
        if meta.get("order", "C") != "C":
            raise ValueError("only C-order chunks are supported")
        if meta.get("filters"):
            raise UnsupportedCodecError(f"zarr filters are not supported: {meta['filters']}")
        self.dtype = np.dtype(meta["dtype"])
        self.shape = tuple(int(x) for x in meta["shape"])
        self.chunks = tuple(int(x) for x in meta["chunks"])
        if len(self.shape) != 2 or len(self.chunks) != 2:
            raise ValueError(f"expected a 2-D [rows, dim] array, got shape {self.shape}")
        self.compressor = meta.get("compressor")
        self.separator = meta.get("dimension_separator", ".")
        fill = meta.get("fill_value")
        self.fill_value = self.dtype.type({"NaN": np.nan, "Infinity": np.inf, "-Infinity": -np.inf}.get(fill, fill or 0))
        self._threads = max(1, int(threads))
        self._pool: concurrent.futures.ThreadPoolExecutor | None = None
        if self.compressor is not None and self.compressor.get("id") not in ("zlib", "gzip", "blosc", "zstd"):
            raise UnsupportedCodecError(f"zarr compressor `{self.compressor.get('id')}` is not supported")

    # pickled into workers like the reference's lazy array (lazy_array.py:121-128): only the path travels
    def __getstate__(self) -> dict:
        return {"path": str(self.path), "threads": self._threads}

    def __setstate__(self, state: dict) -> None:
        self.__init__(state["path"], threads=state["threads"])

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def ndim(self) -> int:
        return 2

    def _chunk(self, ci: int, cj: int) -> np.ndarray:
        """One decoded chunk as a [chunk_rows, chunk_cols] array (edge chunks are stored full size)."""
        f = self.path / f"{ci}{self.separator}{cj}"
        cr, cc = self.chunks
        if not f.exists():
            return np.full((cr, cc), self.fill_value, self.dtype)
        data = _decode_chunk(f.read_bytes(), self.compressor, cr * cc * self.dtype.itemsize)
        if len(data) != cr * cc * self.dtype.itemsize:
            raise ValueError(f"chunk `{f.name}` decodes to {len(data)} bytes, expected {cr * cc * self.dtype.itemsize}")
        return np.frombuffer(data, self.dtype).reshape(cr, cc)

    def _rows(self, start: int, stop: int) -> np.ndarray:
        n, d = self.shape
        cr, cc = self.chunks
        out = np.empty((max(stop - start, 0), d), self.dtype)
        if stop <= start:
            return out
        jobs = [(ci, cj) for ci in range(start // cr, (stop - 1) // cr + 1) for cj in range(-(-d // cc))]

        def load(job):
            ci, cj = job
            block = self._chunk(ci, cj)
            r0, r1 = max(start, ci * cr), min(stop, (ci + 1) * cr)
            c0, c1 = cj * cc, min(d, (cj + 1) * cc)
            out[r0 - start:r1 - start, c0:c1] = block[r0 - ci * cr:r1 - ci * cr, :c1 - c0]

        if len(jobs) > 1 and self._threads > 1:
            if self._pool is None:
                self._pool = concurrent.futures.ThreadPoolExecutor(self._threads, thread_name_prefix="vodb-zarr")
            list(self._pool.map(load, jobs))
        else:
            for job in jobs:
                load(job)
        return out

    def __getitem__(self, item: typ.Any) -> np.ndarray:
        n = self.shape[0]
        if isinstance(item, (int, np.integer)):
            i = int(item) + (n if item < 0 else 0)
            if not 0 <= i < n:
                raise IndexError(f"row {item} out of range for {n} rows")
            return self._rows(i, i + 1)[0]
        if isinstance(item, slice):
            start, stop, step = item.indices(n)
            block = self._rows(start, stop) if step > 0 else self._rows(stop + 1, start + 1)[::-1]
            return block[::abs(step)] if abs(step) != 1 else block
        idx = np.asarray(item, dtype=np.int64).reshape(-1)  # `store[idx]` as the predict loop addresses it
        if idx.size and np.all(np.diff(idx) == 1):
            return self._rows(int(idx[0]), int(idx[-1]) + 1)
        return np.stack([self[int(i)] for i in idx]) if idx.size else np.empty((0, self.shape[1]), self.dtype)


def open_vectors(path: str | pathlib.Path, **kwargs: typ.Any) -> ZarrV2Array:
    """Open the store directory the reference's predict step wrote: `TensorStoreFactory.from_path(path)` reads
    `<path>/factory.json` (ts_factory.py:92-101) whose kvstore path holds the zarr array; a bare zarr directory
    (`.zarray` inside) is accepted as well."""
    path = pathlib.Path(path)
    cfg = path / "factory.json"
    if cfg.exists():
        spec = json.loads(cfg.read_text())
        if spec.get("driver") != "zarr":
            raise UnsupportedCodecError(f"store driver `{spec.get('driver')}` is not supported (zarr is)")
        kv = spec.get("kvstore", {})
        if kv.get("driver") != "file":
            raise UnsupportedCodecError(f"kvstore driver `{kv.get('driver')}` is not supported (file is)")
        target = pathlib.Path(kv["path"])
        if not (target / ".zarray").exists() and (path / ".zarray").exists():
            target = path  # the store was moved after it was written: the recorded absolute path is stale
        return ZarrV2Array(target, **kwargs)
    return ZarrV2Array(path, **kwargs)


def write_zarr_v2(path: str | pathlib.Path, vectors: np.ndarray, *, chunk_size: int = 100,
                  compressor: str | None = None) -> pathlib.Path:
    """Write `vectors` [N, D] in the layout of `TensorStoreFactory.instantiate` (ts_factory.py:54-90): `factory.json`,
    `.zarray` with chunks [chunk_size, D] and fill value NaN, one file per chunk. `compressor`: None or "zlib"."""
    path = pathlib.Path(path)
    a = np.ascontiguousarray(vectors)
    if a.ndim != 2:
        raise ValueError("expected a [N, D] array")
    if a.dtype.str not in ("<f2", "<f4", "<f8"):
        raise ValueError(f"dtype {a.dtype} is not one of float16 / float32 / float64")
    path.mkdir(parents=True, exist_ok=True)
    n, d = a.shape
    comp = None if compressor is None else {"id": "zlib", "level": 1}
    meta = {"zarr_format": 2, "shape": [n, d], "chunks": [chunk_size, d], "dtype": a.dtype.str, "fill_value": "NaN",
            "order": "C", "filters": None, "compressor": comp, "dimension_separator": "."}
    (path / ".zarray").write_text(json.dumps(meta, indent=2))
    factory = {"driver": "zarr", "kvstore": {"driver": "file", "path": str(path.expanduser().absolute())},
               "metadata": {"dtype": a.dtype.str, "shape": [n, d], "chunks": [chunk_size, d], "fill_value": "NaN"}}
    (path / "factory.json").write_text(json.dumps(factory, indent=2))
    for ci in range(-(-n // chunk_size)):
        block = np.full((chunk_size, d), np.nan, a.dtype)
        rows = a[ci * chunk_size:(ci + 1) * chunk_size]
        block[:len(rows)] = rows
        data = block.tobytes()
        (path / f"{ci}.0").write_bytes(zlib.compress(data, 1) if comp else data)
    return path


def ingest(store: typ.Any, vectors: typ.Any, *, batch_rows: int = 1 << 16, row0: int = 0) -> int:
    """Stream a (lazy) [N, D] array into a `CorpusStore` / `MultiGpuStore`, block by block, reading block i+1 from
    disk while block i is uploaded and converted on the GPU. Returns the number of rows added. The per-period index
    refresh of the reference without the faiss index file in between."""
    n = len(vectors)
    if n == 0:
        return 0
    with concurrent.futures.ThreadPoolExecutor(1, thread_name_prefix="vodb-ingest") as reader:
        nxt = reader.submit(lambda a, b: np.ascontiguousarray(vectors[a:b]), 0, min(n, batch_rows))
        at = 0
        while at < n:
            block = nxt.result()
            end = at + len(block)
            if end < n:
                nxt = reader.submit(lambda a, b: np.ascontiguousarray(vectors[a:b]), end, min(n, end + batch_rows))
            store.add(block, row0=row0 + at)
            at = end
    return n
