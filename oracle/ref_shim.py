"""oracle/ref_shim.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Loads the reference's OWN sampling / merge / normalise code by file path from
`/root/reference` so that golden vectors can be generated from it and the CPU twin
can be pinned against it. The reference packages cannot be imported as packages in
this image (lightning, tensorstore, faiss ... are absent and `np.float_` is gone in
numpy 2), so the five files on the hot path are loaded individually under stub
package modules:

    src/vod_types/retrieval.py
    src/vod_dataloaders/core/numpy_ops.py
    src/vod_dataloaders/core/sample.py
    src/vod_dataloaders/core/merge.py
    src/vod_dataloaders/core/normalize.py

`/root/reference` exists only in the build container, never on the GPU box: nothing
that runs there (gpu tests, smoke, bench) imports this module. `available()` says
whether the tree is present.
"""
from __future__ import annotations

import importlib.util
import os
import pathlib
import sys
import types

REFERENCE_ROOT = pathlib.Path(os.environ.get("VOD_REFERENCE_ROOT", "/root/reference"))

_loaded: dict[str, types.ModuleType] = {}


def available() -> bool:
    return (REFERENCE_ROOT / "src" / "vod_dataloaders" / "core" / "sample.py").exists()


def _load(name: str, path: pathlib.Path) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    assert spec is not None and spec.loader is not None
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load() -> dict[str, types.ModuleType]:
    """Return {"retrieval", "numpy_ops", "sample", "merge", "normalize"} reference modules."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    import numpy as np

    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/vodb_numba_cache")  # the reference tree is read-only
    if not hasattr(np, "float_"):  # removed in numpy 2; numpy_ops.py:10-11 uses it in TypeVars only
        np.float_ = np.float64  # type: ignore[attr-defined]

    src = REFERENCE_ROOT / "src"
    # stub packages so that `import vod_types as vt` / `from vod_dataloaders.core import numpy_ops` resolve
    vt = types.ModuleType("vod_types")
    vt.__path__ = []  # type: ignore[attr-defined]
    sys.modules.setdefault("vod_types", vt)
    retrieval = _load("vod_types.retrieval", src / "vod_types" / "retrieval.py")
    for attr in ("RetrievalBatch", "RetrievalSample", "RetrievalTuple", "RetrievalData"):
        setattr(sys.modules["vod_types"], attr, getattr(retrieval, attr))

    vdl = types.ModuleType("vod_dataloaders")
    vdl.__path__ = []  # type: ignore[attr-defined]
    core = types.ModuleType("vod_dataloaders.core")
    core.__path__ = []  # type: ignore[attr-defined]
    sys.modules.setdefault("vod_dataloaders", vdl)
    sys.modules.setdefault("vod_dataloaders.core", core)

    numpy_ops = _load("vod_dataloaders.core.numpy_ops", src / "vod_dataloaders" / "core" / "numpy_ops.py")
    numpy_ops.CACHE_NUMBA_JIT = False  # read-only tree: no on-disk numba cache
    sys.modules["vod_dataloaders.core"].numpy_ops = numpy_ops  # type: ignore[attr-defined]
    sample = _load("vod_dataloaders.core.sample", src / "vod_dataloaders" / "core" / "sample.py")
    merge = _load("vod_dataloaders.core.merge", src / "vod_dataloaders" / "core" / "merge.py")
    normalize = _load("vod_dataloaders.core.normalize", src / "vod_dataloaders" / "core" / "normalize.py")
    _loaded.update(retrieval=retrieval, numpy_ops=numpy_ops, sample=sample, merge=merge, normalize=normalize)
    return _loaded
