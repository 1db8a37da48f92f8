"""Config surface of the B200 backend: what a VOD experiment config selects with `backend: "b200"`.

The reference describes every search engine by a pydantic factory config with a `backend` literal, an optional
"diff" with all-optional fields, `config + diff`, and a `fingerprint()`; `vod_search.factory` dispatches on the
config type and returns a `SearchMaster`:

    BaseSearchFactoryConfig / FaissFactoryConfig / FaissFactoryDiff   src/vod_configs/search.py:91-153
    FaissGpuConfig (devices [-1] = all, shard=True, add_batch_size)    src/vod_configs/search.py:48-89
    build_faiss_index(vectors, config=..., skip_setup, free_resources) -> FaissMaster   src/vod_search/factory.py:131-190
    _init_dense_search_engine (isinstance dispatch)                    src/vod_search/factory.py:240-271

`B200FactoryConfig` / `B200FactoryDiff` / `build_b200_search` are the same three pieces for this backend. There is
no index file and no cache directory: the "index" is the HBM store the master fills from the vectors when it is
entered, so `cache_dir` and `barrier_fn` are accepted for signature compatibility and unused.
"""
from __future__ import annotations

import hashlib
import json
import typing as typ

import pydantic

from .search import B200SearchMaster

B200_METRICS = {"inner_product": 0}  # the only metric of the hot path (faiss.METRIC_INNER_PRODUCT == 0)


class _Strict(pydantic.BaseModel):
    model_config = pydantic.ConfigDict(extra="forbid", frozen=False)


class B200FactoryDiff(_Strict):
    """Relative configuration (cf. FaissFactoryDiff, search.py:110-121): unset fields keep the base value."""

    backend: typ.Literal["b200"] = "b200"
    group_key: None | str = None
    section_id_key: None | str = None
    factory: None | str = None
    metric: None | int = None
    dtype: None | str = None
    mode: None | str = None
    devices: None | list[int] = None
    add_batch_size: None | int = None
    serve: None | bool = None


class B200FactoryConfig(_Strict):
    """Configures the building of a B200 search master (cf. FaissFactoryConfig, search.py:124-153)."""

    backend: typ.Literal["b200"] = "b200"
    subset_id_key: None | str = "subset_id"   # BaseSearchFactoryConfig fields: accepted, unused by a dense engine
    section_id_key: None | str = "id"
    factory: str = "Flat"                     # exact search only; anything else is rejected at build time
    metric: int = 0                           # inner product
    dtype: typ.Literal["float32", "bfloat16", "float16"] = "bfloat16"  # storage dtype in HBM
    mode: None | typ.Literal["auto", "exact", "tensor", "tensor2", "tensor3"] = None  # None = auto (fp32-exact results)
    devices: list[int] = [0]                  # [-1] = every visible GPU, rows sharded over them (FaissGpuConfig.devices)
    add_batch_size: int = 2**18               # FaissGpuConfig.add_batch_size
    serve: bool = True                        # Unix-socket endpoint for clients pickled into DataLoader workers

    @pydantic.field_validator("metric", mode="before")
    @classmethod
    def _validate_metric(cls, v: str | int) -> int:
        if isinstance(v, str):
            if v not in B200_METRICS:
                raise ValueError(f"metric `{v}` is not supported by the b200 backend (inner_product is)")
            return B200_METRICS[v]
        if int(v) != 0:
            raise ValueError("the b200 backend implements the inner-product metric only")
        return int(v)

    @pydantic.field_validator("devices", mode="before")
    @classmethod
    def _validate_devices(cls, v: None | list[int]) -> list[int]:
        if v is None or list(v) == [-1]:  # search.py:60-65
            from . import _lib

            return list(range(max(_lib.load().vodb_device_count(), 1)))
        return list(v)

    def __add__(self, diff: None | B200FactoryDiff) -> "B200FactoryConfig":
        if diff is None:
            return self
        updates = {k: v for k, v in diff if v is not None and k != "group_key"}
        return self.model_copy(update=updates)

    def fingerprint(self) -> str:
        """Stable hash of everything that changes the built index (cf. search.py:150-153: serving details excluded)."""
        payload = self.model_dump(exclude={"serve", "devices", "add_batch_size"})
        return hashlib.sha256(json.dumps(payload, sort_keys=True).encode()).hexdigest()[:16]


def build_b200_search(vectors: typ.Any, *, config: B200FactoryConfig | dict, cache_dir: typ.Any = None,  # noqa: ARG001
                      skip_setup: bool = False, barrier_fn: None | typ.Callable[[str], None] = None,
                      serve_on_gpu: bool = True, free_resources: bool = False) -> B200SearchMaster:  # noqa: ARG001
    """Drop-in for `vod_search.factory.build_faiss_index` (factory.py:131-190): returns the (not yet entered) master;
    `with master as m: client = m.get_client()` uploads the vectors and serves them."""
    if isinstance(config, dict):
        config = B200FactoryConfig(**config)
    if config.factory != "Flat":
        raise ValueError(f"only the exact `Flat` (IndexFlatIP) factory is supported, got `{config.factory}`")
    if barrier_fn is not None:
        barrier_fn(f"b200 build: `{config.fingerprint()}`")
    devices = config.devices
    return B200SearchMaster(vectors, dtype=config.dtype, device=devices[0], devices=devices if len(devices) > 1 else None,
                            mode=None if config.mode in (None, "auto") else config.mode,
                            add_batch_size=config.add_batch_size, skip_setup=skip_setup, free_resources=free_resources,
                            serve=config.serve)
