"""The C-ABI library loads and exports every symbol include/vodb.h declares (no GPU needed)."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def declared_symbols() -> list[str]:
    text = (ROOT / "include" / "vodb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vodb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for required in ("vodb_store_create", "vodb_store_add", "vodb_search", "vodb_merge_topk", "vodb_sample",
                     "vodb_store_destroy", "vodb_last_error"):
        assert required in syms


def test_library_exports_every_declared_symbol():
    from vod_b200 import _lib

    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libvodb.so does not export {name}"
        getattr(lib, name)
    assert lib.vodb_abi_version() == 1


def test_python_binding_covers_header():
    from vod_b200 import _lib

    assert _lib.exported_symbols() == declared_symbols()


def test_no_torch_types_in_signatures():
    text = (ROOT / "include" / "vodb.h").read_text()
    assert "torch" not in text.lower().replace("pytorch", "") or "no torch" in text.lower()
    assert 'extern "C"' in text


def test_argument_validation_without_gpu():
    """Pure host-side argument checks return VODB_EINVAL with a message (no compute, no GPU)."""
    from vod_b200 import _lib

    lib = _lib.load()
    assert lib.vodb_store_create(None, 0, 10, 8, 0, 0) == -1
    assert "out is NULL" in _lib.last_error()
    h = ctypes.c_void_p()
    assert lib.vodb_store_create(ctypes.byref(h), 0, -5, 8, 0, 0) == -1
    assert lib.vodb_store_create(ctypes.byref(h), 0, 10, 0, 0, 0) == -1
    assert lib.vodb_store_create(ctypes.byref(h), 0, 10, 8, 7, 0) == -1
    assert "dtype" in _lib.last_error()
    assert lib.vodb_search(None, None, 0, 0, 1, 10, 0, None, None, 0, None) == -1
    assert lib.vodb_sample(0, None, None, None, 1, 10, 5, 3, 1, 1.0, -1, 1, 0, 0, None, None, None, None, 0, None) == -1
    assert "k_positive" in _lib.last_error()
    assert lib.vodb_merge_topk(0, None, None, 0, 1, 1, 1, None, None, 0, None) == -1


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU fallback): no Python import of it, no
    #include / dlopen / path reference from the CUDA sources. The only mentions allowed are comments that name the
    twin as the shared specification of the sampler arithmetic."""
    for path in (ROOT / "vod_b200").rglob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{path} imports oracle/"
        assert "oracle" not in re.sub(r"#.*", "", src).replace("sample_twin", ""), f"{path} refers to oracle/ outside comments"
    for path in (ROOT / "vod_b200" / "csrc").iterdir():
        text = path.read_text()
        code = re.sub(r"//.*", "", re.sub(r"/\*.*?\*/", "", text, flags=re.S))  # strip comments
        assert "oracle" not in code, f"{path.name} refers to oracle/ in code"
        assert "twin" not in code.lower() or path.name == "vodb_math.h", f"{path.name} refers to the CPU twin in code"
    import vod_b200.build as vbuild

    assert not any("oracle" in s for s in vbuild.SOURCES + vbuild.NVCC_FLAGS), "libvodb.so is built from oracle sources"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vod_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libvodb.so")
    with pytest.raises(_lib.VodbUnavailableError):
        _lib.load()
