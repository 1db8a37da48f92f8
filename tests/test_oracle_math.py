"""The shared IEEE-only math (vod_b200/csrc/vodb_math.h) evaluated by the CPU twin: known answers and accuracy."""
import ctypes

import numpy as np


def test_philox_known_answers(twin):
    """Philox-4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    L = twin.lib()

    def ph(c, k):
        out = (ctypes.c_uint32 * 4)()
        L.twin_philox(*c, *k, out)
        return [int(x) for x in out]

    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert ph([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert ph([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def _ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    return np.abs(got.astype(np.float64) - ref64) / np.spacing(np.abs(ref32)).astype(np.float64)


def test_logf_accuracy(twin):
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(-87, 88, 5000)).astype(np.float32)
    assert _ulp_err(twin.unary("logf", x), np.log(x.astype(np.float64))).max() < 1.0
    sub = np.array([1e-45, 1e-40, 1.1754942e-38], np.float32)
    assert np.allclose(twin.unary("logf", sub), np.log(sub.astype(np.float64)), rtol=1e-6)
    sp = twin.unary("logf", np.array([0.0, -0.0, -1.0, np.inf, np.nan, 1.0], np.float32))
    assert sp[0] == -np.inf and sp[1] == -np.inf and np.isnan(sp[2]) and sp[3] == np.inf and np.isnan(sp[4])
    assert sp[5] == 0.0


def test_expf_accuracy(twin):
    rng = np.random.default_rng(1)
    x = rng.uniform(-87, 88.7, 5000).astype(np.float32)
    assert _ulp_err(twin.unary("expf", x), np.exp(x.astype(np.float64))).max() < 1.0
    sp = twin.unary("expf", np.array([-np.inf, np.inf, np.nan, 0.0, -200.0, 89.0, -0.0], np.float32))
    assert sp[0] == 0.0 and sp[1] == np.inf and np.isnan(sp[2]) and sp[3] == 1.0 and sp[4] == 0.0
    assert sp[5] == np.inf and sp[6] == 1.0
    den = twin.unary("expf", np.array([-95.0, -100.0, -103.0], np.float32))
    assert np.allclose(den, np.exp(np.array([-95.0, -100.0, -103.0])), rtol=0.05)


def test_log1pf_accuracy(twin):
    rng = np.random.default_rng(2)
    x = np.concatenate([-np.exp(rng.uniform(-30, -1e-3, 3000)), np.exp(rng.uniform(-30, 10, 2000))]).astype(np.float32)
    ref = np.log1p(x.astype(np.float64))
    ok = np.isfinite(ref)
    assert _ulp_err(twin.unary("log1pf", x)[ok], ref[ok]).max() < 4.0
    sp = twin.unary("log1pf", np.array([-1.0, 0.0, -2.0, np.inf, 1e-20], np.float32))
    assert sp[0] == -np.inf and sp[1] == 0.0 and np.isnan(sp[2]) and sp[3] == np.inf and sp[4] == np.float32(1e-20)


def test_dtype_rounding_matches_torch(twin):
    import torch

    L = twin.lib()
    L.twin_f32_to_f16.restype = ctypes.c_uint16
    L.twin_f32_to_bf16.restype = ctypes.c_uint16
    L.twin_f32_to_f16.argtypes = [ctypes.c_float]
    L.twin_f32_to_bf16.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(3)
    v = (rng.standard_normal(4000) * np.exp(rng.uniform(-30, 12, 4000))).astype(np.float32)
    v = np.concatenate([v, np.array([0.0, -0.0, 65504.0, 65520.0, 1e-8, 6e-8, 5.96e-8, 2.98e-8, 1e38, -1e38], np.float32)])
    t = torch.from_numpy(v)
    bf = t.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    hf = t.to(torch.float16).view(torch.int16).numpy().view(np.uint16)
    mb = np.array([L.twin_f32_to_bf16(float(a)) for a in v], np.uint16)
    mh = np.array([L.twin_f32_to_f16(float(a)) for a in v], np.uint16)
    assert np.array_equal(mb, bf)
    assert np.array_equal(mh, hf)


def test_noise_is_exponential(twin):
    e = twin.exp1_noise(seed=7, offset=3, B=8, K=512)
    assert e.min() > 0 and np.isfinite(e).all()
    assert abs(e.mean() - 1.0) < 0.06 and abs(e.var() - 1.0) < 0.15
    e2 = twin.exp1_noise(seed=7, offset=3, B=8, K=512)
    assert np.array_equal(e, e2)
    assert not np.array_equal(e, twin.exp1_noise(seed=7, offset=4, B=8, K=512))


def test_synthetic_rows_statistics(twin):
    x = twin.synth_rows(1234, 0, 512, 768, dtype=0)
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert np.array_equal(x[10:20], twin.synth_rows(1234, 10, 10, 768, dtype=0))  # counter based: row-addressable
    xb = twin.synth_rows(1234, 0, 16, 768, dtype=1)
    assert np.abs(xb - x[:16]).max() <= 2 ** -8 * np.abs(x[:16]).max()
    xn = twin.synth_rows(1234, 0, 16, 768, dtype=0, unit_norm=True)
    assert np.allclose(np.linalg.norm(xn, axis=1), 1.0, atol=1e-5)
