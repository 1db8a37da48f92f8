"""Host-side pieces of bench.py that both arms share (no GPU): the query generator, the bf16 rounding helper, the
config dict the driver compares between the arms, and the CPU arm end to end on a tiny corpus."""
import json
import os
import pathlib
import subprocess
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402


def test_round_to_store_is_round_to_nearest_even():
    import torch

    x = np.random.default_rng(0).standard_normal((257, 768), dtype=np.float32) * 37.0
    x[0, :4] = [1.00390625, 1.01171875, -1.00390625, 3.0e-39]        # ties to even (both ways), a denormal
    for name, td in (("bfloat16", torch.bfloat16), ("float16", torch.float16)):
        want = torch.from_numpy(x).to(td).to(torch.float32).numpy()
        assert np.array_equal(bench.round_to_store(np, x, name), want), name


def test_both_arms_draw_the_same_queries_and_config():
    a = bench.make_queries(np, 3, 64, "bfloat16")
    b = bench.make_queries(np, 3, 64, "bfloat16")
    assert np.array_equal(a, b) and a.shape == (3, 64, 768) and a.dtype == np.float32
    assert np.array_equal(bench.round_to_store(np, a, "bfloat16"), a)            # representable in the store dtype
    full = bench.make_queries(np, 3, 64, None)
    assert not np.array_equal(bench.round_to_store(np, full, "bfloat16"), full)  # really needs its correction terms
    assert np.array_equal(bench.round_to_store(np, full, "bfloat16"), a)         # same draw, rounded once
    args = bench.parse_args.__wrapped__() if hasattr(bench.parse_args, "__wrapped__") else None
    ns = type("A", (), {"rows": 10_000_000, "gpus": 8, "store_dtype": "bfloat16", "exchange": "p2p"})()
    assert bench.workload_config(ns) == bench.workload_config(ns) and "mode" not in bench.workload_config(ns)
    assert args is None or args.gpus == 1


def test_reference_arm_scans_every_row_and_uses_all_host_threads(tmp_path):
    """`--impl reference` under a torchrun-like environment (OMP_NUM_THREADS=1, RANK set): rank 0 prints one line
    whose time is a full scan (no extrapolation), BLAS threads are restored; other ranks exit 0 without output."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--rows", "200000"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["steps"] == 1
    assert line["cpu_baseline"]["blas_threads"] in (None, os.cpu_count())
    assert "200000 of 200000 rows" in line["cpu_baseline"]["sample"] and "scaled" not in line["cpu_baseline"]["sample"]
    assert line["e2e"]["value"] == line["value"] and line["config"]["rows"] == 200000
    env["RANK"] = "1"
    r1 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                        text=True, env=env, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
