#!/usr/bin/env python
"""bench.py — exact MIPS top-k throughput of the B200-native retrieval hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port), rank 0 only

Workload (BASELINE.json configs[1]): exact MIPS top-100 over a 10M x 768 bf16 corpus resident in HBM, 64-query
batches (HBM-bandwidth regime). One "step" = one search of a fresh 64-query batch. With N GPUs the same 10M-row
corpus is row-sharded over the N ranks (strong scaling): local top-k per shard, exchange fused into the final
select kernel (peer-mapped stores over NVLink), one merge kernel.

Output: ONE JSON line on rank 0 (see the task contract):
  value         queries/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e           queries/s through the reference-facing client call `B200SearchClient.search(vector=np.ndarray)` with
                HOST float32 queries, pinned H2D + D2H inside the timed region; at N>1 the client fronts the sharded
                corpus, every rank calls it in lockstep. `value`: queries holding bf16-representable values, what the
                reference hands over when the encoder runs in bf16-mixed as in its shipped recipes (the default mode
                then skips the empty correction terms on the device); `value_full_mantissa_f32_queries`: the same call
                with full-mantissa float32 queries, for which the default mode really scores three query terms
  roofline      scoring kernel: algorithmic bytes (rows*dim*2 per search and GPU) / CUDA-event kernel time
  cpu_baseline  the oracle port (`--impl reference` in a subprocess) over the same corpus on the host cores
  parity        (N>1, untimed) merged results: fused exchange == NCCL all-gather + merge, bit for bit, on every rank;
                float64 re-scoring of the returned ids; merged k-th score >= every shard's local k-th; k=100 and 1000
  target_config (N>1) BASELINE configs[2] / the north-star shape: 100M x 768 fp16 row-sharded, 64- and 8192-query
                batches, top-100 and top-1000, as fractions of the HBM / tensor-core roofline
  config1, config4_retrieve_and_sample, large_batch, latency: the other BASELINE configs on one GPU
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_ROWS = 10_000_000
DIM = 768
TOP_K = 100
Q_SMALL = 64
Q_LARGE = 8192
CORPUS_SEED = 1234
QUERY_SEED = 5678
TARGET_ROWS = 100_000_000   # BASELINE configs[2]
CPU_BLOCK_ROWS = 500_000


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=N_ROWS, help="override the corpus size (development only)")
    p.add_argument("--no-large", action="store_true", help="skip the 8192-query section")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    p.add_argument("--no-config4", action="store_true", help="skip the retrieve-and-sample section (profiling runs)")
    p.add_argument("--no-config1", action="store_true", help="skip the 100k x 768 fp32 section")
    p.add_argument("--no-target", action="store_true", help="skip the 100M-row target-config section (N>1)")
    p.add_argument("--no-parity", action="store_true", help="skip the untimed multi-GPU parity section (N>1)")
    p.add_argument("--target-rows", type=int, default=TARGET_ROWS)
    p.add_argument("--large-steps", type=int, default=5)
    p.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="cross-shard exchange (N>1)")
    p.add_argument("--top-k", type=int, default=TOP_K, help="results per query (BASELINE configs[2] uses 1000)")
    p.add_argument("--store-dtype", default="bfloat16", choices=["bfloat16", "float16"],
                   help="dtype of the HBM store (BASELINE configs[2] uses float16)")
    return p.parse_args()


def load_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def traffic_entry(key: str, algorithmic: float):
    """(bytes, provenance) of the DRAM traffic per search: the ncu `--set full` ratio of dram__bytes_read.sum +
    dram__bytes_write.sum to the algorithmic bytes (profiles/traffic.json, one capture per kernel) times the
    algorithmic bytes of this run — derived, not re-measured: ncu cannot run inside the benchmark."""
    path = ROOT / "profiles" / "traffic.json"
    if not path.exists():
        return None, None
    d = json.loads(path.read_text())
    if key not in d:
        return None, None
    return algorithmic * d[key], f"derived: algorithmic x {d[key]:.3f} (ncu --set full ratio, {d.get('source', 'profiles/')})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- synthetic queries: one numpy generator shared by both arms (same values on every rank and in the CPU arm) ----

def round_to_store(np, x, store_dtype: str):
    """float32 values rounded (RNE) to the store dtype and widened again: what an encoder running in bf16 / fp16 emits."""
    if store_dtype == "float16":
        return x.astype(np.float16).astype(np.float32)
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000  # bfloat16 round-to-nearest-even (finite values)
    return u.astype(np.uint32).view(np.float32)


def make_queries(np, n_batches: int, nq: int, store_dtype: str | None):
    """[n_batches, nq, DIM] float32 N(0,1) queries. `store_dtype` given: exactly representable in that dtype (the
    workload of `value`); None: full float32 mantissas (what a float32 encoder hands to the client)."""
    rng = np.random.default_rng([QUERY_SEED, nq])
    q = rng.standard_normal((n_batches, nq, DIM), dtype=np.float32)
    return q if store_dtype is None else round_to_store(np, q, store_dtype)


_WORKER_CODE = """
import pickle, sys, time
import numpy as np
from vod_b200.transport import RemoteSearch
address, authkey, n_batches, seed = pickle.loads(bytes.fromhex(sys.argv[1]))
v = np.random.default_rng(seed).standard_normal((32, 768), dtype=np.float32)
c = RemoteSearch(address, authkey)
c.search(v, 1000, None)                      # connect + warm
print("ready", flush=True)
sys.stdin.readline()                          # start signal
t0 = time.perf_counter()
for _ in range(n_batches):
    s, i = c.search(v, 1000, None)
assert s.shape == (32, 1000) and i.shape == (32, 1000) and (i >= 0).all()
print(time.perf_counter() - t0, flush=True)
"""


def worker_clients_section(store, n_workers: int = 8, n_batches: int = 12, repeats: int = 3):
    """configs[3] as the DataLoader sees it: `n_workers` worker PROCESSES, each holding its own Unix-socket
    connection to the GPU-owning master (like the reference's forkserver workers holding a FaissClient,
    src/vod_exps/train.py:15), each issuing 32-query top-1000 searches back to back. Requests that queue up during a
    scan share the next scan (vod_b200/transport.py ScanCoalescer); the uncoalesced line is the
    one-request-per-scan behaviour of the reference's single uvicorn worker (server.py:98). Each variant runs
    `repeats` times (median reported) with the server's per-scan GPU / host times."""
    import pickle

    from vod_b200.transport import SearchServer

    out = {"workers": n_workers, "queries_per_request": 32, "top_k": 1000, "requests_per_worker": n_batches,
           "unit": "queries/s", "repeats": repeats}
    root = str(pathlib.Path(__file__).resolve().parent)
    for label, coalesce in (("coalesced", True), ("one_scan_per_request", False)):
        runs = []
        for _ in range(repeats):
            scan_log: list[tuple[int, float]] = []

            def search_fn(v, k, mode, _log=scan_log):
                t0 = time.perf_counter()
                res = store.search(v, k, mode=mode or "tensor3")
                _log.append((len(v), (time.perf_counter() - t0) * 1e3))
                return res

            server = SearchServer(search_fn, lambda: True, coalesce=coalesce)
            server.start()
            procs = []
            try:
                for w in range(n_workers):
                    arg = pickle.dumps((server.address, server.authkey, n_batches, w)).hex()
                    procs.append(subprocess.Popen([sys.executable, "-c", _WORKER_CODE, arg], stdin=subprocess.PIPE,
                                                  stdout=subprocess.PIPE, text=True, cwd=root))

                def line_from(p, timeout_s):  # a stuck worker must not hang the bench
                    import select

                    if not select.select([p.stdout], [], [], timeout_s)[0]:
                        raise TimeoutError("search worker did not answer")
                    return p.stdout.readline()

                for p in procs:
                    if line_from(p, 120).strip() != "ready":
                        raise RuntimeError("search worker failed to start")
                scan_log.clear()
                t0 = time.perf_counter()
                for p in procs:
                    p.stdin.write("go\n")
                    p.stdin.flush()
                for p in procs:
                    float(line_from(p, 120))
                dt = time.perf_counter() - t0
                widths = sorted(n for n, _ in scan_log)
                ms = sorted(m for _, m in scan_log)
                runs.append({"qps": n_workers * n_batches * 32 / dt, "scans": len(scan_log),
                             "scan_ms_p50": ms[len(ms) // 2] if ms else None,
                             "queries_per_scan_p50": widths[len(widths) // 2] if widths else None,
                             "store_busy_frac": sum(ms) / (dt * 1e3) if ms else None})
            finally:
                for p in procs:
                    try:
                        p.stdin.close()
                        p.wait(timeout=30)
                    except Exception:  # noqa: BLE001
                        p.kill()
                server.stop()
        runs.sort(key=lambda r: r["qps"])
        mid = runs[len(runs) // 2]
        out[label] = mid["qps"]
        out[label + "_detail"] = {**mid, "qps_all_runs": [r["qps"] for r in runs]}
    return out


def workload_config(args):
    """Identical in both arms (the driver compares the dicts)."""
    short = {"bfloat16": "bf16", "float16": "fp16"}[args.store_dtype]
    return {
        "workload": (f"BASELINE configs[{1 if short == 'bf16' and TOP_K == 100 else 2}]: exact MIPS top-{TOP_K}, {args.rows} x {DIM} "
                     f"{short} corpus, {Q_SMALL}-query batches"),
        "rows": args.rows, "dim": DIM, "store_dtype": short, "queries_per_batch": Q_SMALL, "top_k": TOP_K,
        "n_gpus": args.gpus,
        "l2": "inputs larger than L2: the corpus shard streamed every step is >= 1.9 GB (L2 = 126 MB); fresh queries per step",
    }


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path for this workload = faiss IndexFlatIP.search, restated by oracle/flat_ip.py
# ---------------------------------------------------------------------------------------------------------------

def host_threads_env():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is a one-process job that should use the whole
    host. Must run before numpy is imported (OpenBLAS reads the variable when it is loaded)."""
    n = str(os.cpu_count() or 1)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = n


def run_reference(args):
    """Full scan of the same corpus with the same queries on the host cores; every step scans ALL rows (no
    extrapolation). The corpus is generated once (untimed), block by block with all host threads, and kept in RAM
    as float32 (what `index.add` stores, build.py:67-73); a step times `flat_ip.search` over every block plus the
    running merge. If the host cannot hold the corpus, the part that fits is scanned and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    host_threads_env()
    import concurrent.futures

    import numpy as np

    from oracle import flat_ip, twin

    cores = os.cpu_count() or 1
    blas_threads = None
    try:
        import threadpoolctl

        blas_threads = max((p.get("num_threads", 0) for p in threadpoolctl.threadpool_info() if p.get("user_api") == "blas"),
                           default=None)
    except Exception:  # noqa: BLE001
        pass
    dcode = {"bfloat16": 1, "float16": 2}[args.store_dtype]
    rows_total = args.rows
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:  # noqa: BLE001
        avail = 64 << 30
    rows_fit = int(avail * 0.7) // (DIM * 4)
    rows = min(rows_total, max(CPU_BLOCK_ROWS, rows_fit // CPU_BLOCK_ROWS * CPU_BLOCK_ROWS))
    t_gen = time.perf_counter()
    xb = np.empty((rows, DIM), np.float32)
    piece = 20_000

    def gen(r0):
        n = min(piece, rows - r0)
        xb[r0:r0 + n] = twin.synth_rows(CORPUS_SEED, r0, n, DIM, dtype=dcode)

    with concurrent.futures.ThreadPoolExecutor(cores) as ex:  # the C generator releases the GIL
        list(ex.map(gen, range(0, rows, piece)))
    t_gen = time.perf_counter() - t_gen
    queries = make_queries(np, args.warmup + args.steps, Q_SMALL, args.store_dtype)

    def scan(xq, threads):
        state = None
        for b0 in range(0, rows, CPU_BLOCK_ROWS):
            state = flat_ip.search(xb[b0:b0 + CPU_BLOCK_ROWS], xq, TOP_K, row_offset=b0, state=state, threads=threads)
        return state

    # collection threads: whichever of {1, half the cores} is faster on this host (decided on one block, untimed)
    probe = xb[:min(rows, CPU_BLOCK_ROWS)]
    pick, best = 1, None
    for th in (1, max(1, cores // 2)):
        flat_ip.search(probe, queries[0], TOP_K, threads=th)
        t0 = time.perf_counter()
        flat_ip.search(probe, queries[0], TOP_K, threads=th)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            pick, best = th, dt
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        scan(queries[step], pick)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    scale = rows_total / rows
    t_step = sum(times) / len(times) * scale
    value = Q_SMALL / t_step
    sample = (f"every step scans {rows} of {rows_total} rows x {DIM} fp32 (numpy/OpenBLAS sgemm in 65536-row blocks + "
              f"threshold collection + exact top-{TOP_K}, oracle/flat_ip.py); same queries as the GPU arm; BLAS threads "
              f"{blas_threads}, collection threads {pick}; corpus generated once in {t_gen:.1f} s (untimed)"
              + ("" if rows == rows_total else f"; host RAM holds only part of the corpus: time scaled x{scale:.2f}"))
    line = {
        "impl": "reference", "metric": f"mips_top{TOP_K}_queries_per_sec", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "blas_threads": blas_threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(args):
    """The CPU arm inside our own run (rank 0, N=1): the same `--impl reference` code path in a fresh process (its
    BLAS thread count must be set before numpy loads), fewer steps so that it stays a bounded sample."""
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "1",
           "--rows", str(args.rows), "--top-k", str(TOP_K), "--store-dtype", args.store_dtype]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------

def percentiles(xs):
    xs = sorted(xs)
    return {"p10": xs[len(xs) // 10], "p50": xs[len(xs) // 2], "p90": xs[(len(xs) * 9) // 10], "n": len(xs)}


def config1_section(np, torch, vod_b200, dev_index):
    """BASELINE configs[0]: IndexFlatIP exact top-100, 100k x 768 fp32, 256 queries — the reference's own CPU-runnable
    case (examples/search/faiss.py:27,41-58 with factory "Flat"): fresh queries per iteration, p10/p50/p90 of the
    search call. CPU = the oracle port on the host cores; GPU = `client.search` on a float32 store, default mode,
    host buffers in and out; and the two results must agree at the fp32-exact tolerance (north-star gate)."""
    import concurrent.futures

    from oracle import flat_ip, twin

    n, nq, k = 100_000, 256, 100
    xb = np.empty((n, DIM), np.float32)

    def gen(r0):
        xb[r0:r0 + 10_000] = twin.synth_rows(CORPUS_SEED, r0, 10_000, DIM, dtype=0)

    with concurrent.futures.ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        list(ex.map(gen, range(0, n, 10_000)))
    rng = np.random.default_rng([QUERY_SEED, 1])
    batches = rng.standard_normal((16, nq, DIM), dtype=np.float32)
    cpu_ms = []
    for i in range(2 + 6):
        t0 = time.perf_counter()
        ref_s, ref_i = flat_ip.search(xb, batches[i], k)
        if i >= 2:
            cpu_ms.append((time.perf_counter() - t0) * 1e3)
    out = {"workload": "BASELINE configs[0]: IndexFlatIP exact top-100, 100000 x 768 fp32, 256 queries, fresh queries per call",
           "cpu_ms": percentiles(cpu_ms), "cpu_cores": os.cpu_count(), "cpu_kind": "port (oracle/flat_ip.py)"}
    with vod_b200.B200SearchMaster(xb, dtype="float32", device=dev_index, serve=False) as master:
        client = master.get_client()
        for mode_name, mode in (("auto", None), ("exact_cuda_cores", "exact")):
            client.mode = mode
            gpu_ms = []
            for i in range(4 + 12):
                t0 = time.perf_counter()
                res = client.search(vector=batches[i], top_k=k)
                if i >= 4:
                    gpu_ms.append((time.perf_counter() - t0) * 1e3)
            res = client.search(vector=batches[7], top_k=k)
            rep = flat_ip.compare_topk(xb, batches[7], res.scores, res.indices, ref_s, ref_i)
            out[f"gpu_{mode_name}_ms"] = percentiles(gpu_ms)
            out[f"gpu_{mode_name}_parity"] = {"ok": rep["ok"], "max_score_rel_err": rep["max_score_rel_err"],
                                              "index_mismatches_all_near_ties": rep["n_index_mismatch"],
                                              "max_tie_gap": rep["max_tie_gap"]}
    out["speedup_p50_auto"] = out["cpu_ms"]["p50"] / out["gpu_auto_ms"]["p50"]
    return out


def sampler_baseline(np):
    """CPU baseline of the sampler beside the kernel: the C twin (bit-identical restatement of the reference's numba
    `_labeled_priority_sampling_2d_`, sample.py:323-352) on configs[3]'s shape, 32 x 1000 -> 8, one host thread."""
    from oracle import twin

    rng = np.random.default_rng(7)
    sc = np.sort(rng.normal(size=(32, 1000)).astype(np.float32) * 5, axis=1)[:, ::-1].copy()
    us = []
    for i in range(5 + 30):
        t0 = time.perf_counter()
        twin.sample(sc, None, k_positive=3, k_total=8, seed=42, offset=i)
        if i >= 5:
            us.append((time.perf_counter() - t0) * 1e6)
    return percentiles(us)


def main():
    global TOP_K
    args = parse_args()
    TOP_K = args.top_k
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import vod_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_true(ok: bool) -> bool:
        if world == 1:
            return ok
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed_section(corpus, nq: int, k: int, steps: int, warmup: int, sample_clocks: bool, store_dtype: str):
        """`steps` searches of fresh nq-query batches, device-resident in and out: CUDA events around the whole loop
        (max over ranks) + a second pass with per-kernel events for the scoring-kernel time."""
        queries = torch.from_numpy(make_queries(np, warmup + steps, nq, store_dtype)).to(dev)
        for i in range(warmup):
            corpus.search_device(queries[i], k, mode="tensor")
        torch.cuda.synchronize()
        assert not corpus.any_overflow(), "candidate list overflow during warm-up"
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            out = corpus.search_device(queries[warmup + i], k, mode="tensor")
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if sampler else None
        assert not corpus.any_overflow(), "candidate list overflow in the timed region (results invalid)"
        stats = corpus.store.stats()
        corpus.store.set_profiling(True)
        for i in range(steps):
            corpus.search_device(queries[warmup + i], k, mode="tensor")
        prof = corpus.store.profile()
        corpus.store.set_profiling(False)
        del queries
        return ms / steps, clocks, stats, prof, out

    # ---- corpus: row shard of the global synthetic corpus, generated on the device ----
    corpus = vod_b200.ShardedCorpus(args.rows, DIM, dtype=args.store_dtype, device=local_rank, rank=rank, world_size=world,
                                    exchange=args.exchange, max_queries=Q_LARGE, max_k=max(TOP_K, 100))
    corpus.fill_synthetic(CORPUS_SEED)
    torch.cuda.synchronize()
    shard_rows = corpus.hi - corpus.lo
    shard_bytes = shard_rows * DIM * 2

    ms_step, clocks, stats, prof, last_out = timed_section(corpus, Q_SMALL, TOP_K, args.steps, args.warmup, True, args.store_dtype)
    value = Q_SMALL / (ms_step * 1e-3)
    score_ms_per_search = max_over_ranks(prof["score_ms"] / args.steps)
    achieved_gbs = shard_bytes / (score_ms_per_search * 1e-3) / 1e9
    # kernels per search on this rank: prepare (query staging + list reset) + (score, select) per segment
    # (with N>1 the merge kernel is one more launch; the p2p exchange itself adds none, NCCL adds two collectives)
    launches_per_step = int(stats["launches"]) + (1 if world > 1 else 0)
    traffic, traffic_src = traffic_entry("score_q64_dram_over_algorithmic", shard_bytes) if world == 1 else (None, None)
    roofline = {
        "bound": "hbm",
        "kernel": "score_tc2_kernel<64,1,resident> (cta_group::2 pair, tcgen05 + TMA, query tile resident in shared memory, "
                  "fused top-k filter)",
        "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"],
        "algorithmic_bytes_per_search_per_gpu": shard_bytes, "score_kernel_ms_per_search": score_ms_per_search,
        "select_kernel_ms_per_search": prof["select_ms"] / args.steps, "score_launches_per_search": prof["score_launches"] / args.steps,
        "whole_step_frac": (shard_bytes / (ms_step * 1e-3) / 1e9) / peaks["hbm_gbs"],
    }

    # ---- e2e: the reference-facing client call with host buffers (pinned H2D + D2H inside the timed region) ----
    # One drop-in client at every N: `B200SearchMaster(store=...).get_client().search(vector=np.ndarray)`. At N>1 the
    # master fronts the sharded corpus (ShardedCorpus.search: fused exchange, all-rank overflow retry inside) and
    # every rank calls it in lockstep. Two query sets through the same default-mode call: float32 arrays holding
    # bf16-representable values (headline) and full-mantissa float32 values (reported beside it).
    def e2e_run(store_dtype_for_queries):
        q_host = torch.from_numpy(make_queries(np, args.warmup + args.steps, Q_SMALL, store_dtype_for_queries)).pin_memory()
        for i in range(args.warmup):
            client.search(vector=q_host[i].numpy(), top_k=TOP_K)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            res = client.search(vector=q_host[args.warmup + i].numpy(), top_k=TOP_K)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0), res

    master = vod_b200.B200SearchMaster(store=corpus.store if world == 1 else corpus, serve=False)
    master.__enter__()
    client = master.get_client()
    e2e_exact_s, e2e_res = e2e_run(args.store_dtype)
    e2e_full_s, _ = e2e_run(None)
    e2e = {"value": Q_SMALL * args.steps / e2e_exact_s, "unit": "queries/s", "h2d_bytes_per_step": Q_SMALL * DIM * 4,
           "d2h_bytes_per_step": Q_SMALL * TOP_K * 12, "ms_per_step": e2e_exact_s / args.steps * 1e3,
           "queries": "float32 numpy holding values a bf16 encoder produced (the reference's shipped recipes run bf16-mixed, "
                      "hydra/patch/arch/*.yaml, and widen bf16 vectors to float32, predict/compute.py:128-129): the default "
                      "mode finds the correction terms empty on the device and skips them",
           "value_full_mantissa_f32_queries": Q_SMALL * args.steps / e2e_full_s,
           "ms_per_step_full_mantissa_f32_queries": e2e_full_s / args.steps * 1e3,
           "full_mantissa_note": "float32 queries NOT representable in the store dtype (a float32 encoder): the default mode "
                                 "really scores three query terms; same call, same parity guarantee",
           # the round-1 key names, kept for readers of older lines
           "value_store_dtype_exact_queries": Q_SMALL * args.steps / e2e_exact_s,
           "ms_per_step_store_dtype_exact_queries": e2e_exact_s / args.steps * 1e3,
           "api": "B200SearchClient.search(vector=np.ndarray[64,768] f32, top_k) -> RetrievalBatch, default (auto) mode"
                  + ("" if world == 1 else "; the master fronts the row-sharded corpus, all ranks call in lockstep")}
    master.__exit__(None, None, None)

    # ---- per-call latency (search enqueue -> results ready on the device), p10 / p50 / p90 over fresh batches ----
    lat_q = torch.from_numpy(make_queries(np, 40, Q_SMALL, args.store_dtype)).to(dev)
    lat = []
    for i in range(40):
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        corpus.search_device(lat_q[i], TOP_K, mode="tensor")
        a1.record()
        torch.cuda.synchronize()
        if i >= 8:
            lat.append(max_over_ranks(a0.elapsed_time(a1)))
    latency = {"unit": "ms", **percentiles(lat),
               "what": "one 64-query search, device-resident in and out, CUDA events, max over ranks"}

    # ---- sustained rate: 250 searches back to back (the timed region above is 20 steps = 45 ms at N=1, before the 1 kW
    # power cap has stepped the clocks down). Events every 50 searches show the steps. ----
    sustained = None
    try:
        n_sus, ring = 250, 10
        sus_q = torch.from_numpy(make_queries(np, ring, Q_SMALL, args.store_dtype)).to(dev)
        torch.cuda.synchronize()
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_sus // 50 + 1)]
        marks[0].record()
        for i in range(n_sus):
            corpus.search_device(sus_q[i % ring], TOP_K, mode="tensor")
            if (i + 1) % 50 == 0:
                marks[(i + 1) // 50].record()
        barrier()
        assert not corpus.any_overflow()
        per50 = [max_over_ranks(marks[j].elapsed_time(marks[j + 1])) / 50 for j in range(len(marks) - 1)]
        sus_ms = sum(per50) / len(per50)
        sustained = {"steps": n_sus, "ms_per_step": sus_ms, "value": Q_SMALL / (sus_ms * 1e-3), "unit": "queries/s",
                     "ms_per_step_by_50": per50,
                     "whole_step_frac": (shard_bytes / (sus_ms * 1e-3) / 1e9) / peaks["hbm_gbs"],
                     "what": "same step as `value`, 250 in a row without a host wait; later blocks run under the power cap"}
        del sus_q
    except Exception as exc:  # noqa: BLE001  a side measurement
        sustained = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- multi-GPU parity (untimed): the merged result is right, on every rank ----
    parity = None
    if world > 1 and not args.no_parity:
        try:
            parity = parity_section(np, torch, dist, corpus, args, rank, world, dev, all_ranks_true)
        except Exception as exc:  # a failed check must show up in the line, not take the line down
            parity = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}

    # ---- BASELINE configs[3]: RealmCollate-style chain, 32 queries -> top-1000 -> priority sampling of 8 (rank 0) ----
    config4 = None
    if world == 1 and not args.no_config4:
        try:
            pipe = vod_b200.DenseRetrievalSampler(corpus.store, top_k=1000, total=8, max_pos_sections=3, mode="tensor")
            q4 = torch.from_numpy(make_queries(np, 30, 32, "bfloat16")).pin_memory()
            times, samp = [], []
            for i in range(30):
                t0 = time.perf_counter()
                pipe(q4[i], seed=42, offset=i)
                times.append((time.perf_counter() - t0) * 1e3)
            sc4 = torch.randn((32, 1000), device=dev)
            for i in range(30):
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                vod_b200.sample_device(sc4, None, k_positive=3, k_total=8, seed=42, offset=i)
                b1.record()
                torch.cuda.synchronize()
                samp.append(b0.elapsed_time(b1) * 1e3)
            # the reference-facing sampling call on host arrays (what RealmCollate calls after a hybrid merge)
            rb = vod_b200.RetrievalBatch(scores=sc4.cpu().numpy(), indices=np.tile(np.arange(1000, dtype=np.int64), (32, 1)))
            host_call = []
            for i in range(30):
                t0 = time.perf_counter()
                vod_b200.sample_search_results(search_results=rb, raw_scores={"dense": rb.scores}, total=8,
                                               max_pos_sections=3, seed=42, offset=i)
                host_call.append((time.perf_counter() - t0) * 1e3)
            host_call = sorted(host_call[5:])
            times, samp = sorted(times[5:]), sorted(samp[5:])
            config4 = {"workload": "32 queries -> exact top-1000 over the 10M x 768 bf16 shard -> labeled priority sampling of 8 "
                                   "(host queries in, [32,8] picks + log-weights out, one D2H)",
                       "chain_ms_p50": times[len(times) // 2], "chain_ms_p90": times[(len(times) * 9) // 10],
                       "sampler_kernel_us_p50": samp[len(samp) // 2],
                       "sample_search_results_host_call_ms_p50": host_call[len(host_call) // 2],
                       "sampler_cpu_twin_us": sampler_baseline(np),
                       "sampler_cpu_kind": "port: C twin of the reference's numba sampler, 1 host thread, 32 x 1000 -> 8"}
            try:
                config4["dataloader_workers"] = worker_clients_section(corpus.store)
            except Exception as exc:  # worker processes are a side measurement: never lose the chain numbers over them
                config4["dataloader_workers"] = {"error": f"{type(exc).__name__}: {exc}"}
        except Exception as exc:
            config4 = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- large-batch (tensor-core regime) section ----
    large = None
    if not args.no_large:
        try:
            ms_l, _, stats_l, prof_l, _ = timed_section(corpus, Q_LARGE, TOP_K, args.large_steps, 3, False, args.store_dtype)
            flops = 2.0 * Q_LARGE * shard_rows * DIM
            score_ms_l = max_over_ranks(prof_l["score_ms"] / args.large_steps)
            ach = flops / (score_ms_l * 1e-3) / 1e12
            tr_l, tr_l_src = traffic_entry("score_tc2_dram_over_algorithmic", shard_bytes) if world == 1 else (None, None)
            large = {
                "queries_per_batch": Q_LARGE, "value": Q_LARGE / (ms_l * 1e-3), "unit": "queries/s", "ms_per_step": ms_l,
                "steps": args.large_steps,
                "roofline": {"bound": "tensor",
                             "kernel": "score_tc_kernel<256,1> (1-CTA)" if os.environ.get("VODB_TC2", "1")[0] == "0"
                             else "score_tc2_kernel (cta_group::2 pair, 256 rows x 256 queries per MMA)",
                             "achieved": ach, "peak": peaks["bf16_tflops"],
                             "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"], "traffic": tr_l, "traffic_source": tr_l_src,
                             "peak_sustained": peaks["bf16_tflops_sustained"],
                             "frac_of_sustained": (ach / peaks["bf16_tflops_sustained"]) if peaks["bf16_tflops_sustained"] else None,
                             "score_kernel_ms_per_search": score_ms_l, "select_kernel_ms_per_search": prof_l["select_ms"] / args.large_steps,
                             "whole_step_frac": flops / (ms_l * 1e-3) / 1e12 / peaks["bf16_tflops"]},
                "segments": int(stats_l["segments"]), "cap": int(stats_l["cap"]),
            }
        except Exception as exc:  # keep the headline line even if the big batch fails
            large = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- BASELINE configs[0] on one GPU: CPU p50 beside GPU p50 through the client, and their parity ----
    config1 = None
    if world == 1 and rank == 0 and not args.no_config1:
        try:
            config1 = config1_section(np, torch, vod_b200, local_rank)
        except Exception as exc:  # noqa: BLE001
            config1 = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- the north-star target shape (N>1): 100M x 768 fp16 row-sharded, top-100 and top-1000 ----
    target = None
    if world > 1 and not args.no_target:
        corpus.close()
        corpus = None
        torch.cuda.empty_cache()
        try:
            target = target_section(np, torch, vod_b200, args, rank, world, local_rank, peaks, timed_section, max_over_ranks,
                                    all_ranks_true)
        except Exception as exc:  # noqa: BLE001
            target = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference arm's code path, fewer steps ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(args)

    if rank == 0:
        line = {
            "metric": f"mips_top{TOP_K}_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": {"bfloat16": "bf16", "float16": "fp16"}[args.store_dtype],
            "data": "synthetic", "config": workload_config(args),
            "arithmetic": f"tcgen05 tensor cores, {args.store_dtype} inputs, fp32 accumulate",
            "sharding": (f"rows split over {world} ranks; exchange={args.exchange} (p2p = final select stores epoch-tagged "
                         "words into peer-mapped buffers, merge kernel waits on the tags; nccl = all-gather + merge)")
            if world > 1 else "single shard",
            "corpus_gb_per_s": args.rows * DIM * 2 / (ms_step * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step, "segments": int(stats["segments"]), "cap": int(stats["cap"]),
            "clocks": clocks, "latency": latency, "sustained": sustained, "parity": parity, "target_config": target, "config1": config1,
            "config4_retrieve_and_sample": config4, "large_batch": large,
        }
        print(json.dumps(line), flush=True)
    if corpus is not None:
        corpus.close()
    if world > 1:
        dist.destroy_process_group()


def parity_section(np, torch, dist, corpus, args, rank, world, dev, all_ranks_true, tag="configs[1] corpus"):
    """Untimed correctness of the merged multi-GPU result, for k = 100 and k = 1000, on every rank:
      a. fused peer-store exchange == NCCL all-gather + merge kernel, bit for bit (scores and ids);
      b. every rank holds the same merged result (checked against rank 0's by broadcast);
      c. the merged k-th score >= every shard's local k-th score, and the merged list is sorted, ids unique and in
         range (a shard's best rows cannot be missing from the merge);
      d. merged == exact merge (lexicographic sort on the host) of the all-gathered local lists, bit for bit;
      e. float64 re-scoring of returned ids from the counter-based generator (rows regenerated on the CPU by the
         oracle's generator, rank 0): |score - true| <= 1e-5 * max|true|.
    """
    from oracle import twin

    dcode = {"bfloat16": 1, "float16": 2}[args.store_dtype]
    out = {"what": tag, "checked": [], "ok": True}
    nq = Q_SMALL
    xq_np = make_queries(np, 1, nq, args.store_dtype)[0]
    xq = torch.from_numpy(xq_np).to(dev)
    has_p2p = corpus._xchg is not None
    out["exchange_checked"] = "p2p vs nccl" if has_p2p else "nccl only (corpus created with exchange=nccl)"
    for k in (100, 1000):
        if nq * k > corpus._xchg_limits[0] * corpus._xchg_limits[1]:
            continue
        # (with --exchange nccl the corpus has no peer-mapped buffers: the NCCL result then stands in for both)
        s_nccl, i_nccl = corpus.search_device(xq, k, mode="tensor", exchange="nccl")
        s_p2p, i_p2p = corpus.search_device(xq, k, mode="tensor", exchange="p2p") if has_p2p else (s_nccl, i_nccl)
        s_loc, i_loc = corpus.store.search_device(xq, k, mode="tensor") if corpus.hi > corpus.lo else (None, None)
        torch.cuda.synchronize()
        over = corpus.any_overflow()
        a = bool(torch.equal(s_p2p, s_nccl) and torch.equal(i_p2p, i_nccl))
        ref_s, ref_i = s_p2p.clone(), i_p2p.clone()
        dist.broadcast(ref_s, 0)
        dist.broadcast(ref_i, 0)
        b = bool(torch.equal(ref_s, s_p2p) and torch.equal(ref_i, i_p2p))
        c = bool((s_p2p[:, :-1] >= s_p2p[:, 1:]).all()) and bool((i_p2p >= 0).all()) and bool((i_p2p < corpus.n_total).all())
        c = c and all(len(set(row.tolist())) == k for row in i_p2p[:8].cpu())
        if s_loc is not None:
            c = c and bool((s_p2p[:, -1] >= s_loc[:, -1]).all())
        # d. exact host merge of the gathered local lists
        if s_loc is None:
            s_loc = torch.full((nq, k), -3.4028234663852886e38, device=dev)
            i_loc = torch.full((nq, k), -1, dtype=torch.int64, device=dev)
        g_s = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
        g_i = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(g_s.view(world * nq, k), s_loc.contiguous())
        dist.all_gather_into_tensor(g_i.view(world * nq, k), i_loc.contiguous())
        hs = g_s.permute(1, 0, 2).reshape(nq, world * k).cpu().numpy()
        hi = g_i.permute(1, 0, 2).reshape(nq, world * k).cpu().numpy()
        order = np.lexsort((np.where(hi < 0, np.iinfo(np.int64).max, hi), -hs.astype(np.float64)), axis=1)[:, :k]
        d = bool(np.array_equal(np.take_along_axis(hs, order, 1), s_p2p.cpu().numpy())
                 and np.array_equal(np.take_along_axis(hi, order, 1), i_p2p.cpu().numpy()))
        e, err = True, None
        if rank == 0:
            ids = i_p2p[:4].cpu().numpy()
            sc = s_p2p[:4].cpu().numpy().astype(np.float64)
            true = np.zeros_like(sc)
            for q in range(ids.shape[0]):
                rows = np.stack([twin.synth_rows(CORPUS_SEED, int(r), 1, DIM, dtype=dcode)[0] for r in ids[q]])
                true[q] = rows.astype(np.float64) @ xq_np[q].astype(np.float64)
            err = float(np.abs(sc - true).max() / np.abs(true).max())
            e = err <= 1e-5
        ok = all_ranks_true(a and b and c and d and e and not over)
        out["checked"].append({"k": k, "queries": nq, "p2p_equals_nccl": all_ranks_true(a), "same_on_every_rank": all_ranks_true(b),
                               "sorted_unique_in_range_kth_ge_local_kth": all_ranks_true(c),
                               "equals_exact_host_merge_of_local_lists": all_ranks_true(d),
                               "fp64_rescore_rel_err_rank0": err, "fp64_rescore_ok": all_ranks_true(e), "overflow": over})
        out["ok"] = bool(out["ok"] and ok)
    # f. the host-buffer client path (all-rank overflow retry inside the call) returns the same bits as the device path
    s_h, i_h = corpus.search(xq_np, 100, mode="tensor")
    s_d, i_d = corpus.search_device(xq, 100, mode="tensor")
    f = all_ranks_true(bool(np.array_equal(s_h, s_d.cpu().numpy()) and np.array_equal(i_h, i_d.cpu().numpy())))
    out["checked"].append({"host_client_path_equals_device_path": f})
    out["ok"] = bool(out["ok"] and f)
    return out


def target_section(np, torch, vod_b200, args, rank, world, local_rank, peaks, timed_section, max_over_ranks, all_ranks_true):
    """BASELINE configs[2] = the north-star target shape: 100M x 768 fp16 row-sharded over the N ranks, 64-query
    (HBM-bound) and 8192-query (tensor-bound) batches, top-100 and top-1000, fused exchange + merge included.
    Fractions: scoring kernels and whole step against the measured HBM peak (64 queries) and the measured burst /
    sustained bf16 tensor peaks (8192 queries)."""
    import torch.distributed as dist

    rows = args.target_rows
    corpus = vod_b200.ShardedCorpus(rows, DIM, dtype="float16", device=local_rank, rank=rank, world_size=world,
                                    exchange=args.exchange, max_queries=Q_LARGE, max_k=1000)
    t0 = time.perf_counter()
    corpus.fill_synthetic(CORPUS_SEED)
    torch.cuda.synchronize()
    fill_s = time.perf_counter() - t0
    shard_rows = corpus.hi - corpus.lo
    shard_bytes = shard_rows * DIM * 2
    out = {"workload": f"BASELINE configs[2]: {rows} x {DIM} fp16 row-sharded over {world} GPUs, fused exchange + merge",
           "rows": rows, "shard_rows_rank0": shard_rows, "store_gb_per_gpu": shard_bytes / 1e9,
           "synthetic_fill_gb_per_s_per_gpu": shard_bytes / 1e9 / fill_s, "runs": []}
    # the two HBM-bound measurements first, on a GPU that is not yet heated up by the tensor-bound ones (the 1 kW power
    # cap lowers the SM clock, and with it the L2 -> SM bandwidth, for a while after a long 8192-query batch)
    for k in (100, 1000):
        ms, _, st, prof, _ = timed_section(corpus, Q_SMALL, k, 20, 5, False, "float16")
        score_ms = max_over_ranks(prof["score_ms"] / 20)
        out["runs"].append({
            "top_k": k, "queries_per_batch": Q_SMALL, "ms_per_step": ms, "queries_per_s": Q_SMALL / (ms * 1e-3),
            "corpus_tb_per_s_aggregate": rows * DIM * 2 / (ms * 1e-3) / 1e12, "bound": "hbm",
            "score_kernel_frac": shard_bytes / (score_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "select_ms": prof["select_ms"] / 20,
            "whole_step_frac": shard_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "target_whole_step_frac": 0.80,
            "segments": int(st["segments"])})
    for k in (100, 1000):
        ms, _, st, prof, _ = timed_section(corpus, Q_LARGE, k, 2, 1, False, "float16")
        score_ms = max_over_ranks(prof["score_ms"] / 2)
        flops = 2.0 * Q_LARGE * shard_rows * DIM
        out["runs"].append({
            "top_k": k, "queries_per_batch": Q_LARGE, "ms_per_step": ms, "queries_per_s": Q_LARGE / (ms * 1e-3),
            "tflops_per_gpu_whole_step": flops / (ms * 1e-3) / 1e12, "bound": "tensor",
            "score_kernel_frac": flops / (score_ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
            "select_ms": prof["select_ms"] / 2,
            "whole_step_frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
            "whole_step_frac_of_sustained": flops / (ms * 1e-3) / 1e12 / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
            "target_whole_step_frac": 0.60, "segments": int(st["segments"])})
    targs = argparse.Namespace(**{**vars(args), "store_dtype": "float16"})
    out["parity"] = parity_section(np, torch, dist, corpus, targs, rank, world, torch.device(f"cuda:{local_rank}"),
                                   all_ranks_true, tag="configs[2] corpus (100M x 768 fp16)")
    corpus.close()
    return out


if __name__ == "__main__":
    main()
