"""Regenerates profiles/r02_summary.md from the committed round-2 bench lines, sweeps and ncu raw pages.
    python scripts/make_profile_summary_r02.py"""
import csv
import json
import pathlib

P = pathlib.Path(__file__).resolve().parents[1] / "profiles"


def line(name):
    f = P / name
    if not f.exists():
        return None
    for ln in reversed(f.read_text().strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return None


def jsonl(name):
    f = P / name
    return [json.loads(l) for l in f.read_text().splitlines() if l.startswith("{")] if f.exists() else []


def ncu_rows(name, want):
    f = P / name
    if not f.exists():
        return []
    rows = list(csv.reader(f.open()))
    hdr = rows[0]
    idx = {w: hdr.index(w) for w in want if w in hdr}
    kn = hdr.index("Kernel Name")
    return [(r[kn], {w: r[i] for w, i in idx.items()}) for r in rows[2:]]


out = []
w = out.append
w("# Round 2 — profile summary (B200, sm_100a)\n")
w("All numbers from `gpurun` boxes. Peaks: `MEASURED_PEAKS.json` — HBM 6534.8 GB/s (copy kernel), bf16 1671.7 TFLOP/s burst / "
  "1404.1 sustained (cuBLAS). Workload unless noted: BASELINE configs[1], 10M x 768 bf16, exact top-100. The same build varies by "
  "about +-5% between boxes and between moments on one box (2.19-2.33 ms for the 64-query step): the L2 -> SM path that feeds these "
  "kernels scales with the SM clock, which the 1 kW power cap moves. Regenerate with `python scripts/make_profile_summary_r02.py`.\n")
w("Files: `r02a_*` ncu of the 64-query scan of a 1.25M-row shard BEFORE the counter fix (raw page + source-level stall table); "
  "`r02b_*` ncu raw pages of the CTA-pair kernels (8192 one-term queries; 64 three-term queries); `r02c_*` final single-GPU "
  "state: bench line, launch list of the bench command, ncu raw page of the headline kernel, batch-size and schedule sweeps, "
  "2-GPU bench lines (both arms) and multi-GPU test log; `r02d_*` 4- and 8-GPU bench lines and test log; `r02e_*` 8-GPU bench "
  "lines after the exchange buffer became plane-major (fused and NCCL exchange, reference arm under torchrun), k = 1000 probe; "
  "`r02f_*` ncu raw page of the kernels beside the scan, launch list of a small-shard search; `r02g_bench.json` bench line of the "
  "final build (another box: 2.215 ms), `r02_compute_sanitizer.txt` memcheck / racecheck of the final build; "
  "`r02k_*` probes of the power-cap steps inside a loop of identical searches and of the idle gap before a search; "
  "`r02m_*` bench lines (both arms) of the last build of the round, launch list of the bench command and ncu raw page of its select kernels, A/B of the resident three-term query tile; "
  "`r02l_*` the thread-maximum selection: A/B of the select time, ncu raw page of its first version, bench lines of the final build at 1 / 2 / 8 GPUs "
  "(8 GPUs without the target-shape and 8192-query sections), k = 1000 probe, 2-GPU test log; "
  "`r02j_bench.json` bench line after host-resident queries are classified on the host (e2e 2.44 ms next to a 2.22 ms device step on "
  "that box; 2.57-2.67 ms before), `r02j_test_multigpu_n2.log`; `r02h_*` ncu raw pages of the valley shapes (256 and 384 queries) and the A/B sweep of the dropped wide pair kernel; "
  "`r02_sass_opcodes.txt` opcode histogram of `libvodb.so`; `traffic.json` the DRAM-traffic ratios bench.py multiplies with.\n")

b1 = line("r02c_bench.json")
if b1:
    r, e, lb = b1["roofline"], b1["e2e"], b1["large_batch"]
    c1, c4 = b1["config1"], b1["config4_retrieve_and_sample"]
    w("## Headline, 1 GPU (r02c_bench.json; round-1 driver run BENCH_r01 in brackets)\n")
    w("| quantity | value |\n|---|---|")
    w(f"| 64-query batches, inputs resident in HBM | {b1['value']:.0f} queries/s, {b1['ms_per_step']:.4f} ms/step (28290, 2.2623) |")
    w(f"| corpus scanned, whole step | {b1['corpus_gb_per_s']:.0f} GB/s = {100 * r['whole_step_frac']:.1f}% of measured HBM peak (103.9%) |")
    w(f"| scoring kernel `score_tc2_kernel<64,1,resident>` alone (CUDA events) | {r['score_kernel_ms_per_search']:.4f} ms = {r['achieved']:.0f} GB/s = "
      f"{100 * r['frac']:.1f}% (105.7%) |")
    w(f"| DRAM traffic of those launches (ncu, `r02c_q64_resident_ncu_raw.csv`) | 15.361 GB read + 13 MB written for 15.360 GB algorithmic = 1.001x; "
      "`l1tex__m_xbar2l1tex_read_bytes` 15.40 GB: the L2 -> SM path carries the corpus and nothing else |")
    w(f"| kernels per search | {b1['gpu_launches_per_step']} (prepare + {b1['segments']} x (score, select)), list capacity {b1['cap']} (9, 4 segments) |")
    w(f"| select kernels per search | {1e3 * r['select_kernel_ms_per_search']:.1f} us |")
    full_v = e.get("value_full_mantissa_f32_queries", e["value"])
    full_ms = e.get("ms_per_step_full_mantissa_f32_queries", e["ms_per_step"])
    w(f"| e2e `B200SearchClient.search(np.ndarray)`, float32 queries holding bf16 values (bf16-mixed encoder, as round 1 timed; correction terms "
      f"skipped on the device) | {e['value_store_dtype_exact_queries']:.0f} queries/s, {e['ms_per_step_store_dtype_exact_queries']:.3f} ms (25149, 2.545) |")
    w(f"| e2e, float32 queries with full mantissas (3 query terms really scored) | {full_v:.0f} queries/s, {full_ms:.3f} ms |")
    w(f"| per-call latency p10 / p50 / p90 | {b1['latency']['p10']:.3f} / {b1['latency']['p50']:.3f} / {b1['latency']['p90']:.3f} ms |")
    w(f"| 8192-query batches (`score_tc2_kernel<256,1>`) | {lb['value']:.0f} queries/s, {lb['ms_per_step']:.1f} ms; scoring kernels "
      f"{lb['roofline']['achieved']:.0f} TFLOP/s = {100 * lb['roofline']['frac']:.1f}% of burst, {100 * lb['roofline']['frac_of_sustained']:.1f}% of sustained; "
      f"whole step {100 * lb['roofline']['whole_step_frac']:.1f}% |")
    cb = b1["cpu_baseline"]
    w(f"| CPU arm (oracle port, every step scans all 10M rows, same queries) | {cb['value']:.1f} queries/s on {cb['cores']} cores, BLAS threads {cb['blas_threads']} |")
    w(f"| configs[0] (100k x 768 fp32, 256 queries) through the client | GPU auto (tensor3 on bf16 planes) p50 {c1['gpu_auto_ms']['p50']:.3f} ms, "
      f"CUDA-core exact {c1['gpu_exact_cuda_cores_ms']['p50']:.3f} ms, CPU port {c1['cpu_ms']['p50']:.1f} ms; parity ok: "
      f"{c1['gpu_auto_parity']['ok']} / {c1['gpu_exact_cuda_cores_parity']['ok']} (max score error {c1['gpu_auto_parity']['max_score_rel_err']:.1e}) |")
    dw = c4["dataloader_workers"]
    w(f"| configs[3] chain (32 queries -> top-1000 -> sample 8, one call) | p50 {c4['chain_ms_p50']:.3f} ms; sampler kernel {c4['sampler_kernel_us_p50']:.0f} us "
      f"vs C twin on one host thread {c4['sampler_cpu_twin_us']['p50']:.0f} us; `sample_search_results` host call {1e3 * c4['sample_search_results_host_call_ms_p50']:.0f} us |")
    w(f"| configs[3] as 8 DataLoader worker processes see it | shared scans {dw['coalesced']:.0f} queries/s (runs: "
      f"{', '.join(f'{x:.0f}' for x in dw['coalesced_detail']['qps_all_runs'])}; {dw['coalesced_detail']['queries_per_scan_p50']} queries per scan, "
      f"{dw['coalesced_detail']['scan_ms_p50']:.2f} ms per scan) vs one scan per request {dw['one_scan_per_request']:.0f} "
      f"({dw['one_scan_per_request_detail']['scan_ms_p50']:.2f} ms per 32-query scan) |")
    w(f"| clocks during the timed region | {b1['clocks']} |\n")

w("## Rehearsal of the driver's SCALE procedure (final build, one 8-GPU box, both arms at N = 1, 2, 4, 8 back to back with default flags: r02i_scale_*.json)\n")
w("| GPUs | queries/s (ms/step) | efficiency | scoring kernels / whole step vs HBM peak | e2e queries/s (ms), efficiency | e2e ms, full-mantissa queries | 8192 q: q/s, TFLOP/s per GPU, whole step | CPU arm q/s (threads) | value / CPU | parity | same config |")
w("|---|---|---|---|---|---|---|---|---|---|---|")
_b = line("r02i_scale_n1.json")
for n in (1, 2, 4, 8):
    d, rr = line(f"r02i_scale_n{n}.json"), line(f"r02i_scale_reference_n{n}.json")
    if not d or not rr or not _b:
        continue
    rf, lb, e = d["roofline"], d["large_batch"], d["e2e"]
    w(f"| {n} | {d['value']:.0f} ({d['ms_per_step']:.4f}) | {d['value'] / _b['value'] / n:.3f} | {100 * rf['frac']:.1f}% / {100 * rf['whole_step_frac']:.1f}% | "
      f"{e['value']:.0f} ({e['ms_per_step']:.3f}), {e['value'] / _b['e2e']['value'] / n:.3f} | {e['ms_per_step_full_mantissa_f32_queries']:.3f} | "
      f"{lb['value']:.0f}, {lb['roofline']['achieved']:.0f}, {100 * lb['roofline']['whole_step_frac']:.1f}% | {rr['value']:.1f} ({rr['cpu_baseline']['blas_threads']}) | "
      f"{d['value'] / rr['value']:.0f}x | {d['parity']['ok'] if d.get('parity') else '-'} | {d['config'] == rr['config']} |")
w("\n(The 4-GPU 8192-query whole-step figure of this run, 60%, is an outlier of a 3-step measurement: 24.6 ms / 76% in r02d_bench_n4.json; the default is now 5 steps.)\n")
for n in (2, 4, 8):
    d = line(f"r02i_scale_n{n}.json")
    if not d or not d.get("target_config") or "runs" not in d["target_config"]:
        continue
    t = d["target_config"]
    w(f"Target shape at {n} GPUs (`r02i_scale_n{n}.json` -> `target_config`; parity ok = {t['parity']['ok']}): " + "; ".join(
        f"top-{run['top_k']} x {run['queries_per_batch']} q: {run['ms_per_step']:.2f} ms = {100 * run['whole_step_frac']:.1f}% (target {100 * run['target_whole_step_frac']:.0f}%)"
        for run in t["runs"]) + "\n")

# final build (thread-maximum selection)
fl = [(n, line(f)) for n, f in ((1, "r02m_bench.json"), (2, "r02l_bench_n2.json"), (8, "r02l_bench_n8.json"))]
if all(d for _, d in fl):
    w("## Final build: thread-maximum selection in front of the radix select (r02m_bench.json, r02l_bench_n2.json, r02l_bench_n8.json, r02l_select_fast_ab.jsonl)\n")
    w("| GPUs | queries/s (ms/step) | scoring kernels / whole step vs HBM peak | select kernels per search | e2e ms | 250 searches in a row, ms/step by blocks of 50 | parity |")
    w("|---|---|---|---|---|---|---|")
    for n, d in fl:
        rf = d["roofline"]
        w(f"| {n} | {d['value']:.0f} ({d['ms_per_step']:.4f}) | {100 * rf['frac']:.1f}% / {100 * rf['whole_step_frac']:.1f}% | {rf['select_kernel_ms_per_search'] * 1e3:.1f} us | "
          f"{d['e2e']['ms_per_step']:.3f} | {', '.join(f'{x:.3f}' for x in d['sustained']['ms_per_step_by_50'])} | {d['parity']['ok'] if d.get('parity') else '-'} |")
    ab = jsonl("r02m_select_fast_ab.jsonl")
    if ab:
        a0 = next(r for r in ab if r["fast_select"] == "0")
        a1 = next(r for r in ab if r["fast_select"] == "1" and r["threads"] == "default")
        w(f"\nSame box, 1.25M-row shard, 64 queries (`r02m_select_fast_ab.jsonl`): radix select only {a0['search_ms']:.4f} ms per search "
          f"({a0['select_ms_per_search'] * 1e3:.1f} us in two selects), with the thread-maximum bound {a1['search_ms']:.4f} ms ({a1['select_ms_per_search'] * 1e3:.1f} us). "
          "The 1-GPU line above comes from a box that ran the whole bench under `sw_power_cap` (2.24-2.27 ms on the last two boxes; 2.17-2.22 on the other boxes of the round, same kernels); "
          "8 GPUs: 0.344 -> 0.334 ms, whole step 85.4% -> 88.0% of the HBM roofline against the SCALE rehearsal below.\n")

def _launch_rows(name):
    import csv as _c, re as _r
    f = P / name
    if not f.exists():
        return []
    rows = list(_c.reader(f.open()))
    h = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
    if h is None:
        return []
    hd = rows[h]
    ki, vi = hd.index("Kernel Name"), hd.index("Metric Value")
    out_ = []
    for r in rows[h + 1:]:
        if len(r) == len(hd):
            m = _r.search(r"(\w+_kernel(?:<[^(]*?>)?)\(", r[ki])
            out_.append((m.group(1) if m else r[ki][:40], float(r[vi]) / 1e3))
    return out_
_lr = _launch_rows("r02m_launches_ncu.csv")
_starts = [i for i, (k_, _) in enumerate(_lr) if k_.startswith("prepare_kernel")]
if len(_starts) >= 5:
    seq = _lr[_starts[3]:_starts[4]]
    tot = sum(us for _, us in seq)
    w("## One search of the bench command, final build, per launch (r02m_launches_ncu.csv: `ncu --metrics gpu__time_duration.sum` over `bench.py --steps 2 --warmup 1`; cold cache, serialised)\n")
    w("| kernel | us | share |\n|---|---|---|")
    for k_, us in seq:
        w(f"| `{k_}` | {us:.1f} | {100 * us / tot:.1f}% |")
    w(f"\nScoring kernels {100 * sum(us for k_, us in seq if k_.startswith('score')) / tot:.1f}% of the search (CUDA events in the bench line: "
      "`roofline.score_kernel_ms_per_search` / `ms_per_step`, same share); the selects were 18-20 us each before the thread-maximum bound (r02c_launches_ncu.csv).\n")

w("## Strong scaling during the round (bench lines r02c / r02d / r02e; fused exchange)\n")
w("| GPUs | file | 64 q: queries/s (ms) | vs 1 GPU | scoring kernels / whole step vs HBM peak | e2e ms (f32 / bf16-exact queries) | 8192 q: q/s, TFLOP/s per GPU | parity |")
w("|---|---|---|---|---|---|---|---|")
base = b1["value"] if b1 else None
for n, f in ((1, "r02c_bench.json"), (2, "r02c_bench_n2.json"), (4, "r02d_bench_n4.json"), (8, "r02d_bench_n8.json"),
             (8, "r02e_bench_n8.json"), ("8 (NCCL all-gather + merge)", "r02e_bench_n8_nccl.json")):
    d = line(f)
    if not d:
        continue
    r, lb = d["roofline"], d.get("large_batch")
    par = d["parity"]["ok"] if d.get("parity") else "-"
    ng = int(str(n).split()[0])
    big = f"{lb['value']:.0f}, {lb['roofline']['achieved']:.0f}" if lb else "-"
    w(f"| {n} | {f} | {d['value']:.0f} ({d['ms_per_step']:.4f}) | {d['value'] / base:.2f}x = {d['value'] / base / ng:.3f} | {100 * r['frac']:.1f}% / {100 * r['whole_step_frac']:.1f}% | "
      f"{d['e2e'].get('ms_per_step_full_mantissa_f32_queries', d['e2e']['ms_per_step']):.3f} / {d['e2e']['ms_per_step_store_dtype_exact_queries']:.3f} | {big} | {par} |")
ref1, ref8 = line("r02c_bench_reference_n2.json"), line("r02e_bench_reference_n8.json")
if ref1 and ref8:
    w(f"\nReference arm under torchrun (rank 0 alone, BLAS threads set before numpy loads): {ref1['value']:.1f} queries/s at `--gpus 2` "
      f"({ref1['cpu_baseline']['blas_threads']} threads), {ref8['value']:.1f} at `--gpus 8` ({ref8['cpu_baseline']['blas_threads']} threads); "
      "round 1 fell from 31 to 6 queries/s there because torchrun exports OMP_NUM_THREADS=1.")
w("")
for n, f in ((2, "r02c_bench_n2.json"), (8, "r02d_bench_n8.json"), (8, "r02e_bench_n8.json")):
    d = line(f)
    if not d or not d.get("target_config") or "runs" not in d["target_config"]:
        continue
    t = d["target_config"]
    w(f"Target shape at {n} GPUs (`{f}` -> `target_config`: {t['workload']}; {t['store_gb_per_gpu']:.1f} GB per GPU, filled at "
      f"{t['synthetic_fill_gb_per_s_per_gpu']:.0f} GB/s per GPU; parity ok = {t['parity']['ok']}):\n")
    w("| top-k | queries | ms/step | queries/s | scoring kernels | whole step | target |")
    w("|---|---|---|---|---|---|---|")
    for run in t["runs"]:
        extra = f" ({100 * run['whole_step_frac_of_sustained']:.1f}% of sustained)" if "whole_step_frac_of_sustained" in run else ""
        w(f"| {run['top_k']} | {run['queries_per_batch']} | {run['ms_per_step']:.3f} | {run['queries_per_s']:.0f} | {100 * run['score_kernel_frac']:.1f}% | "
          f"{100 * run['whole_step_frac']:.1f}%{extra} | >= {100 * run['target_whole_step_frac']:.0f}% of the {run['bound']} roofline |")
    w("")

sw = jsonl("r02c_batch_sweep.jsonl")
r1 = {d["nq"]: d for d in jsonl("r01k_batch_sweep.jsonl")}
if sw:
    w("## Roofline curve over the batch size (r02c_batch_sweep.jsonl vs r01k_batch_sweep.jsonl; back-to-back searches)\n")
    w("| queries | ms | roofline ms (bound) | fraction | round 1 ms (fraction) | segments |")
    w("|---|---|---|---|---|---|")
    for d in sw:
        o = r1.get(d["nq"])
        w(f"| {d['nq']} | {d['ms']:.3f} | {d['roofline_ms']:.3f} ({d['bound']}) | {100 * d['frac']:.1f}% | "
          + (f"{o['ms']:.3f} ({100 * o['frac']:.1f}%)" if o else "-") + f" | {d['segments']} |")
    w("\nWhat moved it: counters one per 256-byte line (64 -> 128 queries), MMA width and query box following the query count, "
      "narrow last tiles (96, 192, 384), queries resident in shared memory (<= 128). 256-512 queries remain the valley: HBM, the "
      "L2 -> SM path (2 x corpus bytes on a pair) and the tensor pipe are all within 10% of each other there.\n")

sched = jsonl("r02c_schedule_sweep.jsonl")
if sched:
    w("## Scan schedule after the counter fix (r02c_schedule_sweep.jsonl, 64 queries; `first` rows dumped, growth g, list capacity)\n")
    w("| shard rows | first | growth | cap | segments | ms | scoring ms | select ms |")
    w("|---|---|---|---|---|---|---|---|")
    for d in sched:
        w(f"| {d['rows']} | {d['first'] or 'default'} | {d['growth'] or '-'} | {d['cap']} | {d['segments']} | {d['ms']:.3f} | {d['score_ms']:.3f} | {d['select_ms']:.3f} |")
    w("\n(`default` = the round-1 schedule at the time of the sweep: 4096 rows, growth 20, capacity 16384. Before the counter fix the "
      "two-segment schedules cost 0.40 ms on the 1.25M-row shard against 0.34 ms; after it they are the fastest, which is why the "
      "planner now dumps 16384 rows and grows up to 96-fold.)\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread"]
w("## ncu `--set full` raw pages (per launch: ms, DRAM GB read, DRAM active %, L2 throughput %, tensor pipe active %, L2->SM GB)\n")
for name, what in (("r02a_small_shard_score_tc64_ncu_raw.csv", "64 queries, 1.25M-row shard, round-1 kernel and packed counters: the middle segment runs at 33% DRAM"),
                   ("r02c_q64_resident_ncu_raw.csv", "64 queries, 10M rows, headline kernel (resident queries): 3 segments"),
                   ("r02b_pair64x3_q64_ncu_raw.csv", "64 float32 queries with three real bf16 terms, `score_tc2_kernel<64,3>`: neither DRAM nor tensor pipe saturated"),
                   ("r02b_pair256_q8192_ncu_raw.csv", "8192 queries, `score_tc2_kernel<256,1>` with the blocked item order, segments 3-7: DRAM read = 1.02x algorithmic (2.86x in round 1)"),
                   ("r02f_aux_kernels_ncu_raw.csv", "the kernels beside the 16-bit tensor scan (scripts/ncu_aux_probe.py, 1M x 768): `score_exact_kernel` after its round-1 rewrite (64 queries: "
                    "2.29 ms = 43 TFLOP/s fp32 — test / fallback path only since `auto` serves fp32 stores from the tensor cores), `split_planes_kernel`, the fp32-store tensor "
                    "kernel `score_tc_kernel<64,3,3>` (0.73 ms for the same search, DRAM 76% on its 4.6 GB of planes), selects, sampler chain, 8-list merge")):
    rows = ncu_rows(name, want)
    if not rows:
        continue
    w(f"`{name}` — {what}\n")
    w("| kernel | ms | DRAM GB | DRAM % | L2 % | tensor % | L2->SM GB | regs |")
    w("|---|---|---|---|---|---|---|---|")
    for k, m in rows:
        kname = k.split("(")[0].replace("void vodb::<unnamed>::", "").replace("void unnamed>::", "")
        w(f"| `{kname}` | {float(m[want[0]]):.3f} | {float(m[want[1]]):.3f} | {float(m[want[2]]):.1f} | {float(m[want[3]]):.1f} | {float(m[want[4]]):.1f} | "
          f"{float(m.get(want[5], 'nan')):.2f} | {m[want[6]]} |")
    w("")

f = P / "r02c_launches_ncu.csv"
if f.exists():
    rows = list(csv.reader(f.open()))
    hdr, launches = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            launches.append((d["Kernel Name"].split("(")[0].replace("void vodb::<unnamed>::", "").replace("void unnamed>::", ""), d["Grid Size"],
                             d["Block Size"], float(d["Metric Value"]) / 1e3))
    # first complete 64-query search after warm-up: prepare ... until the next prepare
    starts = [i for i, l in enumerate(launches) if l[0].startswith("prepare_kernel")]
    if len(starts) > 3:
        a, b = starts[2], starts[3]
        w("## One 64-query search, per launch (r02c_launches_ncu.csv: `ncu --metrics gpu__time_duration.sum` over the bench command; cold cache, serialised)\n")
        w("| kernel | grid x block | us |\n|---|---|---|")
        tot = sc = 0.0
        for k, g, blk, us in launches[a:b]:
            w(f"| `{k}` | {g} x {blk} | {us:.1f} |")
            tot += us
            sc += us if k.startswith("score_") else 0.0
        ev = b1["roofline"]["score_kernel_ms_per_search"] / b1["ms_per_step"] if b1 else float("nan")
        w(f"\nScoring kernels' share of the search under ncu: {100 * sc / tot:.1f}% ({sc:.0f} of {tot:.0f} us); from bench.py's CUDA events without "
          f"profiler: {100 * ev:.1f}%. The shares agree.\n")

f = P / "r02f_launches_small_shard_ncu.csv"
if f.exists():
    rows = list(csv.reader(f.open()))
    hdr, launches = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            launches.append((d["Kernel Name"].split("(")[0].replace("void vodb::<unnamed>::", "").replace("void unnamed>::", ""), d["Grid Size"],
                             d["Block Size"], float(d["Metric Value"]) / 1e3))
    starts = [i for i, l in enumerate(launches) if l[0].startswith("prepare_kernel")]
    if len(starts) > 2:
        a, b = starts[-2], starts[-1]
        w("## One 64-query search over a 1.25M-row shard = one rank of the 8-GPU split, per launch (r02f_launches_small_shard_ncu.csv; cold cache, serialised)\n")
        w("| kernel | grid x block | us |\n|---|---|---|")
        for k, g, blk, us in launches[a:b]:
            w(f"| `{k}` | {g} x {blk} | {us:.1f} |")
        w("\nRound 1 (r01m_launches_ncu.csv scaled to this shard): 4096-row dump 10.4 us, 86k-row segment 50 us, 1.16M-row segment ~265 us, three selects of "
          "~11 us; now a 16384-row dump, one scan and two selects. At 8 GPUs the merge kernel (fused exchange) follows; a multi-rank command cannot be "
          "wrapped in ncu, so the 8-GPU step is the CUDA-event time of `r02e_bench_n8.json` (0.350 ms whole step, 0.303 ms in the scoring kernels).\n")

# rank 0 of a two-rank search under ncu (single-pass collections only; rank 1 runs unprofiled beside it)
import csv as _csv
def _ncu_long(name):
    f = P / name
    if not f.exists():
        return []
    rows = list(_csv.reader(f.open()))
    h = next((i for i, r in enumerate(rows) if r and r[0] == "ID"), None)
    if h is None:
        return []
    hd = rows[h]
    return [dict(zip(hd, r)) for r in rows[h + 1:] if len(r) == len(hd)]
xl, xn = _ncu_long("r02k_xchg_launches_rank0_ncu.csv"), _ncu_long("r02k_xchg_nvlink_rank0_ncu.csv")
if xl and xn:
    ours = [r for r in xl if "vodb" in r["Kernel Name"] and "synth_fill" not in r["Kernel Name"]]
    def short(nm):
        import re as _re
        m = _re.search(r"(\w+_kernel(?:<[^(]*?>)?)\(", nm)
        return m.group(1) if m else nm[:40]
    w("## Fused exchange, rank 0 of a 2-rank search under ncu (r02k_xchg_launches_rank0_ncu.csv, r02k_xchg_nvlink_rank0_ncu.csv; scripts/r02_ncu_exchange.sh)\n")
    w("2 x 1.25M rows, 64 queries. Only rank 0's command line carries ncu and rank 1 enqueues its search 0.3 s earlier, so the words rank 0's merge waits for are "
      "already there. Collections that need kernel replay (`--set full`) fail in a process with CUDA-IPC peer mappings (ncu: UnknownError); single-pass ones work.\n")
    w("| search | kernel | us |\n|---|---|---|")
    names = ["warm-up, top-100", "top-100", "top-1000"]
    si = -1
    for r in ours:
        if "prepare_kernel" in r["Kernel Name"]:
            si += 1
        if si >= 1:
            w(f"| {names[min(si, 2)]} | `{short(r['Kernel Name'])}` | {float(r['Metric Value']) / 1e3:.1f} |")
    tx = {}
    for r in xn:
        tx.setdefault(r["ID"], {"k": short(r["Kernel Name"])})[r["Metric Name"]] = float(r["Metric Value"])
    w("\nNVLink bytes per launch (`nvltx__bytes.sum` / `nvlrx__bytes.sum`, launches of the top-100 then the top-1000 search):\n")
    w("| kernel | NVLink tx bytes | NVLink rx bytes |\n|---|---|---|")
    for i in sorted(tx, key=int):
        w(f"| `{tx[i]['k']}` | {tx[i].get('nvltx__bytes.sum', 0):.0f} | {tx[i].get('nvlrx__bytes.sum', 0):.0f} |")
    w("\nThe final select is the only kernel that touches NVLink: 188 KB sent for the 64 x 100 x 3 tagged words (153.6 KB) it stores into the peer's gather buffer "
      "(1.23x: 32-byte packet granularity + headers), 1.91 MB for top-1000 (1.536 MB of words); the merge kernel reads its own HBM only (0 bytes on the links), "
      "and nothing flows back but acknowledgements. With the exchange the final select takes 25.1 us against 19.9 us without it (1-GPU list above), the merge 10.0 us "
      "(30.8 us for 2 x 1000 entries per query).\n")

mg, ing = line("r02g_multigpu_store_n2_probe.json"), line("r02g_ingest_probe.json")
if mg and ing:
    w("## Single-process multi-GPU master and ingest (r02g_multigpu_store_n2_probe.json, r02g_ingest_probe.json)\n")
    w(f"`B200SearchMaster(store=MultiGpuStore(10M x 768 bf16, devices=[0, 1]))` — ONE process owning both GPUs, the reference's server shape — through "
      f"`client.search(np.ndarray)`: p50 {mg['bf16_exact_queries']['ms_p50']:.3f} ms per 64-query batch ({mg['bf16_exact_queries']['queries_per_s']:.0f} queries/s) with "
      f"bf16-exact float32 queries, {mg['full_f32_queries']['ms_p50']:.3f} ms with full-mantissa queries; identical to one store holding everything: "
      f"{mg['equals_single_store']}. (One process per GPU with the fused exchange, same box class: 1.335 / 1.496 ms, table above.) The per-search "
      "torch.stack / .cpu() chain of round 1 is gone: scans on all devices are enqueued first, lists reach the first device by event-ordered peer copies, one host wait.\n")
    w(f"Ingest, one GPU: float32 rows in pinned host memory -> bf16 store {ing['pinned_f32_to_bf16_store_gb_per_s']:.1f} GB/s of host data in 2^18-row `add` calls, "
      f"{ing['one_call_4gb_gb_per_s']:.1f} GB/s for one 4 GB call (two staging buffers, copy stream overlapping the conversion kernel; PCIe-bound: 52 GB/s in round 1); "
      f"`zarr_io.ingest` from an uncompressed zarr-v2 store with the reference's 100-row chunks (2000 files of 300 KB on tmpfs): {ing['zarr_uncompressed_chunks100_gb_per_s']:.2f} GB/s, "
      f"read-back equal: {ing['zarr_roundtrip_ok']}.\n")

w("## Where the time of a small-shard search went (r02a_small_shard_seg1_stalls.txt)\n")
w("Source-level stall sampling of the 86k-row middle segment before the fix: 31% of the warp samples are epilogue warps waiting for the "
  "next accumulator (the MMA warp is itself waiting for stages that the epilogue has not released), 16% wait for the threshold loads that "
  "queued behind the flush's atomics, 5% wait for atomic results — all symptoms of ~2000 same-line atomics per query counter. After "
  "spreading the counters (one per 256 bytes) and loading the thresholds before the atomics the 1.25M-row search went from 0.343 to "
  "0.327 ms with the old schedule and to 0.321 ms with two segments; 64-query searches over 10M rows from 2.43 to 2.22 ms.\n")
(P / "r02_summary.md").write_text("\n".join(out) + "\n")
print("wrote", P / "r02_summary.md", len(out), "lines")
