"""Builds profiles/r01_summary.md from the committed captures (bench lines, ncu launch list, ncu raw pages)."""
import csv, json, pathlib
P = pathlib.Path(__file__).resolve().parents[1] / "profiles"
def bench(name): return json.loads([l for l in (P / name).read_text().splitlines() if l.startswith('{')][-1])
out = []; A = out.append
A("# Round 1 — profile summary (B200, sm_100a)\n")
A("All numbers from `gpurun` boxes (1x B200 unless noted). Peaks: `MEASURED_PEAKS.json` — HBM 6534.8 GB/s (copy kernel), bf16 1671.7 TFLOP/s burst / 1404.1 sustained (cuBLAS). Workload: BASELINE configs[1], 10M x 768 bf16, exact top-100. Box-to-box variation of the same build is about +-4% (2.19-2.29 ms for the 64-query step).\n")
A("Files: `r01a_*` first working tcgen05 path; `r01b_*` dump mode + staged survivors; `r01c..g_*` bench lines along the way; `r01h_*` final state of the round (launch list, `ncu --set full` raw pages of `score_tc_kernel<64,1>` and the 2-CTA `score_tc2_kernel`, bench lines of both arms); `r01i_*` multi-GPU bench lines with the flag/fence exchange; `r01j_*` final bench lines of the round (1 GPU both arms, 2/4/8 GPUs with the epoch-tagged LL exchange, 8 GPUs NCCL); `r01l_*`/`r01m_*` bench lines, launch list and `ncu --set full` raw pages of the final tensor-core kernels; `r01n_bench*.json` the last bench lines of the round (after the CUDA-core kernel rework); `r01k_*` BASELINE configs[2] at full size (100M x 768 fp16 over 2/4/8 GPUs, top-1000), `ncu --set full` raw page of the kernels beside the tensor-core scan, latency probe; probes: `r01_schedule_sweep.jsonl`, `r01_query_terms_probe.json`, `r01_configs_3_5_probe.json`, `r01_compute_sanitizer.txt`. Regenerate this file with `python scripts/make_profile_summary.py`.\n")
d = bench("r01m_bench.json"); f = bench("r01h_bench.json")
A("## Headline (r01m_bench.json = final build of the round, same gpurun call as the r01m ncu captures below; r01h_bench.json = an earlier build on another box)\n")
A("| quantity | r01m | r01h |\n|---|---|---|")
A(f"| 64-query batches, inputs resident in HBM: queries/s (ms/step) | {d['value']:.0f} ({d['ms_per_step']:.4f}) | {f['value']:.0f} ({f['ms_per_step']:.4f}) |")
A(f"| corpus scanned, whole step | {d['corpus_gb_per_s']:.0f} GB/s = {d['roofline']['whole_step_frac']*100:.1f}% of measured HBM peak | {f['corpus_gb_per_s']:.0f} GB/s = {f['roofline']['whole_step_frac']*100:.1f}% |")
A(f"| scoring kernel alone (CUDA events around its launches) | {d['roofline']['score_kernel_ms_per_search']:.4f} ms = {d['roofline']['achieved']:.0f} GB/s = {d['roofline']['frac']*100:.1f}% | {f['roofline']['score_kernel_ms_per_search']:.4f} ms = {f['roofline']['achieved']:.0f} GB/s = {f['roofline']['frac']*100:.1f}% |")
A(f"| select kernels per search | {d['roofline']['select_kernel_ms_per_search']*1e3:.1f} us | {f['roofline']['select_kernel_ms_per_search']*1e3:.1f} us |")
A(f"| e2e through `B200SearchClient.search(np.ndarray)` (pinned H2D 196 KB + D2H 77 KB inside): queries/s (ms) | {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.3f}) | {f['e2e']['value']:.0f} ({f['e2e']['ms_per_step']:.3f}) |")
A(f"| per-call latency p10 / p50 / p90 (ms) | {d['latency']['p10']:.3f} / {d['latency']['p50']:.3f} / {d['latency']['p90']:.3f} | {f['latency']['p10']:.3f} / {f['latency']['p50']:.3f} / {f['latency']['p90']:.3f} |")
c4 = d['config4_retrieve_and_sample']
c4f = f['config4_retrieve_and_sample']
A(f"| config 4 chain (32 queries -> top-1000 -> sample 8, host in, [32,8] out; one `vodb_retrieve_sample` call in r01m) | p50 {c4['chain_ms_p50']:.3f} ms; sampler kernel p50 {c4['sampler_kernel_us_p50']:.1f} us; `sample_search_results` on host arrays p50 {c4['sample_search_results_host_call_ms_p50']*1e3:.0f} us | p50 {c4f['chain_ms_p50']:.3f} ms |")
w = c4['dataloader_workers']
A(f"| config 4 as {w['workers']} DataLoader worker processes see it (32-query top-1000 requests over the Unix socket) | {w['coalesced']:.0f} queries/s with shared scans ({w['requests_served']} requests in {w['scans_issued']} scans) vs {w['one_scan_per_request']:.0f} one scan per request | - |")
lb, lf = d['large_batch'], f['large_batch']
A(f"| 8192-query batches (2-CTA kernel): queries/s (ms/step) | {lb['value']:.0f} ({lb['ms_per_step']:.1f}) | {lf['value']:.0f} ({lf['ms_per_step']:.1f}) |")
A(f"| 8192-query scoring kernels | {lb['roofline']['achieved']:.0f} TFLOP/s = {lb['roofline']['frac']*100:.1f}% of burst peak, {lb['roofline']['frac_of_sustained']*100:.1f}% of sustained | {lf['roofline']['achieved']:.0f} TFLOP/s = {lf['roofline']['frac']*100:.1f}% |")
A(f"| CPU baseline (oracle port: numpy/OpenBLAS sgemm + exact top-k, {d['cpu_baseline']['cores']} host cores, 500k-row sample x20) | {d['cpu_baseline']['value']:.1f} queries/s | - |")
A(f"| kernels per 64-query search | {d['gpu_launches_per_step']} (prepare + 4 x (score, select)), all launched with PDL | |")
A(f"| clocks during the timed region | {d['clocks']} | |\n")
A("## Strong scaling, same 10M-row corpus (queries/s, 64-query batches)\n")
A("| GPUs | exchange | file | queries/s | ms/step | vs 1 GPU of the same series | 8192-query batch q/s |\n|---|---|---|---|---|---|---|")
series = [("early (before select/PDL work)", [(1, 'r01c_bench.json'), (2, 'r01c_bench_n2_p2p.json'), (2, 'r01c_bench_n2_nccl.json'), (4, 'r01d_bench_n4_p2p.json'), (8, 'r01d_bench_n8_p2p.json'), (8, 'r01d_bench_n8_nccl.json')]),
          ("flag + fence exchange", [(1, 'r01h_bench.json'), (2, 'r01i_bench_n2_p2p.json'), (4, 'r01i_bench_n4_p2p.json'), (8, 'r01i_bench_n8_p2p.json'), (8, 'r01i_bench_n8_nccl.json')]),
          ("final (epoch-tagged LL exchange)", [(1, 'r01j_bench.json'), (2, 'r01j_bench_n2_p2p.json'), (4, 'r01j_bench_n4_p2p.json'), (8, 'r01j_bench_n8_p2p.json'), (8, 'r01j_bench_n8_nccl.json')])]
for label, files in series:
    base = None
    for n, fn in files:
        if not (P / fn).exists(): continue
        x = bench(fn)
        if n == 1: base = x['value']
        ex = '-' if n == 1 else ('p2p (fused into select / merge kernels)' if 'p2p' in fn else 'NCCL all-gather')
        l = (x.get('large_batch') or {}).get('value')
        A(f"| {n} ({label}) | {ex} | {fn} | {x['value']:.0f} | {x['ms_per_step']:.4f} | {x['value']/base:.2f}x | {'%.0f' % l if l else '-'} |")
A("")
lines = [l for l in open(P / 'r01m_launches_ncu.csv') if not l.startswith('==')]
r = list(csv.DictReader(lines))
# first complete search = first 'prepare' after the synthetic fill
start = next(i for i, row in enumerate(r) if 'prepare_kernel' in row['Kernel Name'])
start = next(i for i, row in enumerate(r) if 'prepare_kernel' in row['Kernel Name'] and i > start)
A("## One 64-query search, per launch (ncu launch list r01m_launches_ncu.csv: cold cache, serialised)\n")
A("| # | kernel | grid x block | us |\n|---|---|---|---|")
tot = sc = 0
for row in r[start:start + 9]:
    name = row['Kernel Name']; short = name.split('<unnamed>::', 1)[-1].split('(')[0]
    us = float(row['Metric Value']) / 1e3; tot += us
    if 'score_tc' in short: sc += us
    A(f"| {row['ID']} | `{short}` | {row['Grid Size']} x {row['Block Size']} | {us:.1f} |")
A(f"\nScoring kernel share of the search under ncu: {sc/tot*100:.1f}% ({sc:.0f} of {tot:.0f} us); from bench.py's CUDA events without profiler: {d['roofline']['score_kernel_ms_per_search']/d['ms_per_step']*100:.1f}%. The shares agree.\n")
def table(path, title, algo_rows, elt=1536):
    rows = list(csv.reader(open(P / path))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    A(f"## {title}\n")
    A("| segment | rows | ms | DRAM read GB (algorithmic GB) | DRAM write MB | DRAM % of peak | tensor pipe active % | L2 throughput % | L2 hit % | SM clock GHz |\n|---|---|---|---|---|---|---|---|---|---|")
    for j, rr in enumerate(rows[2:]):
        g = lambda k: rr[idx[k]]
        nrows = algo_rows[j] if j < len(algo_rows) else 0
        A(f"| {j} | {nrows} | {float(g('gpu__time_duration.sum')):.4f} | {float(g('dram__bytes_read.sum')):.3f} ({nrows*elt/1e9:.3f}) | {float(g('dram__bytes_write.sum')):.1f} | {float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(g('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(g('lts__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(g('lts__t_sector_hit_rate.pct')):.1f} | {float(g('sm__cycles_elapsed.avg.per_second')):.3f} |")
    A("")
table('r01m_score_tc64_ncu_raw.csv', "ncu --set full, `score_tc_kernel<64,1,1>` (64 queries), the 4 segments of one search", [4096, 83840, 1800960, 8111104])
A("DRAM traffic equals the algorithmic bytes (15.36 GB read per search, +0.1%): nothing is re-read, the score matrix never exists. The two large segments run at 6.9-7.0 TB/s (84-85% of ncu's DRAM peak, 106% of the copy-kernel figure) with the tensor pipe 16-20% busy: HBM-bound as designed. 128 registers, 1 CTA/SM, 192 threads.\n")
table('r01m_score_tc2_pair_ncu_raw.csv', "ncu --set full, `score_tc2_kernel` (2-CTA pairs, 8192 queries), the 7 segments of one search", [4096, 12288, 49152, 196608, 786432, 3145728, 5805696])
A("The large segments keep the tensor pipe 83-87% active at 1.41-1.44 GHz (sw_power_cap): the kernel sits at the MMA issue limit at the clock the 1 kW budget allows. Against the 1-CTA capture (`r01e_score_tc256_ncu_raw.csv`: 86-90% active at 1.38 GHz, L2 throughput 67-72%) the pair kernel moves 1/3 less data per flop (L2 throughput 50%), which buys the higher clock. Early segments (dense survivors, few items per pair) are below that; they cover 10% of the rows.\n")
A("## BASELINE configs[2] at full size: 100M x 768 fp16 row-sharded over 2/4/8 B200, top-1000, fused peer-memory exchange (r01k_bench_c3_n*_p2p.json)\n")
A("The north-star target configuration (>= 80% of the HBM roofline at 64-query batches, >= 60% of the bf16 tensor roofline at 8192-query batches). `python -m torch.distributed.run --nproc-per-node N bench.py --gpus N --rows 100000000 --top-k 1000 --store-dtype float16`.\n")
A("| GPUs | 64-query: queries/s (ms/step) | aggregate corpus GB/s | scoring kernels, % of measured HBM peak | whole step, % | p50 latency ms | 8192-query: queries/s (ms/step) | scoring kernels TFLOP/s per GPU (% of burst / sustained peak) | whole step, % of burst |\n|---|---|---|---|---|---|---|---|---|")
for n in (2, 4, 8):
    x = bench(f"r01k_bench_c3_n{n}_p2p.json"); r = x['roofline']; L = x['large_batch']; R = L['roofline']
    A(f"| {n} | {x['value']:.0f} ({x['ms_per_step']:.3f}) | {x['corpus_gb_per_s']:.0f} | {r['frac']*100:.1f} | {r['whole_step_frac']*100:.1f} | {x['latency']['p50']:.3f} | {L['value']:.0f} ({L['ms_per_step']:.1f}) | {R['achieved']:.0f} ({R['frac']*100:.1f} / {R['frac_of_sustained']*100:.1f}) | {R['whole_step_frac']*100:.1f} |")
A("\nPer-kernel times come from a separate pass with events around every launch (idle gaps between kernels); the whole-step numbers are the back-to-back timed region, where the 1 kW power cap holds the clocks lower (`sw_power_cap` in every run) — that, the selects (k = 1000: 4-8 ms per 8192-query batch) and the merge are the difference between the two columns.\n")
lp = json.loads((P / 'r01k_latency_probe.json').read_text())
A("## Isolated-call latency vs submission pattern (r01k_latency_probe.json, scripts/latency_probe.py; 10M x 768 bf16, 64 queries)\n")
A("| pattern | ms per search (p10 / p50 / p90) |\n|---|---|")
A(f"| 40 searches back to back, one event pair | {lp['back_to_back_ms']:.3f} (mean) |")
for key, label in (("isolated", "one search per synchronize, no pause"), ("isolated_gap_5ms", "one search per synchronize, 5 ms idle between calls"), ("isolated_gap_50ms", "one search per synchronize, 50 ms idle between calls"), ("pairs", "two searches per synchronize"), ("quads", "four searches per synchronize")):
    c = lp[key]['call_ms']; A(f"| {label} | {c['p10']:.3f} / {c['p50']:.3f} / {c['p90']:.3f} |")
A(f"| host time to enqueue one search (9 launches) | {lp['host_enqueue_ms']['p50']*1e3:.0f} us |")
A("\nThe same kernels take 2.19 ms (7.0 TB/s) when the GPU idles a few ms between searches - the situation inside a training loop - and 2.30-2.45 ms under a saturating stream: the step time is set by the board's power / thermal management of HBM + tensor work, not by launch structure (enqueue is 36 us, selects 60 us).\n")
rows = list(csv.reader(open(P / 'r01k_aux_kernels_ncu_raw.csv'))); hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
A("## Kernels beside the tensor-core scan (ncu --set full, r01k_aux_kernels_ncu_raw.csv, scripts/ncu_aux_probe.py)\n")
A("| kernel | grid x block | us | DRAM read MB | regs | warps active % | FMA pipe active % | what it ran |\n|---|---|---|---|---|---|---|---|")
what = {"score_exact_kernel": "fp32-exact scan, 64 queries over 1M x 768 fp32 (3 segments)", "select_kernel": "radix select + sort of the candidate lists (64 x k=100, then 32 x k=1000)",
        "match_labels_kernel": "32 x 1000 ids vs 2 gold ids", "sample_kernel": "labeled priority sampling, 32 x 1000 -> 8", "gather_picks_kernel": "take_along_axis + max_sampling_id",
        "merge_kernel": "8 shards x 64 queries x 100 -> 100", "merge_results_kernel": "hybrid merge, 32 rows, 4 + 1000 + 1000 entries"}
seen = {}
for rr in rows[2:]:
    nm = rr[ix['Kernel Name']].split('unnamed>::', 1)[-1].split('(')[0]; base = nm.split('<')[0]
    key = (base, rr[ix['Grid Size']])
    if key in seen: continue
    seen[key] = 1
    g = lambda k: float(rr[ix[k]])
    A(f"| `{nm}` | {rr[ix['Grid Size']]} x {rr[ix['Block Size']]} | {g('gpu__time_duration.sum')*1e3:.1f} | {g('dram__bytes_read.sum')*1e3:.2f} | {int(g('launch__registers_per_thread'))} | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | {what.get(base, '')} |")
A("\nThe fp32-exact CUDA-core scan of this capture (first version of the kernel) runs the big segment at 34 TFLOP/s (FMA pipe 49% active, LSU pipe 69% busy with shared-memory wavefronts, 4-way conflicts on the transposing stores); the current kernel (8x8 thread tile, swizzled stores) is 18% faster at 64 queries. Still compute-bound at ~15% of HBM bandwidth, which is why a bf16/fp16 store with 3-term queries on the tensor cores (`tensor3`, same 1e-5 parity) is the recommended exact mode. Everything else is a latency-bound single wave (6-45 us).\n")
c5 = json.loads((P / 'r01m_c5_n8_probe.json').read_text())
A("## BASELINE configs[4] at full size: 50M x 1024 index refresh + fp32-exact search on 8 GPUs (r01m_c5_n8_probe.json, scripts/probe_c5_multi.py)\n")
A("| step | result |\n|---|---|")
A(f"| ingest: float32 vectors in pinned host memory (2^18-row batches) -> bf16 row shards, 8 ranks in parallel | {c5['ingest_s']:.2f} s for 50M x 1024 = {c5['ingest_host_GBps_aggregate']:.0f} GB/s of host data in aggregate ({c5['ingest_rows_per_s']/1e6:.0f}M rows/s; one GPU alone: 52 GB/s, so 8 concurrent uploads share the host's memory / PCIe fabric) |")
A(f"| search, 64 float32 queries, top-100, `tensor3` (fp32-exact on tensor cores), fused peer-memory exchange | {c5['search_tensor3_ms']:.3f} ms per batch = {c5['search_tensor3_qps']:.0f} queries/s, {c5['search_tensor3_GBps_aggregate']/1e3:.1f} TB/s scanned in aggregate; recall vs the CUDA-core fp32 kernel {c5['tensor3_recall_vs_cuda_core_exact']:.4f} |")
A(f"| same, `tensor` (queries rounded to bf16) | {c5['search_tensor_ms']:.3f} ms = {c5['search_tensor_qps']:.0f} queries/s, {c5['search_tensor_GBps_aggregate']/1e3:.1f} TB/s |")
ps = [json.loads(l) for l in (P / 'r01m_pitch_sweep.jsonl').read_text().splitlines() if l.startswith('{')]
A("\nRow length vs scan bandwidth (r01m_pitch_sweep.jsonl, scripts/sweep_pitch.py; bf16, 64 queries, `tensor`): dense pitch / one extra unscanned 64-element chunk per row, GB/s:\n")
A("| dim | store GB | dense | padded |\n|---|---|---|---|")
for i in range(0, len(ps) - 1, 2):
    d0, d1 = ps[i], ps[i + 1]
    A(f"| {d0['dim']} | {d0['rows']*d0['dim']*2/1e9:.1f} | {d0['GBps']:.0f} | {d1['GBps']:.0f} |")
A("\nNo penalty for power-of-two row strides (dim 1024 / 2048), and padding never helps: the layout stays dense. (Single runs of the config-4 probe on one shard came out at 1.9 or 2.2-2.3 ms with either layout — run-to-run state of the box, not the stride.)\n")
sw = [json.loads(l) for l in (P / 'r01k_batch_sweep.jsonl').read_text().splitlines() if l.startswith('{')]
A("## Roofline curve over the batch size (r01k_batch_sweep.jsonl, scripts/sweep_batch.py; 10M x 768 bf16, top-100, back-to-back searches)\n")
A("| queries | ms | queries/s | roofline ms = max(bytes / HBM peak, flops / bf16 burst peak) | fraction | bound | segments |\n|---|---|---|---|---|---|---|")
for r in sw:
    A(f"| {r['nq']} | {r['ms']:.3f} | {r['qps']:.0f} | {r['roofline_ms']:.3f} | {r['frac']*100:.0f}% | {r['bound']} | {r['segments']} |")
A("\nQuery tiles are 64 / 128 (1-CTA kernel) or 256 wide (CTA pair), so 257-511 queries cost what 512 cost; around the crossover (128-512 queries) HBM and tensor pipe are both near their limits and share one power budget, which is where the fraction of the max() roofline is lowest. Run-to-run spread on one box at 384-512 queries was +-15% (5.8-7.1 ms at 512) with the clocks under `sw_power_cap`.\n")
fp = json.loads((P / 'r01k_fp32_planes_probe.json').read_text())
A("## float32 store: CUDA-core exact kernel vs tensor cores over bf16 planes (r01k_fp32_planes_probe.json, scripts/probe_fp32_planes.py; 4M x 768 fp32 = 12.3 GB, top-100)\n")
A("| queries | exact (CUDA cores) ms | tensor3 ms (speed-up; GB/s of fp32 bytes; recall vs exact) | tensor2 ms (recall) | tensor ms (recall) |\n|---|---|---|---|---|")
for nq in (64, 256, 1024):
    e = fp[f'q{nq}_exact_ms']; t3 = fp[f'q{nq}_tensor3_ms']
    A(f"| {nq} | {e:.2f} | {t3:.2f} ({e/t3:.1f}x; {fp[f'q{nq}_tensor3_GBps_of_fp32_bytes']:.0f}; {fp[f'q{nq}_tensor3_recall_vs_exact']:.5f}) | {fp[f'q{nq}_tensor2_ms']:.2f} ({fp[f'q{nq}_tensor2_recall_vs_exact']:.5f}) | {fp[f'q{nq}_tensor_ms']:.2f} ({fp[f'q{nq}_tensor_recall_vs_exact']:.4f}) |")
A("\ntensor3 reads 6 bytes per stored element (three bf16 planes); up to 64 queries one MMA per plane and K step covers all query terms (N = 192 / 128 / 64), and the search runs at about 90% of the HBM time of those 18.4 GB. Scores agree with a float64 re-score to <= 4e-6 relative (leading product and corrections are accumulated in separate TMEM column groups).\n")
A("## Epilogue history (8192-query batch, segment with ~19 survivors per 128x256 item)\n")
A("| version | that segment | whole batch |\n|---|---|---|")
A("| r01a: warp-aggregated global atomic per surviving column, L2 round trip inside the epilogue | 93 ms, 20% of the MMA rate | 190.5 ms, 39.7% of peak |")
A("| r01b: dump mode for segment 0, CTA-level smem staging + bulk flush behind block barriers | 60 ms | 141.3 ms, 52.7% |")
A("| + two-phase flush (issue atomics, consume after the item), register bitmask push | 50 ms | 134.2 ms, 58.3% |")
A("| r01c: per-warp staging (no block barriers), select-tree push, growth 3 for large batches | not distinguishable | 98.5 ms, 77.7% |")
A("| r01g/h: 2-CTA pair kernel, PDL | - | 95.3-98.4 ms, 78-80% of burst / 93-95% of sustained peak |\n")
A("ncu source-level sampling that drove this (r01b capture, epilogue warps): 15% of samples on the unrolled per-column bit tests, 15% at the block barrier waiting for warps in the push path, 11% instruction-cache misses in the unrolled push code.\n")
t = json.loads((P / 'r01_query_terms_probe.json').read_text().strip().splitlines()[-1])
A("## Query-term modes (r01_query_terms_probe.json; 10M x 768 bf16, float32 queries NOT representable in bf16)\n")
A("| mode | 64-query ms | 8192-query ms | recall@100 vs fp32 CUDA-core kernel |\n|---|---|---|---|")
for m in ('tensor', 'tensor2', 'tensor3'):
    A(f"| {m} | {t[m+'_q64_ms']:.3f} | {t[m+'_q8192_ms']:.1f} | {t[m+'_recall_vs_exact']:.4f} |")
A(f"| exact (fp32 FMA, CUDA cores) | {t['exact_q64_ms']:.2f} | - | 1 |")
A(f"| tensor3, float32 queries that are exact in bf16 (empty correction terms skipped on the device) | {t['tensor3_q64_bf16_exact_queries_ms']:.3f} (tensor on the same batch: {t['tensor_q64_bf16_exact_queries_ms']:.3f}; identical results: {t['tensor3_equals_tensor_on_bf16_exact_queries']}) | - | - |\n")
c = json.loads((P / 'r01_configs_3_5_probe.json').read_text().strip().splitlines()[-1])
A("## BASELINE configs 3 and 5, one GPU's share (r01_configs_3_5_probe.json, scripts/probe_configs.py)\n")
A("| config | result |\n|---|---|")
A(f"| C3 shard: 12.5M x 768 fp16 (= 100M rows over 8 GPUs), top-1000, 64 queries | {c['c3_q64_k1000_ms']:.3f} ms per batch = {c['c3_q64_GBps']:.0f} GB/s ({c['c3_q64_GBps']/65.348:.0f}% of measured HBM peak), {c['c3_q64_segments']} segments, list capacity {c['c3_q64_cap']} |")
A(f"| C3 shard, 8192 queries, top-1000 | {c['c3_q8192_k1000_ms']:.1f} ms = {c['c3_q8192_TFLOPs']:.0f} TFLOP/s ({c['c3_q8192_TFLOPs']/16.717:.0f}% of measured bf16 peak; 1-CTA kernel at the time) |")
A(f"| C5 shard: 6.25M x 1024 (= 50M rows over 8 GPUs), fp32 host vectors -> bf16 store, 2^18-row chunks from pinned memory | {c['c5_ingest_s']:.3f} s = {c['c5_ingest_host_GBps']:.1f} GB/s of host data (PCIe Gen5 x16 bound); from a CUDA fp32 tensor (encoder output): {c['c5_ingest_from_device_s']*1e3:.1f} ms |")
A(f"| C5 search, fp32-exact, 64 queries, top-100 | tensor cores, 3 query terms: {c['c5_search_tensor3_ms']:.2f} ms; CUDA-core fp32 kernel: {c['c5_search_exact_ms']:.1f} ms (same neighbours: recall {c['c5_tensor3_vs_exact_recall']:.1f}, max relative score difference {c['c5_tensor3_vs_exact_max_rel_score_diff']:.0e}); 1 term: {c['c5_search_tensor_ms']:.2f} ms |\n")
A("## compute-sanitizer\n\n`scripts/sanitizer_probe.py` (every kernel, small sizes): memcheck 0 errors, racecheck 0 hazards (`r01_compute_sanitizer.txt`).")
(P / 'r01_summary.md').write_text("\n".join(out) + "\n")
print("wrote", P / 'r01_summary.md')
