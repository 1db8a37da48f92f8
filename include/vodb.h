/* vodb.h — C ABI of libvodb.so, the B200-native dense-retrieval hot path.
 *
 * This is the drop-in boundary for VOD's dynamic-retrieval path. Every entry
 * point replaces one call the reference makes into third-party native code
 * (faiss C++ / numba-JIT numpy); citations are relative to the reference tree.
 *
 *   vodb_store_create / _add        <- faiss.index_factory(D,"Flat",IP) + index.add(f32 rows)
 *                                      src/vod_search/faiss_search/build.py:60,67-73
 *   vodb_store_add (src_on_device)  <- the faiss-gpu sharded add_with_ids loop
 *                                      src/vod_search/faiss_search/build_gpu.py:294-380
 *   vodb_search                     <- faiss_index.search(query_vec, k)
 *                                      src/vod_search/faiss_search/server.py:72,84
 *   vodb_merge_topk                 <- the host-side IndexShards merge behind
 *                                      faiss.index_cpu_to_all_gpus(index, co.shard=True)
 *                                      src/vod_search/faiss_search/server.py:51-54
 *   vodb_sample                     <- _labeled_priority_sampling_2d_ (numba)
 *                                      src/vod_dataloaders/core/sample.py:323-352 (and :245-320, :160-219)
 *   vodb_retrieve_sample            <- RealmCollate: search -> sample_search_results, dense-only flow
 *                                      src/vod_dataloaders/realm_collate.py:101-122, core/sample.py:22-84
 *   vodb_store_ntotal / _dim        <- index.ntotal / index.d checks, build.py:75-79; server.py:59-66
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types cross this boundary;
 *   - every function returns 0 on success or a negative VODB_E* code; the message
 *     is available from vodb_last_error() (thread-local);
 *   - "on_device" flags say whether a pointer is a CUDA device pointer on the
 *     store's device (1) or host memory (0). Host buffers are copied inside the
 *     call (H2D of queries, D2H of results) and the call returns after the
 *     results have landed; with device buffers the work is only enqueued on
 *     `stream` (a cudaStream_t passed as void*, NULL = legacy default stream);
 *   - one store = one row shard on one GPU. Multi-GPU = one process per GPU,
 *     each owning one store with its `row_offset`; per-shard results are
 *     exchanged by the host code (NCCL all-gather) and reduced by vodb_merge_topk;
 *   - every call on a store takes the store's mutex, so host threads take turns; searches that are enqueued
 *     asynchronously (out_on_device) share the store's candidate lists: calls on one stream are ordered by the
 *     stream, and a call that arrives on a different stream than its predecessor first waits for the device to
 *     drain (correct on any mix of streams; keep one stream per store for full asynchrony).
 */
#ifndef VODB_H_
#define VODB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VODB_ABI_VERSION 1

/* element types of stored rows / queries */
#define VODB_F32 0
#define VODB_BF16 1
#define VODB_F16 2

/* scoring modes of vodb_search */
#define VODB_MODE_EXACT 0  /* fp32 FMA on CUDA cores over the stored values: IndexFlatIP parity mode */
#define VODB_MODE_TENSOR 1 /* tcgen05 tensor cores (bf16/fp16 store; queries rounded to the store dtype) */
#define VODB_MODE_TENSOR_X2 2 /* same, float32 queries split into 2 store-dtype terms (~16 mantissa bits kept) */
#define VODB_MODE_TENSOR_X3 3 /* same, 3 terms: the full float32 query mantissa; products are exact in fp32 */
/* The tensor modes also serve a float32 store: its rows are then mirrored as three bf16 planes (row = p0 + p1 + p2
 * exactly; 6 more bytes per element, allocated by the first such search, kept in step with later adds) and mode X
 * multiplies the first X planes with X query terms. TENSOR_X3 reproduces the fp32 result within the fp32-exact
 * tolerance (1e-5 relative; measured <= 4e-6) 3-4x faster than VODB_MODE_EXACT.
 * Correction terms that are zero for the whole query batch (float32 queries that are exact in the store dtype) are
 * detected on the device and skipped: TENSOR_X2 / _X3 then cost what TENSOR costs and return the same bits.
 * Host-resident float32 queries are classified on the host (a few microseconds): when every value is representable in
 * the 16-bit store dtype the call runs as VODB_MODE_TENSOR outright, on the one-term kernels. */

/* error codes */
#define VODB_OK 0
#define VODB_EINVAL (-1)   /* bad argument */
#define VODB_ECUDA (-2)    /* CUDA runtime / driver error */
#define VODB_ENOMEM (-3)   /* allocation failed */
#define VODB_ESTATE (-4)   /* call not valid in this state (e.g. search on an empty store) */
#define VODB_EUNSUPPORTED (-5)

/* sampling quirk bits (vodb_sample `quirks`) */
#define VODB_QUIRK_INVERTED_SUPPORT 1 /* reproduce sample.py:176-178 literally: the top
                                          `max_support` entries are masked OUT */

#define VODB_MAX_K 2048

typedef struct vodb_store vodb_store;

const char* vodb_last_error(void);
int vodb_abi_version(void);
/* number of CUDA devices visible, or a negative error code */
int vodb_device_count(void);

/* ---- corpus store ------------------------------------------------------ */

/* Allocate an HBM-resident store for `n_rows` x `dim` embeddings of `dtype` on
 * CUDA device `device`. Global ids returned by searches are row_offset + local row. */
int vodb_store_create(vodb_store** out, int device, int64_t n_rows, int dim, int dtype,
                      int64_t row_offset);
void vodb_store_destroy(vodb_store* s);

/* Copy rows [row0, row0+n) into the store, converting src_dtype -> store dtype
 * (round-to-nearest-even) on the device. `rows` is row-major [n, dim]. Rows are appended or overwritten:
 * row0 <= vodb_store_ntotal() (faiss `index.add` appends); a block that would leave a gap is VODB_EINVAL. */
int vodb_store_add(vodb_store* s, const void* rows, int src_dtype, int src_on_device,
                   int64_t row0, int64_t n, void* stream);

/* Fill rows [row0, row0+n) with the deterministic synthetic embedding
 * vodb_synth_value(seed, global_row, col) (csrc/vodb_math.h), generated on the device.
 * If unit_norm != 0 rows are L2-normalised before rounding to the store dtype. */
int vodb_store_fill_synthetic(vodb_store* s, uint64_t seed, int64_t row0, int64_t n,
                              int unit_norm, void* stream);

/* Copy rows [row0,row0+n) back out as float32 (tests / verification). */
int vodb_store_read(vodb_store* s, int64_t row0, int64_t n, float* out, int out_on_device,
                    void* stream);

int64_t vodb_store_ntotal(const vodb_store* s); /* rows added so far (max row0+n seen) */
int vodb_store_dim(const vodb_store* s);
int vodb_store_dtype(const vodb_store* s);
int vodb_store_device(const vodb_store* s);
int64_t vodb_store_bytes(const vodb_store* s);  /* HBM bytes of the row data (pitch included) */

/* ---- exact MIPS top-k --------------------------------------------------- */

/* scores[q, j] = <queries[q], row(idx[q, j])>, the k largest per query, sorted by
 * (score descending, id ascending). Slots beyond the number of stored rows get
 * id -1 and score -FLT_MAX (what faiss returns). out_scores is float32 [nq, k],
 * out_idx is int64 [nq, k] (global ids). 1 <= k <= VODB_MAX_K. */
int vodb_search(vodb_store* s, const void* queries, int q_dtype, int q_on_device, int nq,
                int k, int mode, float* out_scores, int64_t* out_idx, int out_on_device,
                void* stream);

/* Make the store ready for the tensor-core modes now instead of inside the first such search: a float32 store gets
 * its three bf16 planes allocated (6 more bytes per element) and brought up to date with the rows added so far;
 * 16-bit stores need nothing. Returns VODB_ENOMEM when the planes do not fit (the store stays usable with
 * VODB_MODE_EXACT) — the host code's "auto" mode uses this to choose between TENSOR_X3 and EXACT for float32 stores. */
int vodb_store_prepare_tensor(vodb_store* s, void* stream);

/* With device outputs vodb_search only enqueues work. If a per-query candidate list overflowed (adversarial
 * row order / massive duplicate scores) a sticky device flag is set: this call synchronises `stream`, returns 1
 * if any search since the last check overflowed (results of those searches are invalid: re-run them with host
 * outputs, which falls back to the overflow-proof schedule), 0 if all were fine, or a negative error code.
 * After vodb_search_sharded the answer covers every rank's shard (see there). */
int vodb_search_check(vodb_store* s, void* stream);

/* Statistics of the last vodb_search on this store (for bench / tests):
 * out[0] = number of kernel launches, out[1] = number of scan segments,
 * out[2] = candidate-list capacity per query, out[3] = 1 if the safe fallback ran,
 * out[4] = max candidates seen in any list, out[5..7] reserved. */
int vodb_search_stats(const vodb_store* s, int64_t out[8]);

/* Per-kernel timing for bench.py's roofline: while enabled, CUDA events are recorded on the search stream around
 * every scoring and select launch. vodb_store_profile synchronises the device and returns, summed over the
 * searches since it was enabled / last read: out[0] = ms in scoring kernels, out[1] = ms in select kernels,
 * out[2] = number of scoring launches, out[3] = algorithmic corpus bytes those launches scanned (rows*dim*elt). */
int vodb_store_set_profiling(vodb_store* s, int enable);
int vodb_store_profile(vodb_store* s, double out[4]);

/* Merge `n_lists` per-shard results. scores: float32 [n_lists, nq, k_in], idx:
 * int64 [n_lists, nq, k_in] (global ids, -1 = empty slot). Writes the k_out best per
 * query in (score desc, id asc) order. All pointers on `device` if on_device. */
int vodb_merge_topk(int device, const float* scores, const int64_t* idx, int n_lists, int nq,
                    int k_in, int k_out, float* out_scores, int64_t* out_idx, int on_device,
                    void* stream);

/* ---- cross-shard exchange fused into the search (one process per GPU, peer-mapped buffers over NVLink) ------
 *
 * Replaces faiss' IndexShards host-side merge (index_cpu_to_all_gpus with co.shard=True, server.py:51-54) and the
 * NCCL all-gather of the unfused path: the final select kernel of every rank stores its [nq,k] (score, global id)
 * list directly into every peer's gather buffer as epoch-tagged 8-byte words (payload and tag in one atomic store,
 * no fence, no flag); the merge kernel spins on the tags of the entries it reads and reduces world*k -> k. Set-up: every rank calls vodb_xchg_create, the 64-byte handles are all-gathered
 * by the host code (torch.distributed), then every rank calls vodb_xchg_connect with the world*64 handle bytes
 * (rank-major). All ranks must then call vodb_search_sharded the same number of times, in the same order. */
#define VODB_IPC_HANDLE_BYTES 64
typedef struct vodb_xchg vodb_xchg;
int vodb_xchg_create(vodb_xchg** out, int device, int rank, int world, int max_nq, int max_k,
                     unsigned char* handle_out /* VODB_IPC_HANDLE_BYTES */);
int vodb_xchg_connect(vodb_xchg* x, const unsigned char* all_handles /* world * VODB_IPC_HANDLE_BYTES */);
void vodb_xchg_destroy(vodb_xchg* x);
/* Like vodb_search over this rank's shard, but out_scores / out_idx receive the MERGED result of all shards (every
 * rank gets the same [nq,k] arrays). Every rank's overflow flag travels with its list (one more tagged word), and the
 * merge kernel ORs them, so all ranks agree on whether ANY shard overflowed without a collective:
 *   - host outputs: the call itself re-runs the batch on the overflow-proof schedule, on all ranks in lockstep;
 *   - device outputs (enqueue only): vodb_search_check() returns the same answer on every rank; callers that see 1
 *     re-run the batch with safe=1 on ALL ranks.
 * `safe` != 0 selects the overflow-proof schedule from the start (all ranks must pass the same value).
 * The flag word also carries a fingerprint of (nq, k): ranks that run an epoch with different batch shapes do not
 * hang in the merge; the host-output call returns VODB_ESTATE, vodb_search_check() does after a device-output call. */
int vodb_search_sharded(vodb_store* s, vodb_xchg* x, const void* queries, int q_dtype, int q_on_device, int nq,
                        int k, int mode, int safe, float* out_scores, int64_t* out_idx, int out_on_device,
                        void* stream);

/* ---- hybrid merge: normalise + weighted union + raw-score / label gather (the step between search and sampling) --
 *
 * Replaces, per query row, `_subtract_min_score` (src/vod_dataloaders/core/normalize.py:17-20), `result * weight`
 * (src/vod_types/retrieval.py:222-233), the numba union `_nopy_merge_two_search_results` applied engine after
 * engine (src/vod_dataloaders/core/merge.py:31-62, 108-164) and `gather_values_by_indices`
 * (src/vod_dataloaders/core/numpy_ops.py:24-143) as `_merge_search_results` chains them (core/search.py:79-125).
 *   scores[e]   [B, widths[e]] of score_dtype (VODB_F32, or 3 = float64), indices[e] int64 (negative = padding),
 *   labels[e]   int64 or NULL; zero_scores[e] != 0 treats engine e's scores as 0 (the lookup engine)
 *   normalize   subtract each engine's row minimum over finite scores, add `offset`
 *   label_engine  engine whose labels are gathered into out_labels (-1: none; out_labels may be NULL)
 *   outputs     out_scores/out_indices/out_labels [B, out_width], out_raw [n_engines, B, out_width] (NaN where the
 *               engine did not return the id), out_counts[b] = number of distinct ids of row b. Slots past the
 *               count hold id -1 / score -inf. The caller keeps columns [0, min(out_width, max_b count + 1)).
 * Output order = first occurrence (engines in table order); duplicate ids are summed left to right, so results are
 * bit-identical to the reference for equal dtypes. sum(widths) <= 8192. */
int vodb_merge_results(int device, int n_engines, const void* const* scores, const int64_t* const* indices,
                       const int64_t* const* labels, const int* widths, const double* weights,
                       const int* zero_scores, int B, int score_dtype, int normalize, double offset,
                       int label_engine, int out_width, void* out_scores, int64_t* out_indices,
                       int64_t* out_labels, void* out_raw, int* out_counts, int on_device, void* stream);

/* The scan schedule vodb_search would use for a shard of n_rows rows (host logic only, no GPU needed): list capacity
 * per query in *out_cap and the segment boundaries b_0 = 0 < b_1 < ... = n_rows in out_bounds (at most max_bounds
 * written). Segment 0 is scored in dump mode, every later one against the threshold published by the select after
 * its predecessor; `safe` is the overflow-proof schedule used for the re-run. Returns the number of boundaries. */
int vodb_plan_scan(int64_t n_rows, int nq, int k, int safe, int* out_cap, int64_t* out_bounds, int max_bounds);

/* ---- labeled priority sampling ------------------------------------------ */

/* Per row b of scores[B,K]: split entries by labels[b,:] > 0, priority-sample
 * k_positive positives then (k_total - #positives) negatives, compute importance
 * log-weights (self-normalised per label group if `normalized`).
 *   noise      optional float32 [B,K] Exp(1) samples (what the reference draws from
 *              np.random at sample.py:398); NULL -> Philox-4x32-10(seed, offset, b, j)
 *   out_ids    int64  [B,k_total]  local column ids, -1 in unused slots
 *   out_logw   float32[B,k_total]  -inf in unused slots
 *   out_labels uint8  [B,k_total]
 *   out_lse    float32[B,2]        (pos, neg) log-normalisers (sample.py:305-307)
 * Results are bit-identical to oracle/sample_twin.c for equal inputs. K <= 8192. */
int vodb_sample(int device, const float* scores, const uint8_t* labels, const float* noise, int B,
                int K, int k_positive, int k_total, int normalized, float temperature,
                int max_support, int quirks, uint64_t seed, uint64_t offset, int64_t* out_ids,
                float* out_logw, uint8_t* out_labels, float* out_lse, int on_device, void* stream);

/* `sample_search_results` (src/vod_dataloaders/core/sample.py:22-84) on host arrays in one call: vodb_sample with
 * normalized=1 over scores[B,K] / labels[B,K] (uint8, NULL = no positives), then the gathers at the picks
 * (sample.py:57-64: out_idx = indices[b, local], out_scores = scores[b, local]; an unused slot, local = -1, reads the
 * LAST column like numpy's take_along_axis) and out_msid[b] = max_sampling_id (sample.py:66-71), all on the device.
 * out_local (optional) receives the sampled positions so that the caller can gather further per-engine raw scores.
 * Outputs equal vodb_retrieve_sample's for the same lists, bit for bit. 1 <= K <= 8192. */
int vodb_sample_results(int device, const float* scores, const int64_t* indices, const uint8_t* labels,
                        const float* noise /* as vodb_sample; NULL = Philox */, int B, int K, int k_positive, int k_total, float temperature, int max_support, int quirks,
                        uint64_t seed, uint64_t offset, int64_t* out_idx, float* out_scores, float* out_logw,
                        uint8_t* out_labels, float* out_lse, float* out_msid, int64_t* out_local, void* stream);

/* ---- retrieve -> sample chain --------------------------------------------- */

/* What RealmCollate does per training batch with the dense engine alone (realm_collate.py:101-122), in one call and
 * without the [nq, top_k] lists leaving HBM: vodb_search(top_k) -> labels[b,j] = (retrieved id in gold_ids[b,:]) ->
 * vodb_sample(k_positive, k_total, normalized=1) -> sample_search_results' gathers (core/sample.py:57-71).
 *   queries      as vodb_search (host or device pointer)
 *   gold_ids     host int64 [nq, n_gold] ids of the positive sections (negative entries = padding), NULL if n_gold == 0
 *   out_idx      host int64  [nq,k_total]  global row ids at the picks
 *   out_scores   host float32[nq,k_total]  retrieval scores at the picks
 *   out_logw / out_labels / out_lse        as vodb_sample
 *   out_msid     host float32[nq]          max_sampling_id (sample.py:66-71)
 *   out_local    host int64  [nq,k_total]  optional (NULL ok): positions inside the top_k list, -1 in unused slots
 * Unused slots (sampler position -1) gather the LAST retrieved column, like numpy's take_along_axis in the reference.
 * Results equal vodb_search + host labels + vodb_sample + numpy gathers bit for bit. Synchronous; one device->host
 * copy of the packed [nq, k_total] result. top_k <= VODB_MAX_K. */
int vodb_retrieve_sample(vodb_store* store, const void* queries, int q_dtype, int q_on_device, int nq, int top_k,
                         int mode, const int64_t* gold_ids, int n_gold, int k_positive, int k_total,
                         float temperature, int max_support, int quirks, uint64_t seed, uint64_t offset,
                         int64_t* out_idx, float* out_scores, float* out_logw, uint8_t* out_labels, float* out_lse,
                         float* out_msid, int64_t* out_local, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VODB_H_ */
