// common.cuh — shared declarations of libvodb.so (internal; the public ABI is include/vodb.h)
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <mutex>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/vodb.h"
#include "vodb_math.h"

namespace vodb {

// ---- errors -----------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define VODB_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::vodb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VODB_ECUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define VODB_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::vodb::set_error(__VA_ARGS__);    \
      return VODB_EINVAL;                \
    }                                    \
  } while (0)

// ---- element types ------------------------------------------------------------
__host__ __device__ inline int dtype_size(int dtype) { return dtype == VODB_F32 ? 4 : 2; }

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// order-preserving float -> uint32 (bigger float = bigger uint); NaN lowest; -0 == +0
__host__ __device__ __forceinline__ uint32_t ord_u32(float x) {
  uint32_t u = vm_f2u(x);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0u;
  if ((u & 0x7fffffffu) == 0u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return vm_u2f(u);
}

#define VODB_NEG_FLT_MAX (-3.402823466e+38f)

// ---- programmatic dependent launch (PDL) ---------------------------------------------
// Every kernel of the search chain is launched with programmaticStreamSerializationAllowed: its CTAs may be
// scheduled while the previous kernel of the stream is still draining, so launch latency and per-CTA set-up
// (barrier init, TMEM allocation, descriptor prefetch) overlap the predecessor's tail. A kernel calls
// pdl_launch_dependents() as early as possible and pdl_wait() before it touches anything a predecessor wrote.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember what has been set for
// (kernel, current device) so that stores on several GPUs of one process all get it (api.cu)
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);

// ---- corpus store ---------------------------------------------------------------
constexpr int kPitchAlign = 64;  // elements; one 128-byte TMA swizzle atom of 2-byte types

struct Workspace {
  // per-query candidate lists: scores/ids [nq_cap, cap], counts, thresholds
  float* cand_s = nullptr;
  int32_t* cand_i = nullptr;
  int* cnt = nullptr;
  float* tau = nullptr;
  int* term_any = nullptr;     // [kTermSlots] per prepare-block bit mask: bit t set = query term t has a nonzero element
  int* overflow = nullptr;     // device flags: [0] a candidate list of THIS shard ran out of room (sticky);
                               // [1] some shard of a sharded search did (OR over the ranks, set by the merge kernel)
  int* overflow_host = nullptr;  // pinned mirror
  int nq_cap = 0;        // counters / thresholds allocated for this many queries
  size_t list_elems = 0; // candidate slots allocated in total (>= nq * cap of the current call)
  int cap = 0;           // per-query list capacity (= row stride) of the current call
  // staged queries (converted / padded) and staged outputs for host callers
  void* q_stage = nullptr;
  size_t q_stage_bytes = 0;
  void* q_in = nullptr;  // raw copy of host queries
  size_t q_in_bytes = 0;
  // staged outputs of host callers: one device buffer [ids i64 x nqk | scores f32 x nqk | overflow flag] so that
  // results and flag come back with ONE device->host copy into the pinned mirror
  char* out_pack = nullptr;
  char* out_host = nullptr;  // pinned
  size_t out_cap = 0;        // elements (nq * k) the pack can hold
  // retrieve -> sample chain (vodb_retrieve_sample): labels / gold ids / sampler outputs / packed result + pinned mirror
  char* chain_dev = nullptr;
  char* chain_host = nullptr;  // pinned
  size_t chain_bytes = 0;
};

}  // namespace vodb

struct vodb_store {
  // held by every entry point that touches the store: workspace, staging buffers, tensor-map cache, planes and
  // statistics are per store, so host threads take turns (the kernels of concurrent callers additionally share the
  // candidate lists: callers that overlap searches on one store must enqueue them on the same stream)
  std::mutex mu;
  int device = 0;
  int64_t n_rows = 0;      // capacity
  int64_t n_added = 0;     // rows filled so far
  int dim = 0;
  int pitch = 0;           // elements per stored row (dim rounded up to kPitchAlign)
  int dtype = VODB_F32;
  int64_t row_offset = 0;
  void* data = nullptr;    // [n_rows, pitch] of dtype, zero padded
  int sm_count = 0;
  vodb::Workspace ws;
  // host->device adds: two staging buffers; chunk i+1 is copied on `copy_stream` while chunk i is converted on the
  // caller's stream (events order buffer reuse), so PCIe and the conversion kernel overlap
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr};     // chunk landed in stage[b]
  cudaEvent_t converted[2] = {nullptr, nullptr};  // stage[b] consumed by the conversion kernel
  // TMA descriptor cache for the tensor-core path (opaque CUtensorMap storage)
  alignas(64) unsigned char tmap_corpus[128];
  bool tmap_corpus_valid = false;
  // fp32 stores on the tensor cores: bf16 planes [3][n_rows][pitch] with data = p0 + p1 + p2 exactly, built on the
  // first tensor-mode search (6 more bytes per element) and extended after later adds
  void* planes = nullptr;
  int64_t planes_rows_done = 0;  // rows [0, planes_rows_done) of the planes are up to date
  alignas(64) unsigned char tmap_planes[128];
  bool tmap_planes_valid = false;
  int64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // stream of the previous call that touched the workspace: a call arriving on another stream first waits for
  // everything enqueued so far (the lists / staging buffers are shared), api.cu order_streams
  cudaStream_t last_stream = nullptr;
  bool last_stream_set = false;
  // optional per-kernel timing (vodb_store_set_profiling): events recorded around every scan launch
  bool profiling = false;
  void* prof = nullptr;  // vodb::ProfileState*
};

namespace vodb {

// ---- kernels launched by the search pipeline (defined in the .cu files) ----------
struct SegmentArgs {
  const void* corpus;   // [n_rows, pitch]; fp32 store in tensor mode: bf16 planes [3][n_rows][pitch]
  int dtype;            // dtype of what the scoring kernel reads (store dtype, or bf16 for the planes)
  int pitch;
  int64_t row_begin;    // segment [row_begin, row_end) of local rows
  int64_t row_end;
  const void* queries;  // EXACT: float32 [nq, pitch]; TENSOR: store dtype [nq_pad, pitch]
  int nq;
  float* cand_s;
  int32_t* cand_i;
  int* cnt;
  const float* tau;
  int* overflow;
  int cap;
  int terms;            // TENSOR: 16-bit terms per query (1..3), staged as [terms][round_up(nq,256)][pitch]
  int planes;           // TENSOR on an fp32 store: bf16 corpus planes used (= terms); 1 otherwise
  const int* term_any;  // prepare kernel's per-block term masks (which query terms are nonzero at all), term_blocks of them
  int term_blocks;
  bool dump;            // first segment of a scan: lists are empty, every score is stored at slot row-row_begin
};

// Cross-shard exchange fused into the final select (select.cu), NCCL-LL style: every rank's final select kernel
// stores its [nq,k] result straight into slot `rank` of every peer's gather buffer (peer-mapped memory, NVLink
// stores) as three 8-byte words per entry, each carrying 4 payload bytes and the 4-byte epoch tag of the call:
//   w0 = epoch<<32 | score bits, w1 = epoch<<32 | id[31:0], w2 = epoch<<32 | id[63:32]   (one plane per word).
// An aligned 8-byte store is single-copy atomic, so a reader that sees the tag also sees the payload: no fences, no
// flags, no counters. The merge kernel spins on the tags of the entries it needs and reduces world*k -> k.
// Per-query candidate counters live kCntStride ints (256 bytes) apart. Atomics on addresses in one 128-byte line are
// serialised by the same L2 atomic unit (and adjacent lines pair up through address bit 7, B300_MICROARCH.md "L2-atom
// multi-CTA"), so 64 densely packed counters behave like two addresses: the 84k-row segment of a 64-query scan over a
// 1.25M-row shard (2000 appends per query) took 49 us for 19 us worth of HBM traffic, its epilogue warps waiting on
// atomic round trips (ncu source view, profiles/r02_summary.md).
constexpr int kCntStride = 64;
constexpr int kTermSlots = 148 * 32;  // upper bound of the prepare kernel's grid (api.cu grid_for)
constexpr int kMaxPeers = 16;
struct ExchangeDst {
  int world;
  int rank;
  uint32_t epoch;
  // peer r's gather buffer, slot `rank`, current parity: three planes [3][slot_elems] of tagged words (score, id low,
  // id high) — plane-major so that a warp's 32 stores of one word are 256 contiguous bytes on the NVLink write path
  uint64_t* peer_ll[kMaxPeers];
  size_t slot_elems;             // entries per plane (max_nq * max_k of the exchange)
  // one more tagged word per (parity, source rank) behind the entries: epoch<<32 | this shard's overflow flag, so that
  // every rank learns from the exchange itself whether ANY shard overflowed (all ranks then re-run in lockstep)
  uint64_t* peer_flag[kMaxPeers];
  const int* overflow;           // this shard's sticky flag (Workspace::overflow[0])
};

int launch_score_exact(const SegmentArgs& a, int sm_count, cudaStream_t stream);
int launch_score_tensor(vodb_store* s, const SegmentArgs& a, cudaStream_t stream);
bool tensor_path_supported(const vodb_store* s);

// select the k best candidates of every query list; if `final`, sort and write outputs
int launch_select(float* cand_s, int32_t* cand_i, int* cnt, float* tau, int cap, int nq, int k, bool final,
                  float* out_s, int64_t* out_i, int64_t row_offset, cudaStream_t stream,
                  const ExchangeDst* xd = nullptr, int expected_n = 0);
// merge of the gathered per-rank lists; waits for every entry's epoch tag (see ExchangeDst)
// slot_words = words per (parity, source rank) slot (entries + the flag word at offset flag_word)
int launch_merge_exchange(const uint64_t* gather_ll, uint32_t epoch, int world, size_t slot_words, size_t flag_word, int nq,
                          int k, float* out_s, int64_t* out_i, int* overflow_any, cudaStream_t stream);
int launch_merge(const float* scores, const int64_t* idx, int n_lists, int nq, int k_in, int k_out, float* out_s,
                 int64_t* out_i, cudaStream_t stream);

int launch_convert_rows(const void* src, int src_dtype, int src_dim, void* dst, int dst_dtype, int dst_pitch,
                        int64_t n, cudaStream_t stream);
// first kernel of a search: stage the queries (fp32 plane for EXACT, `terms` 16-bit planes for TENSOR) and reset the
// candidate lists (cnt = rows of the first segment, tau = -inf) in one launch
// `term_any[b]` receives block b's mask of nonzero terms; returns the number of blocks launched in *term_blocks
int launch_prepare(const void* src, int src_dtype, int src_dim, void* dst, int dst_dtype, int dst_pitch, int64_t n,
                   int64_t plane_rows, int terms, int* cnt, float* tau, int first_rows, int* term_any, int* term_blocks,
                   cudaStream_t stream);
int launch_split_planes(const float* src, void* planes, int64_t row0, int64_t n, int pitch, int64_t n_rows,
                        cudaStream_t stream);
int launch_fill_synthetic(void* dst, int dtype, int dim, int pitch, uint64_t seed, int64_t global_row0, int64_t n,
                          int unit_norm, cudaStream_t stream);
int launch_read_rows(const void* src, int dtype, int dim, int pitch, int64_t n, float* out, cudaStream_t stream);

int launch_merge_results(int n_engines, const void* const* scores, const int64_t* const* indices,
                         const int64_t* const* labels, const int* widths, const double* weights, const int* zero_scores,
                         int B, int is_f64, int normalize, double offset, int label_engine, int out_width,
                         void* out_scores, int64_t* out_indices, int64_t* out_labels, void* out_raw, int* out_counts,
                         cudaStream_t stream);

int launch_sample(const float* scores, const uint8_t* labels, const float* noise, int B, int K, int k_positive,
                  int k_total, int normalized, float temperature, int max_support, int quirks, uint64_t seed,
                  uint64_t offset, int64_t* out_ids, float* out_logw, uint8_t* out_labels, float* out_lse,
                  cudaStream_t stream);

int launch_match_labels(const int64_t* idx, const int64_t* gold, int n_gold, int B, int K, uint8_t* labels,
                        cudaStream_t stream);
int launch_gather_picks(const float* scores, const int64_t* idx, const uint8_t* labels, int B, int K, int k_total,
                        const int64_t* local, const uint8_t* picked_labels, int64_t* out_idx, float* out_scores,
                        float* out_msid, cudaStream_t stream);

}  // namespace vodb
