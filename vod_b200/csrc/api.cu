// api.cu — C ABI of libvodb.so (include/vodb.h): corpus store, search pipeline, merge, sampling.
//
// Search pipeline (replaces faiss `IndexFlatIP.search`, reference src/vod_search/faiss_search/server.py:84):
//   the shard is scanned in a few geometrically growing row segments. Segment 0 appends every score to the
//   per-query candidate lists (tau = -inf); after each segment `select` keeps the k best and publishes
//   tau[q] = k-th best so far, so later segments append only scores >= tau[q] — for rows in random order the
//   expected number of survivors of a segment is k * rows(segment) / rows(before), i.e. a vanishing fraction
//   of the scores, and the [nq x rows] score matrix never exists in HBM. A list that runs out of room sets a
//   device flag; the call then re-runs in "safe" mode (segments of cap-k rows, which cannot overflow).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace vodb {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}
const char* get_error() { return g_error.c_str(); }

cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = done[{kernel, dev}];
  if (bytes > cur) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    cur = bytes;
  }
  return cudaSuccess;
}

struct ProfileState {
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  // (begin, end, kind) triples; kind 0 = scoring kernel, 1 = select kernel
  std::vector<int> kinds;
  double bytes = 0.0;  // algorithmic corpus bytes of the recorded scoring launches
  cudaEvent_t next() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  void reset() {
    used = 0;
    kinds.clear();
    bytes = 0.0;
  }
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ---- element conversion / synthetic fill kernels -----------------------------------------------

template <typename S, typename D>
__global__ void convert_rows_kernel(const S* __restrict__ src, int src_dim, D* __restrict__ dst, int dst_pitch,
                                    int64_t n) {
  // one thread per destination element; padded columns are written as zero
  int64_t total = n * dst_pitch;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / dst_pitch;
    int c = (int)(e - r * dst_pitch);
    float v = (c < src_dim) ? to_f32<S>(src[r * src_dim + c]) : 0.0f;
    dst[e] = from_f32<D>(v);
  }
}

// prepare = query staging + list reset. dst dtype float: one plane (EXACT mode); 16-bit: `terms` planes, plane t =
// round(q - plane_0 - ... - plane_{t-1}). The remainders are exact in float32 (each has fewer significant bits), so
// 3 bf16 / 2-3 fp16 terms reproduce the float32 query exactly (barring fp16 range limits).
template <typename S, typename D>
__global__ void prepare_kernel(const S* __restrict__ src, int src_dim, D* __restrict__ dst, int dst_pitch, int64_t n,
                               int64_t plane_rows, int64_t fill_rows, int terms, int* __restrict__ cnt,
                               float* __restrict__ tau, int first_rows, int* __restrict__ term_any) {
  pdl_launch_dependents();
  pdl_wait();  // the previous search on this stream may still be reading the staging buffer / lists
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid < n) {
    cnt[gtid * kCntStride] = first_rows;
    tau[gtid] = -INFINITY;
  }
  // rows [n, fill_rows) are the rest of the last query tile: zero them, so that whatever an earlier call left there
  // (possibly inf / NaN bit patterns of another dtype) never reaches the MMAs — a garbage +inf score would pass the
  // `score >= tau` filter of an unused column even against tau = +inf
  int64_t total = fill_rows * dst_pitch;
  int nonzero = 0;  // bit t: this thread wrote a nonzero element of term t
  for (int64_t e = gtid; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / dst_pitch;
    int c = (int)(e - r * dst_pitch);
    float v = (r < n && c < src_dim) ? to_f32<S>(src[r * src_dim + c]) : 0.0f;
    for (int t = 0; t < terms; ++t) {
      D d = from_f32<D>(v);
      dst[(size_t)t * plane_rows * dst_pitch + e] = d;
      const float dv = to_f32<D>(d);
      nonzero |= (dv != 0.0f) ? (1 << t) : 0;
      v = __fsub_rn(v, dv);
    }
  }
  // which terms carry anything at all: float32 queries that are exactly representable in the store dtype (encoder
  // outputs computed in bf16 / fp16 and widened) have empty correction terms, and the scoring kernel skips them
  const int any1 = __syncthreads_or(nonzero & 2), any2 = __syncthreads_or(nonzero & 4);  // logical ORs over the block
  if (threadIdx.x == 0) term_any[blockIdx.x] = 1 | (any1 ? 2 : 0) | (any2 ? 4 : 0);
}

// fp32 store -> 3 bf16 planes with row = p0 + p1 + p2 exactly: p_i = bf16_rn(what p_0..p_{i-1} left), the remainders
// are exact in float32 and the last one (<= 8 significant bits) is itself a bf16
__global__ void split_planes_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t row0, int64_t n,
                                    int pitch, int64_t plane_elems) {
  const int64_t total = n * pitch;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t at = row0 * pitch + e;
    float v = src[at];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const __nv_bfloat16 d = from_f32<__nv_bfloat16>(v);
      dst[(size_t)t * plane_elems + at] = d;
      v = __fsub_rn(v, to_f32<__nv_bfloat16>(d));
    }
  }
}

template <typename D>
__device__ __forceinline__ D round_store(float v);
template <>
__device__ __forceinline__ float round_store<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 round_store<__nv_bfloat16>(float v) {
  return __ushort_as_bfloat16(vodb_f32_to_bf16(v));
}
template <>
__device__ __forceinline__ __half round_store<__half>(float v) { return __ushort_as_half(vodb_f32_to_f16(v)); }

__global__ void synth_norm_kernel(uint64_t seed, int64_t global_row0, int64_t n, int dim, float* __restrict__ inv) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float ss = 0.0f;
  for (int c = 0; c < dim; ++c) {
    float v = vodb_synth_value(seed, (uint64_t)(global_row0 + r), (uint32_t)c);
    ss = VM_ADD(ss, VM_MUL(v, v));
  }
  inv[r] = __fsqrt_rn(ss);
}

template <typename D>
__global__ void synth_fill_kernel(D* __restrict__ dst, int dim, int pitch, uint64_t seed, int64_t global_row0,
                                  int64_t n, const float* __restrict__ norms) {
  // one thread per group of 4 columns (one Philox call)
  const int groups = pitch / 4;
  int64_t total = n * groups;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / groups;
    int cg = (int)(e - r * groups);
    uint64_t grow = (uint64_t)(global_row0 + r);
    vm_u32x4 w = vodb_philox4x32((uint32_t)cg, (uint32_t)grow, (uint32_t)(grow >> 32), 0x53594e54u, (uint32_t)seed,
                                 (uint32_t)(seed >> 32));
    float nrm = norms ? norms[r] : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = cg * 4 + j;
      float v = 0.0f;
      if (c < dim) {
        v = vodb_synth_from_word(w.v[j]);
        if (norms && nrm > 0.0f) v = VM_DIV(v, nrm);
      }
      dst[r * pitch + c] = round_store<D>(v);
    }
  }
}

template <typename S>
__global__ void read_rows_kernel(const S* __restrict__ src, int dim, int pitch, int64_t n, float* __restrict__ out) {
  int64_t total = n * dim;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / dim;
    int c = (int)(e - r * dim);
    out[e] = to_f32<S>(src[r * pitch + c]);
  }
}

inline int grid_for(int64_t total, int threads = 256) {
  int64_t g = (total + threads - 1) / threads;
  return (int)std::min<int64_t>(std::max<int64_t>(g, 1), 148 * 32);
}

template <typename S>
int convert_dispatch_dst(const void* src, int src_dim, void* dst, int dst_dtype, int dst_pitch, int64_t n,
                         cudaStream_t st) {
  int64_t total = n * dst_pitch;
  if (total == 0) return VODB_OK;
  int g = grid_for(total);
  const S* s = reinterpret_cast<const S*>(src);
  switch (dst_dtype) {
    case VODB_F32: convert_rows_kernel<S, float><<<g, 256, 0, st>>>(s, src_dim, (float*)dst, dst_pitch, n); break;
    case VODB_BF16: convert_rows_kernel<S, __nv_bfloat16><<<g, 256, 0, st>>>(s, src_dim, (__nv_bfloat16*)dst, dst_pitch, n); break;
    case VODB_F16: convert_rows_kernel<S, __half><<<g, 256, 0, st>>>(s, src_dim, (__half*)dst, dst_pitch, n); break;
    default: set_error("bad dst dtype %d", dst_dtype); return VODB_EINVAL;
  }
  VODB_CUDA_CHECK(cudaGetLastError());
  return VODB_OK;
}

}  // namespace

int launch_convert_rows(const void* src, int src_dtype, int src_dim, void* dst, int dst_dtype, int dst_pitch,
                        int64_t n, cudaStream_t st) {
  switch (src_dtype) {
    case VODB_F32: return convert_dispatch_dst<float>(src, src_dim, dst, dst_dtype, dst_pitch, n, st);
    case VODB_BF16: return convert_dispatch_dst<__nv_bfloat16>(src, src_dim, dst, dst_dtype, dst_pitch, n, st);
    case VODB_F16: return convert_dispatch_dst<__half>(src, src_dim, dst, dst_dtype, dst_pitch, n, st);
  }
  set_error("bad src dtype %d", src_dtype);
  return VODB_EINVAL;
}

namespace {
template <typename S>
int prepare_dispatch_dst(const void* src, int src_dim, void* dst, int dst_dtype, int dst_pitch, int64_t n,
                         int64_t plane_rows, int terms, int* cnt, float* tau, int first_rows, int* term_any,
                         int* term_blocks, cudaStream_t st) {
  // query tiles are 64, 128 or 256 rows (score_tc.cu launch_score_tensor): clear the rest of the last one
  const int64_t fill_rows = std::min<int64_t>(plane_rows, n <= 64 ? 64 : n <= 128 ? 128 : (n + 255) / 256 * 256);
  int64_t total = fill_rows * dst_pitch;
  if (n == 0) return VODB_OK;
  int g = std::max(grid_for(total), (int)((n + 255) / 256));  // every query needs a thread for the list reset
  if (g > kTermSlots) {
    set_error("launch_prepare: %lld queries need %d blocks (> %d)", (long long)n, g, kTermSlots);
    return VODB_EINVAL;
  }
  *term_blocks = g;
  const S* s = reinterpret_cast<const S*>(src);
  cudaError_t e;
  switch (dst_dtype) {
    case VODB_F32: e = launch_pdl(prepare_kernel<S, float>, dim3(g), dim3(256), 0, st, s, src_dim, (float*)dst, dst_pitch, n, plane_rows, fill_rows, 1, cnt, tau, first_rows, term_any); break;
    case VODB_BF16: e = launch_pdl(prepare_kernel<S, __nv_bfloat16>, dim3(g), dim3(256), 0, st, s, src_dim, (__nv_bfloat16*)dst, dst_pitch, n, plane_rows, fill_rows, terms, cnt, tau, first_rows, term_any); break;
    case VODB_F16: e = launch_pdl(prepare_kernel<S, __half>, dim3(g), dim3(256), 0, st, s, src_dim, (__half*)dst, dst_pitch, n, plane_rows, fill_rows, terms, cnt, tau, first_rows, term_any); break;
    default: set_error("launch_prepare: bad destination dtype %d", dst_dtype); return VODB_EINVAL;
  }
  VODB_CUDA_CHECK(e);
  return VODB_OK;
}
}  // namespace

int launch_prepare(const void* src, int src_dtype, int src_dim, void* dst, int dst_dtype, int dst_pitch, int64_t n,
                   int64_t plane_rows, int terms, int* cnt, float* tau, int first_rows, int* term_any, int* term_blocks,
                   cudaStream_t st) {
  switch (src_dtype) {
    case VODB_F32: return prepare_dispatch_dst<float>(src, src_dim, dst, dst_dtype, dst_pitch, n, plane_rows, terms, cnt, tau, first_rows, term_any, term_blocks, st);
    case VODB_BF16: return prepare_dispatch_dst<__nv_bfloat16>(src, src_dim, dst, dst_dtype, dst_pitch, n, plane_rows, terms, cnt, tau, first_rows, term_any, term_blocks, st);
    case VODB_F16: return prepare_dispatch_dst<__half>(src, src_dim, dst, dst_dtype, dst_pitch, n, plane_rows, terms, cnt, tau, first_rows, term_any, term_blocks, st);
  }
  set_error("bad src dtype %d", src_dtype);
  return VODB_EINVAL;
}

int launch_split_planes(const float* src, void* planes, int64_t row0, int64_t n, int pitch, int64_t n_rows,
                        cudaStream_t st) {
  if (n <= 0) return VODB_OK;
  split_planes_kernel<<<grid_for(n * pitch), 256, 0, st>>>(src, reinterpret_cast<__nv_bfloat16*>(planes), row0, n, pitch,
                                                            n_rows * pitch);
  VODB_CUDA_CHECK(cudaGetLastError());
  return VODB_OK;
}

int launch_fill_synthetic(void* dst, int dtype, int dim, int pitch, uint64_t seed, int64_t global_row0, int64_t n,
                          int unit_norm, cudaStream_t st) {
  if (n == 0) return VODB_OK;
  float* norms = nullptr;
  if (unit_norm) {
    VODB_CUDA_CHECK(cudaMallocAsync(&norms, (size_t)n * sizeof(float), st));
    synth_norm_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seed, global_row0, n, dim, norms);
  }
  int g = grid_for(n * (pitch / 4));
  switch (dtype) {
    case VODB_F32: synth_fill_kernel<float><<<g, 256, 0, st>>>((float*)dst, dim, pitch, seed, global_row0, n, norms); break;
    case VODB_BF16: synth_fill_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((__nv_bfloat16*)dst, dim, pitch, seed, global_row0, n, norms); break;
    case VODB_F16: synth_fill_kernel<__half><<<g, 256, 0, st>>>((__half*)dst, dim, pitch, seed, global_row0, n, norms); break;
    default: set_error("bad dtype %d", dtype); return VODB_EINVAL;
  }
  VODB_CUDA_CHECK(cudaGetLastError());
  if (norms) VODB_CUDA_CHECK(cudaFreeAsync(norms, st));
  return VODB_OK;
}

int launch_read_rows(const void* src, int dtype, int dim, int pitch, int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return VODB_OK;
  int g = grid_for(n * dim);
  switch (dtype) {
    case VODB_F32: read_rows_kernel<float><<<g, 256, 0, st>>>((const float*)src, dim, pitch, n, out); break;
    case VODB_BF16: read_rows_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, dim, pitch, n, out); break;
    case VODB_F16: read_rows_kernel<__half><<<g, 256, 0, st>>>((const __half*)src, dim, pitch, n, out); break;
    default: set_error("bad dtype %d", dtype); return VODB_EINVAL;
  }
  VODB_CUDA_CHECK(cudaGetLastError());
  return VODB_OK;
}

namespace {

inline int mode_terms(int mode) { return mode == VODB_MODE_TENSOR_X3 ? 3 : mode == VODB_MODE_TENSOR_X2 ? 2 : 1; }
inline bool is_tensor_mode(int mode) { return mode >= VODB_MODE_TENSOR && mode <= VODB_MODE_TENSOR_X3; }

int pow2ceil_host(int64_t x) {
  int64_t p = 1;
  while (p < x) p <<= 1;
  return (int)p;
}

// candidate-list capacity: room for ~128 k entries, within a ~4 GB budget for the whole batch
int choose_cap(int nq, int k) {
  static const char* env_cap = std::getenv("VODB_CAP");  // development knob (schedule sweeps)
  if (env_cap) return std::max(pow2ceil_host(4LL * k), pow2ceil_host(std::atoll(env_cap)));
  int64_t cap = pow2ceil_host(std::max<int64_t>(128LL * k, 8192));
  cap = std::min<int64_t>(cap, 65536);
  // small batches: 64k slots per list (<= 134 MB in all) let the segments grow 80-fold, i.e. a 1.25M-row shard is two
  // segments and a 10M-row shard three (one select + two launches less per search; scripts/r02_sweep_schedule2.py)
  if (nq <= 256) cap = 65536;
  const int64_t budget = 4LL << 30;
  while (cap > 4LL * k && cap > 2048 && (int64_t)nq * cap * 8 > budget) cap >>= 1;
  cap = std::max<int64_t>(cap, pow2ceil_host(4LL * k));
  return (int)cap;
}

// The candidate lists, the query staging and the overflow flag are per store, so the kernels of two calls must not
// overlap. Calls on one stream are ordered by the stream; a call that arrives on a DIFFERENT stream than its
// predecessor (a torch side stream after the default stream, the server thread next to the training thread) first
// waits for the device to drain. Rare by construction, and the only host-blocking point of the asynchronous path.
int order_streams(vodb_store* s, cudaStream_t st) {
  if (s->last_stream_set && s->last_stream != st) VODB_CUDA_CHECK(cudaDeviceSynchronize());
  s->last_stream = st;
  s->last_stream_set = true;
  return VODB_OK;
}

int ensure_workspace(vodb_store* s, int nq, int k, int q_elem_bytes, cudaStream_t st) {
  Workspace& w = s->ws;
  int rc_order = order_streams(s, st);
  if (rc_order != VODB_OK) return rc_order;
  int cap = choose_cap(nq, k);
  // lists are laid out [nq, cap] with the cap of THIS call as row stride (a larger stride left over from an earlier
  // call with another k would only spread the lists over more cache lines)
  const size_t need_elems = (size_t)nq * cap;
  if (need_elems > w.list_elems) {
    if (w.cand_s) cudaFree(w.cand_s);
    if (w.cand_i) cudaFree(w.cand_i);
    w.cand_s = nullptr; w.cand_i = nullptr;
    VODB_CUDA_CHECK(cudaMalloc(&w.cand_s, need_elems * sizeof(float)));
    VODB_CUDA_CHECK(cudaMalloc(&w.cand_i, need_elems * sizeof(int32_t)));
    w.list_elems = need_elems;
  }
  if (nq > w.nq_cap) {
    if (w.cnt) cudaFree(w.cnt);
    if (w.tau) cudaFree(w.tau);
    w.cnt = nullptr; w.tau = nullptr;
    VODB_CUDA_CHECK(cudaMalloc(&w.cnt, (size_t)nq * kCntStride * sizeof(int)));
    VODB_CUDA_CHECK(cudaMalloc(&w.tau, (size_t)nq * sizeof(float)));
    w.nq_cap = nq;
  }
  w.cap = cap;
  if (!w.overflow) {
    VODB_CUDA_CHECK(cudaMalloc(&w.overflow, 2 * sizeof(int)));
    VODB_CUDA_CHECK(cudaMalloc(&w.term_any, kTermSlots * sizeof(int)));
    // on the call's stream: a legacy-stream memset is not ordered against work on a non-blocking stream
    VODB_CUDA_CHECK(cudaMemsetAsync(w.overflow, 0, 2 * sizeof(int), st));
    VODB_CUDA_CHECK(cudaMallocHost(&w.overflow_host, 2 * sizeof(int)));
  }
  // staged queries: rows padded to a multiple of 256 so that any TMA box is in bounds; zero filled
  size_t rows_pad = ((size_t)nq + 255) / 256 * 256;
  size_t need = rows_pad * (size_t)s->pitch * 6;  // fp32 staging (4 B) or up to three 16-bit term planes (6 B)
  if (need > w.q_stage_bytes) {
    if (w.q_stage) cudaFree(w.q_stage);
    w.q_stage = nullptr;
    VODB_CUDA_CHECK(cudaMalloc(&w.q_stage, need));
    VODB_CUDA_CHECK(cudaMemsetAsync(w.q_stage, 0, need, st));
    w.q_stage_bytes = need;
  }
  size_t need_in = (size_t)nq * s->dim * q_elem_bytes;
  if (need_in > w.q_in_bytes) {
    if (w.q_in) cudaFree(w.q_in);
    w.q_in = nullptr;
    VODB_CUDA_CHECK(cudaMalloc(&w.q_in, need_in));
    w.q_in_bytes = need_in;
  }
  size_t out_need = (size_t)nq * k;
  if (out_need > w.out_cap) {
    if (w.out_pack) cudaFree(w.out_pack);
    if (w.out_host) cudaFreeHost(w.out_host);
    w.out_pack = nullptr; w.out_host = nullptr;
    VODB_CUDA_CHECK(cudaMalloc(&w.out_pack, out_need * 12 + 16));
    VODB_CUDA_CHECK(cudaMallocHost(&w.out_host, out_need * 12 + 16));
    w.out_cap = out_need;
  }
  return VODB_OK;
}

void free_workspace(Workspace& w) {
  cudaFree(w.cand_s); cudaFree(w.cand_i); cudaFree(w.cnt); cudaFree(w.tau); cudaFree(w.overflow); cudaFree(w.term_any);
  if (w.overflow_host) cudaFreeHost(w.overflow_host);
  cudaFree(w.q_stage); cudaFree(w.q_in); cudaFree(w.out_pack); cudaFree(w.chain_dev);
  if (w.out_host) cudaFreeHost(w.out_host);
  if (w.chain_host) cudaFreeHost(w.chain_host);
  w = Workspace();
}

// row segments [b_i, b_{i+1}) of the scan; boundaries are multiples of 128 (except the end)
std::vector<int64_t> plan_segments(int64_t n, int cap, int k, int nq, bool safe) {
  std::vector<int64_t> b;
  b.push_back(0);
  auto round128 = [](int64_t x) { return std::max<int64_t>(128, x / 128 * 128); };
  if (safe) {
    int64_t step = round128(cap - k);  // cap >= 8192 > k + 128: a segment of cap-k rows cannot overflow a list
    for (int64_t r = step; r < n; r += step) b.push_back(r);
    b.push_back(n);
    return b;
  }
  // Large batches refresh the thresholds more often (smaller first segment, growth 3 instead of up to 96): the
  // number of survivors per query over the whole scan is ~ k*g*log_{1+g}(n/first), every survivor costs epilogue
  // time, and with thousands of queries that outweighs the fixed cost of a few more (select + launch) pairs.
  // (scripts/sweep_schedule_large.py, 8192 queries: first 4096 / 16384 rows x growth 2 / 3 / 5 are within +-3% of
  // each other for k = 100 and k = 1000 — profiles/r01m_schedule_sweep_large.jsonl.)
  const bool large_batch = nq > 256;
  // first segment ("dump": every score stored, then one select). Measured on B200, 64 queries, k=100
  // (profiles/r02c_schedule_sweep.jsonl, after the per-query counters were spread over cache lines): 4096 rows x
  // growth 20-32, 8192 x 64, 16384 x 80-160 and 32768 x 40 are within 2% of each other on a 1.25M-row and on a
  // 10M-row shard; 16384 rows with the widest growth has the fewest launches (2 resp. 3 segments). Still the best
  // after the selects moved to the thread-maximum bound (profiles/r02l_schedule_sweep.jsonl).
  int64_t first = round128(std::min<int64_t>(cap / 2, std::max<int64_t>(large_batch ? 4096 : 16384, 16LL * k)));
  if (large_batch) first = std::min<int64_t>(first, std::max<int64_t>(4096, round128(4LL * k)));
  // tuning knobs (development): VODB_FIRST_ROWS / VODB_GROWTH override the schedule of small batches
  static const char* env_first = std::getenv("VODB_FIRST_ROWS");
  static const char* env_growth = std::getenv("VODB_GROWTH");
  if (env_first && !large_batch) first = round128(std::min<int64_t>(cap / 2, std::max<int64_t>(2LL * k, std::atoll(env_first))));
  static const char* env_first_large = std::getenv("VODB_FIRST_ROWS_LARGE");
  static const char* env_growth_large = std::getenv("VODB_GROWTH_LARGE");
  if (env_first_large && large_batch) first = round128(std::min<int64_t>(cap / 2, std::max<int64_t>(2LL * k, std::atoll(env_first_large))));
  if (first >= n) {
    b.push_back(n);
    return b;
  }
  b.push_back(first);
  // growth: expected survivors of a segment = k * seg/before; keep that below cap/8
  double g = std::max(1.0, std::min((double)cap / (8.0 * k), large_batch ? 3.0 : 96.0));
  if (env_growth && !large_batch) g = std::max(1.0, std::atof(env_growth));
  if (env_growth_large && large_batch) g = std::max(1.0, std::min((double)cap / (8.0 * k), std::atof(env_growth_large)));
  int64_t cur = first;
  while (cur < n) {
    int64_t seg = round128((int64_t)(g * (double)cur));
    int64_t nxt = cur + seg;
    if (nxt >= n || (n - nxt) < seg / 4) nxt = n;  // fold a short tail into the last segment
    b.push_back(nxt);
    cur = nxt;
  }
  return b;
}

// fp32 stores reach the tensor cores through three bf16 planes (6 more bytes per element, allocated on the first
// tensor-mode search and extended after later adds). The 16-bit stores need nothing.
int ensure_planes(vodb_store* s, cudaStream_t st) {
  if (s->dtype != VODB_F32) return VODB_OK;
  if (!s->planes) {
    cudaError_t e = cudaMalloc(&s->planes, (size_t)3 * s->n_rows * s->pitch * sizeof(__nv_bfloat16));
    if (e != cudaSuccess) {
      cudaGetLastError();
      s->planes = nullptr;
      set_error("tensor modes on a float32 store need %.1f GB for the bf16 planes (%s); use VODB_MODE_EXACT or a bf16 store",
                3.0 * s->n_rows * s->pitch * 2 / 1e9, cudaGetErrorString(e));
      return VODB_ENOMEM;
    }
    s->planes_rows_done = 0;
    s->tmap_planes_valid = false;
  }
  if (s->planes_rows_done < s->n_added) {
    int rc = launch_split_planes(reinterpret_cast<const float*>(s->data), s->planes, s->planes_rows_done,
                                 s->n_added - s->planes_rows_done, s->pitch, s->n_rows, st);
    if (rc != VODB_OK) return rc;
    s->planes_rows_done = s->n_added;
  }
  return VODB_OK;
}

int run_scan(vodb_store* s, const void* q_dev, int q_dtype, int nq, int k, int mode, bool safe, float* out_s,
             int64_t* out_i, cudaStream_t st, const ExchangeDst* xd = nullptr) {
  Workspace& w = s->ws;
  const void* q_stage = w.q_stage;
  std::vector<int64_t> b = plan_segments(s->n_added, w.cap, k, nq, safe);
  // one launch stages the queries (fp32 plane, or `terms` 16-bit planes of round_up(nq,256) rows each) and resets the
  // lists; the first segment (dump mode) stores every score, so cnt starts at its row count
  const int64_t rows_pad = ((int64_t)nq + 255) / 256 * 256;
  const bool tensor = is_tensor_mode(mode);
  int term_blocks = 0;
  const bool planes = tensor && s->dtype == VODB_F32;   // fp32 store: bf16 planes x bf16 query terms
  const int tc_dtype = planes ? VODB_BF16 : s->dtype;   // what the tensor-core kernel multiplies
  int rc = planes ? ensure_planes(s, st) : VODB_OK;
  if (rc != VODB_OK) return rc;
  rc = launch_prepare(q_dev, q_dtype, s->dim, w.q_stage, tensor ? tc_dtype : VODB_F32, s->pitch, nq, rows_pad,
                      tensor ? mode_terms(mode) : 1, w.cnt, w.tau, (int)(b[1] - b[0]), w.term_any, &term_blocks, st);
  if (rc != VODB_OK) return rc;
  int64_t launches = 1;
  for (size_t i = 0; i + 1 < b.size(); ++i) {
    SegmentArgs a;
    a.corpus = planes ? s->planes : s->data;
    a.dtype = tensor ? tc_dtype : s->dtype;
    a.pitch = s->pitch;
    a.row_begin = b[i];
    a.row_end = b[i + 1];
    a.queries = q_stage;
    a.nq = nq;
    a.cand_s = w.cand_s;
    a.cand_i = w.cand_i;
    a.cnt = w.cnt;
    a.tau = w.tau;
    a.overflow = w.overflow;
    a.cap = w.cap;
    a.dump = (i == 0);
    a.terms = tensor ? mode_terms(mode) : 1;
    a.planes = planes ? a.terms : 1;
    a.term_any = w.term_any;
    a.term_blocks = term_blocks;
    ProfileState* prof = s->profiling ? static_cast<ProfileState*>(s->prof) : nullptr;
    if (prof) cudaEventRecord(prof->next(), st);
    rc = is_tensor_mode(mode) ? launch_score_tensor(s, a, st) : launch_score_exact(a, s->sm_count, st);
    if (rc != VODB_OK) return rc;
    if (prof) {
      cudaEventRecord(prof->next(), st);
      prof->kinds.push_back(0);
      prof->bytes += (double)(b[i + 1] - b[i]) * s->dim * dtype_size(s->dtype);
      cudaEventRecord(prof->next(), st);
    }
    bool last = (i + 2 == b.size());
    // list length the select will see: the whole first segment (dump), afterwards k + ~k*g survivors (x2 slack)
    const int expected_n = (i == 0) ? (int)(b[1] - b[0])
                                    : (int)std::min<int64_t>(w.cap, 2 * (int64_t)k * (1 + (b[i + 1] - b[i]) / std::max<int64_t>(b[i], 1)) + 512);
    rc = launch_select(w.cand_s, w.cand_i, w.cnt, w.tau, w.cap, nq, k, last, out_s, out_i, s->row_offset, st,
                       last ? xd : nullptr, expected_n);
    if (rc != VODB_OK) return rc;
    if (prof) {
      cudaEventRecord(prof->next(), st);
      prof->kinds.push_back(1);
    }
    launches += 2;
  }
  s->stats[0] = launches;
  s->stats[1] = (int64_t)b.size() - 1;
  s->stats[2] = w.cap;
  s->stats[3] = safe ? 1 : 0;
  return VODB_OK;
}

}  // namespace

}  // namespace vodb

using namespace vodb;

namespace {
// Device scratch of the host-buffer entry points that have no store to hang a workspace on (vodb_sample,
// vodb_merge_topk, vodb_merge_results): one grow-only buffer per device, held for the duration of a call, so that
// a per-batch call costs copies + kernel instead of a cudaMalloc / cudaFree pair (the free alone synchronises the
// device). Released at process exit with the context.
struct CallScratch {
  std::mutex mu;
  char* dev = nullptr;
  size_t bytes = 0;
};
CallScratch& call_scratch(int device) {
  static CallScratch table[64];
  return table[device >= 0 && device < 64 ? device : 0];
}
// returns nullptr (error set) when the allocation fails; the caller holds `cs.mu`
char* scratch_reserve(CallScratch& cs, size_t bytes, const char* who) {
  if (bytes > cs.bytes) {
    if (cs.dev) cudaFree(cs.dev);
    cs.dev = nullptr;
    cs.bytes = 0;
    const size_t want = bytes + bytes / 4;  // head-room: batch shapes repeat, widths wobble
    cudaError_t e = cudaMalloc(&cs.dev, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("%s: cudaMalloc(%zu bytes) failed: %s", who, want, cudaGetErrorString(e));
      return nullptr;
    }
    cs.bytes = want;
  }
  return cs.dev;
}
}  // namespace

// Peer-mapped exchange buffer of one rank (vodb_xchg_*): [2 parities][world source ranks][slot entries][3 tagged
// 8-byte words] (see ExchangeDst), zero-initialised, exported to the peers through CUDA IPC.
struct vodb_xchg {
  int device = 0, rank = 0, world = 1;
  int max_nq = 0, max_k = 0;
  size_t slot = 0;            // entries per (parity, source rank) slot = max_nq * max_k
  size_t slot_words = 0;      // 8-byte words per slot: 3 per entry + the overflow-flag word (padded to 8 words)
  size_t bytes = 0;
  char* local = nullptr;      // this rank's buffer
  char* peer[kMaxPeers] = {}; // peer-mapped base pointers (peer[rank] == local)
  bool connected = false;
  uint32_t epoch = 0;
};

namespace {
// host queries are copied to the device here; staging (convert / split / pad) happens in run_scan's prepare kernel.
// Rows of the staging buffer past nq keep whatever an earlier call left there: their score columns are never read
// (tau = +inf / q < nq guards in the scoring kernels).
int upload_queries(vodb_store* s, const void* queries, int q_dtype, int q_on_device, int nq, cudaStream_t st,
                   const void** q_dev) {
  Workspace& w = s->ws;
  *q_dev = queries;
  if (!q_on_device) {
    VODB_CUDA_CHECK(cudaMemcpyAsync(w.q_in, queries, (size_t)nq * s->dim * dtype_size(q_dtype), cudaMemcpyHostToDevice, st));
    *q_dev = w.q_in;
  }
  return VODB_OK;
}

// Host-resident float32 queries that are exactly representable in a 16-bit store dtype (what a bf16 / fp16 encoder
// hands over after widening, reference src/vod_ops/workflows/predict/compute.py:128-129) have empty correction terms:
// the multi-term modes then return the same bits as VODB_MODE_TENSOR, which runs the faster one-term kernels. The
// scan over the host buffer costs a few microseconds (49k floats at 64 x 768) and stops at the first counterexample.
// Device-resident queries are classified on the device instead (prepare_kernel's term masks).
int effective_mode(const vodb_store* s, const void* queries, int q_dtype, int q_on_device, int nq, int mode) {
  if (q_on_device || q_dtype != VODB_F32 || (mode != VODB_MODE_TENSOR_X2 && mode != VODB_MODE_TENSOR_X3)) return mode;
  if (s->dtype != VODB_BF16 && s->dtype != VODB_F16) return mode;
  const uint32_t* bits = reinterpret_cast<const uint32_t*>(queries);
  const size_t n = (size_t)nq * s->dim;
  if (s->dtype == VODB_BF16) {
    uint32_t low = 0;
    for (size_t i = 0; i < n; i += 4096) {  // blocks: vectorisable OR, early exit between blocks
      const size_t e = std::min(n, i + 4096);
      for (size_t j = i; j < e; ++j) low |= bits[j] & 0xffffu;
      if (low) return mode;
    }
    return VODB_MODE_TENSOR;
  }
  const float* q = reinterpret_cast<const float*>(queries);
  for (size_t i = 0; i < n; ++i) {
    const float back = vodb_f16_to_f32(vodb_f32_to_f16(q[i]));
    if (vm_f2u(back) != bits[i]) return mode;
  }
  return VODB_MODE_TENSOR;
}

int check_search_args(vodb_store* s, const void* queries, int q_dtype, int nq, int k, int mode, const float* out_scores,
                      const int64_t* out_idx, const char* fn) {
  VODB_REQUIRE(s != nullptr, "%s: store is NULL", fn);
  VODB_REQUIRE(queries != nullptr || nq == 0, "%s: queries is NULL", fn);
  VODB_REQUIRE(nq >= 0, "%s: nq=%d < 0", fn, nq);
  VODB_REQUIRE(k >= 1 && k <= VODB_MAX_K, "%s: k=%d outside [1, %d]", fn, k, VODB_MAX_K);
  VODB_REQUIRE(q_dtype == VODB_F32 || q_dtype == VODB_BF16 || q_dtype == VODB_F16, "%s: bad query dtype %d", fn, q_dtype);
  VODB_REQUIRE(mode == VODB_MODE_EXACT || is_tensor_mode(mode), "%s: bad mode %d", fn, mode);
  VODB_REQUIRE(nq == 0 || (out_scores != nullptr && out_idx != nullptr), "%s: output pointer is NULL", fn);
  return VODB_OK;
}
}  // namespace

extern "C" {

const char* vodb_last_error(void) { return get_error(); }
int vodb_abi_version(void) { return VODB_ABI_VERSION; }

int vodb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return VODB_ECUDA;
  }
  return n;
}

int vodb_store_create(vodb_store** out, int device, int64_t n_rows, int dim, int dtype, int64_t row_offset) {
  VODB_REQUIRE(out != nullptr, "vodb_store_create: out is NULL");
  *out = nullptr;
  VODB_REQUIRE(n_rows >= 0 && n_rows < (1LL << 31), "vodb_store_create: n_rows=%lld out of range [0, 2^31)", (long long)n_rows);
  VODB_REQUIRE(dim > 0 && dim <= 16384, "vodb_store_create: dim=%d out of range (0, 16384]", dim);
  VODB_REQUIRE(dtype == VODB_F32 || dtype == VODB_BF16 || dtype == VODB_F16, "vodb_store_create: bad dtype %d", dtype);
  VODB_REQUIRE(row_offset >= 0, "vodb_store_create: negative row_offset");
  int ndev = 0;
  VODB_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  VODB_REQUIRE(device >= 0 && device < ndev, "vodb_store_create: device %d not in [0,%d)", device, ndev);
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("cudaSetDevice(%d) failed", device);
    return VODB_ECUDA;
  }
  vodb_store* s = new (std::nothrow) vodb_store();
  if (!s) {
    set_error("out of host memory");
    return VODB_ENOMEM;
  }
  s->device = device;
  s->n_rows = n_rows;
  s->dim = dim;
  s->pitch = (dim + kPitchAlign - 1) / kPitchAlign * kPitchAlign;
  // Development knob: VODB_PITCH_PAD=2 adds one unscanned 64-element chunk per row. Tested as a remedy for
  // power-of-two row strides (dim 1024 bf16 = 2 KB): a controlled sweep over row lengths and store sizes showed no
  // stride penalty and a 0-8% loss from the padding (profiles/r01m_pitch_sweep.jsonl), so the layout stays dense.
  static const char* env_pad = std::getenv("VODB_PITCH_PAD");
  if (env_pad && env_pad[0] == '2') s->pitch += kPitchAlign;
  s->dtype = dtype;
  s->row_offset = row_offset;
  cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
  size_t bytes = (size_t)std::max<int64_t>(n_rows, 1) * s->pitch * dtype_size(dtype);
  cudaError_t e = cudaMalloc(&s->data, bytes);
  if (e != cudaSuccess) {
    set_error("vodb_store_create: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    delete s;
    cudaGetLastError();
    return VODB_ENOMEM;
  }
  *out = s;
  return VODB_OK;
}

void vodb_store_destroy(vodb_store* s) {
  if (!s) return;
  DeviceGuard guard(s->device);
  cudaDeviceSynchronize();
  free_workspace(s->ws);
  if (s->prof) {
    ProfileState* p = static_cast<ProfileState*>(s->prof);
    for (cudaEvent_t e : p->pool) cudaEventDestroy(e);
    delete p;
  }
  for (int b = 0; b < 2; ++b) {
    if (s->stage[b]) cudaFree(s->stage[b]);
    if (s->copied[b]) cudaEventDestroy(s->copied[b]);
    if (s->converted[b]) cudaEventDestroy(s->converted[b]);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  if (s->planes) cudaFree(s->planes);
  if (s->data) cudaFree(s->data);
  delete s;
}

int vodb_store_add(vodb_store* s, const void* rows, int src_dtype, int src_on_device, int64_t row0, int64_t n,
                   void* stream) {
  VODB_REQUIRE(s != nullptr, "vodb_store_add: store is NULL");
  std::lock_guard<std::mutex> store_lock(s->mu);
  VODB_REQUIRE(n >= 0 && row0 >= 0 && row0 + n <= s->n_rows, "vodb_store_add: rows [%lld, %lld) outside the store (%lld rows)",
               (long long)row0, (long long)(row0 + n), (long long)s->n_rows);
  VODB_REQUIRE(src_dtype == VODB_F32 || src_dtype == VODB_BF16 || src_dtype == VODB_F16, "vodb_store_add: bad src dtype %d", src_dtype);
  if (n == 0) return VODB_OK;
  VODB_REQUIRE(rows != nullptr, "vodb_store_add: rows is NULL");
  // append or overwrite only (faiss `index.add` appends): rows in a gap would be searchable uninitialised memory
  VODB_REQUIRE(row0 <= s->n_added, "vodb_store_add: row0=%lld leaves a gap after the %lld rows added so far (add blocks in order)",
               (long long)row0, (long long)s->n_added);
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t esz = dtype_size(src_dtype);
  char* dst = reinterpret_cast<char*>(s->data) + (size_t)row0 * s->pitch * dtype_size(s->dtype);
  if (src_on_device) {
    int rc = launch_convert_rows(rows, src_dtype, s->dim, dst, s->dtype, s->pitch, n, st);
    if (rc != VODB_OK) return rc;
  } else {
    // chunked upload through two device staging buffers (64 MiB each), converted on the device: the copy of chunk
    // i+1 (copy stream) overlaps the conversion of chunk i (caller's stream); one host wait at the end
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / ((size_t)s->dim * esz));
    size_t need = std::min<size_t>((size_t)n, chunk_rows) * s->dim * esz;
    if (need > s->stage_bytes) {
      for (int b = 0; b < 2; ++b) {
        if (s->stage[b]) cudaFree(s->stage[b]);
        s->stage[b] = nullptr;
      }
      s->stage_bytes = 0;
      VODB_CUDA_CHECK(cudaMalloc(&s->stage[0], need));
      VODB_CUDA_CHECK(cudaMalloc(&s->stage[1], need));
      s->stage_bytes = need;
    }
    if (!s->copy_stream) {
      VODB_CUDA_CHECK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        VODB_CUDA_CHECK(cudaEventCreateWithFlags(&s->copied[b], cudaEventDisableTiming));
        VODB_CUDA_CHECK(cudaEventCreateWithFlags(&s->converted[b], cudaEventDisableTiming));
      }
    }
    int64_t chunk = 0;
    for (int64_t r = 0; r < n; r += (int64_t)chunk_rows, ++chunk) {
      const int b = (int)(chunk & 1);
      int64_t m = std::min<int64_t>((int64_t)chunk_rows, n - r);
      const char* src = reinterpret_cast<const char*>(rows) + (size_t)r * s->dim * esz;
      if (chunk >= 2) VODB_CUDA_CHECK(cudaStreamWaitEvent(s->copy_stream, s->converted[b], 0));  // buffer free again
      VODB_CUDA_CHECK(cudaMemcpyAsync(s->stage[b], src, (size_t)m * s->dim * esz, cudaMemcpyHostToDevice, s->copy_stream));
      VODB_CUDA_CHECK(cudaEventRecord(s->copied[b], s->copy_stream));
      VODB_CUDA_CHECK(cudaStreamWaitEvent(st, s->copied[b], 0));
      int rc = launch_convert_rows(s->stage[b], src_dtype, s->dim, dst + (size_t)r * s->pitch * dtype_size(s->dtype),
                                   s->dtype, s->pitch, m, st);
      if (rc != VODB_OK) return rc;
      VODB_CUDA_CHECK(cudaEventRecord(s->converted[b], st));
    }
    VODB_CUDA_CHECK(cudaStreamSynchronize(st));  // host buffers are consumed when the call returns (include/vodb.h)
  }
  s->n_added = std::max(s->n_added, row0 + n);
  s->planes_rows_done = std::min(s->planes_rows_done, row0);  // bf16 planes of an fp32 store: redo from here
  return VODB_OK;
}

int vodb_store_fill_synthetic(vodb_store* s, uint64_t seed, int64_t row0, int64_t n, int unit_norm, void* stream) {
  VODB_REQUIRE(s != nullptr, "vodb_store_fill_synthetic: store is NULL");
  std::lock_guard<std::mutex> store_lock(s->mu);
  VODB_REQUIRE(n >= 0 && row0 >= 0 && row0 + n <= s->n_rows, "vodb_store_fill_synthetic: rows outside the store");
  VODB_REQUIRE(row0 <= s->n_added, "vodb_store_fill_synthetic: row0=%lld leaves a gap after the %lld rows added so far",
               (long long)row0, (long long)s->n_added);
  DeviceGuard guard(s->device);
  char* dst = reinterpret_cast<char*>(s->data) + (size_t)row0 * s->pitch * dtype_size(s->dtype);
  int rc = launch_fill_synthetic(dst, s->dtype, s->dim, s->pitch, seed, s->row_offset + row0, n, unit_norm,
                                 reinterpret_cast<cudaStream_t>(stream));
  if (rc != VODB_OK) return rc;
  s->n_added = std::max(s->n_added, row0 + n);
  s->planes_rows_done = std::min(s->planes_rows_done, row0);
  return VODB_OK;
}

int vodb_store_read(vodb_store* s, int64_t row0, int64_t n, float* out, int out_on_device, void* stream) {
  VODB_REQUIRE(s != nullptr && out != nullptr, "vodb_store_read: NULL argument");
  std::lock_guard<std::mutex> store_lock(s->mu);
  VODB_REQUIRE(n >= 0 && row0 >= 0 && row0 + n <= s->n_rows, "vodb_store_read: rows outside the store");
  if (n == 0) return VODB_OK;
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* src = reinterpret_cast<const char*>(s->data) + (size_t)row0 * s->pitch * dtype_size(s->dtype);
  if (out_on_device) return launch_read_rows(src, s->dtype, s->dim, s->pitch, n, out, st);
  float* tmp = nullptr;
  VODB_CUDA_CHECK(cudaMalloc(&tmp, (size_t)n * s->dim * sizeof(float)));
  int rc = launch_read_rows(src, s->dtype, s->dim, s->pitch, n, tmp, st);
  if (rc == VODB_OK) {
    cudaError_t e = cudaMemcpyAsync(out, tmp, (size_t)n * s->dim * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("vodb_store_read: %s", cudaGetErrorString(e));
      rc = VODB_ECUDA;
    }
  }
  cudaFree(tmp);
  return rc;
}

int64_t vodb_store_ntotal(const vodb_store* s) { return s ? s->n_added : 0; }
int vodb_store_dim(const vodb_store* s) { return s ? s->dim : 0; }
int vodb_store_dtype(const vodb_store* s) { return s ? s->dtype : -1; }
int vodb_store_device(const vodb_store* s) { return s ? s->device : -1; }
int64_t vodb_store_bytes(const vodb_store* s) {
  return s ? (int64_t)s->n_added * s->pitch * dtype_size(s->dtype) : 0;
}

int vodb_search(vodb_store* s, const void* queries, int q_dtype, int q_on_device, int nq, int k, int mode,
                float* out_scores, int64_t* out_idx, int out_on_device, void* stream) {
  int rc = check_search_args(s, queries, q_dtype, nq, k, mode, out_scores, out_idx, "vodb_search");
  if (rc != VODB_OK) return rc;
  std::lock_guard<std::mutex> store_lock(s->mu);
  if (nq == 0) return VODB_OK;
  if (s->n_added <= 0) {
    set_error("vodb_search: the store is empty (faiss health check: 'ERROR: Index is empty')");
    return VODB_ESTATE;
  }
  if (is_tensor_mode(mode) && !tensor_path_supported(s)) {
    set_error("vodb_search: VODB_MODE_TENSOR* needs a driver exporting cuTensorMapEncodeTiled");
    return VODB_EUNSUPPORTED;
  }
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = ensure_workspace(s, nq, k, dtype_size(q_dtype), st);
  if (rc != VODB_OK) return rc;
  Workspace& w = s->ws;
  const void* q_dev = nullptr;
  mode = effective_mode(s, queries, q_dtype, q_on_device, nq, mode);
  rc = upload_queries(s, queries, q_dtype, q_on_device, nq, st, &q_dev);
  if (rc != VODB_OK) return rc;

  const size_t nqk = (size_t)nq * k;
  int64_t* pack_i = reinterpret_cast<int64_t*>(w.out_pack);
  float* pack_s = reinterpret_cast<float*>(w.out_pack + nqk * 8);
  if (out_on_device) {
    // asynchronous: only enqueue. A list overflow leaves the sticky device flag set; the caller polls it
    // with vodb_search_check() (bench / pipelined callers) and re-runs synchronously if it fired.
    return run_scan(s, q_dev, q_dtype, nq, k, mode, /*safe=*/false, out_scores, out_idx, st);
  }
  // host outputs: ids, scores and the overflow flag come back with ONE device->host copy (pinned mirror) and ONE
  // synchronisation; if a list overflowed (rare) the batch is re-run on the overflow-proof schedule
  bool safe = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    rc = run_scan(s, q_dev, q_dtype, nq, k, mode, safe, pack_s, pack_i, st);
    if (rc != VODB_OK) return rc;
    VODB_CUDA_CHECK(cudaMemcpyAsync(w.out_pack + nqk * 12, w.overflow, sizeof(int), cudaMemcpyDeviceToDevice, st));
    VODB_CUDA_CHECK(cudaMemcpyAsync(w.out_host, w.out_pack, nqk * 12 + sizeof(int), cudaMemcpyDeviceToHost, st));
    VODB_CUDA_CHECK(cudaStreamSynchronize(st));
    int flag;
    std::memcpy(&flag, w.out_host + nqk * 12, sizeof(int));
    if (flag == 0) break;
    VODB_CUDA_CHECK(cudaMemsetAsync(w.overflow, 0, 2 * sizeof(int), st));
    if (safe) {
      set_error("vodb_search: candidate list overflow in safe mode (internal error)");
      return VODB_ESTATE;
    }
    safe = true;  // re-run with segments that cannot overflow
  }
  std::memcpy(out_idx, w.out_host, nqk * 8);
  std::memcpy(out_scores, w.out_host + nqk * 8, nqk * 4);
  return VODB_OK;
}

int vodb_xchg_create(vodb_xchg** out, int device, int rank, int world, int max_nq, int max_k,
                     unsigned char* handle_out) {
  VODB_REQUIRE(out != nullptr && handle_out != nullptr, "vodb_xchg_create: NULL argument");
  *out = nullptr;
  VODB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "vodb_xchg_create: bad rank %d / world %d (max %d)", rank, world, kMaxPeers);
  VODB_REQUIRE(max_nq >= 1 && max_k >= 1 && max_k <= VODB_MAX_K, "vodb_xchg_create: bad max_nq / max_k");
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("cudaSetDevice(%d) failed", device);
    return VODB_ECUDA;
  }
  vodb_xchg* x = new (std::nothrow) vodb_xchg();
  if (!x) return VODB_ENOMEM;
  x->device = device; x->rank = rank; x->world = world; x->max_nq = max_nq; x->max_k = max_k;
  x->slot = (size_t)max_nq * max_k;
  x->slot_words = x->slot * 3 + 8;
  x->bytes = (size_t)2 * world * x->slot_words * sizeof(uint64_t);
  cudaError_t e = cudaMalloc(&x->local, x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, x->local);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("vodb_xchg_create: %s", cudaGetErrorString(e));
    if (x->local) cudaFree(x->local);
    delete x;
    cudaGetLastError();
    return VODB_ECUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == VODB_IPC_HANDLE_BYTES, "IPC handle size");
  std::memcpy(handle_out, &h, VODB_IPC_HANDLE_BYTES);
  x->peer[rank] = x->local;
  *out = x;
  return VODB_OK;
}

int vodb_xchg_connect(vodb_xchg* x, const unsigned char* all_handles) {
  VODB_REQUIRE(x != nullptr && all_handles != nullptr, "vodb_xchg_connect: NULL argument");
  DeviceGuard guard(x->device);
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank || x->peer[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_handles + (size_t)r * VODB_IPC_HANDLE_BYTES, VODB_IPC_HANDLE_BYTES);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("vodb_xchg_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      return VODB_ECUDA;
    }
    x->peer[r] = reinterpret_cast<char*>(ptr);
  }
  x->connected = true;
  return VODB_OK;
}

void vodb_xchg_destroy(vodb_xchg* x) {
  if (!x) return;
  DeviceGuard guard(x->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->local) cudaFree(x->local);
  delete x;
}

int vodb_search_sharded(vodb_store* s, vodb_xchg* x, const void* queries, int q_dtype, int q_on_device, int nq, int k,
                        int mode, int safe, float* out_scores, int64_t* out_idx, int out_on_device, void* stream) {
  int rc = check_search_args(s, queries, q_dtype, nq, k, mode, out_scores, out_idx, "vodb_search_sharded");
  if (rc != VODB_OK) return rc;
  std::lock_guard<std::mutex> store_lock(s->mu);
  VODB_REQUIRE(x != nullptr && x->connected, "vodb_search_sharded: exchange is NULL or not connected");
  VODB_REQUIRE(x->device == s->device, "vodb_search_sharded: exchange and store live on different devices");
  VODB_REQUIRE(nq >= 1 && (size_t)nq * k <= x->slot, "vodb_search_sharded: nq*k=%lld exceeds the exchange slot (%zu)", (long long)nq * k, x->slot);
  if (is_tensor_mode(mode) && !tensor_path_supported(s)) {
    set_error("vodb_search_sharded: VODB_MODE_TENSOR* needs a driver exporting cuTensorMapEncodeTiled");
    return VODB_EUNSUPPORTED;
  }
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = ensure_workspace(s, nq, k, dtype_size(q_dtype), st);
  if (rc != VODB_OK) return rc;
  Workspace& w = s->ws;
  const void* q_dev = nullptr;
  mode = effective_mode(s, queries, q_dtype, q_on_device, nq, mode);
  rc = upload_queries(s, queries, q_dtype, q_on_device, nq, st, &q_dev);
  if (rc != VODB_OK) return rc;

  const size_t nqk = (size_t)nq * k;
  float* o_s = out_on_device ? out_scores : reinterpret_cast<float*>(w.out_pack + nqk * 8);
  int64_t* o_i = out_on_device ? out_idx : reinterpret_cast<int64_t*>(w.out_pack);
  bool safe_run = safe != 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    // every rank calls this the same number of times: the epoch (and its parity = buffer half) stay in lockstep
    x->epoch += 1;
    if (x->epoch == 0) x->epoch = 2;  // 32-bit wrap: 0 is the tag of a zero-initialised buffer; 2 keeps the parity sequence
    const int parity = (int)(x->epoch & 1u);
    ExchangeDst xd{};
    xd.world = x->world;
    xd.rank = x->rank;
    xd.epoch = x->epoch;
    xd.overflow = w.overflow;
    xd.slot_elems = x->slot;
    for (int r = 0; r < x->world; ++r) {
      xd.peer_ll[r] = reinterpret_cast<uint64_t*>(x->peer[r]) + ((size_t)parity * x->world + x->rank) * x->slot_words;
      xd.peer_flag[r] = xd.peer_ll[r] + x->slot * 3;
    }
    if (s->n_added > 0) {
      rc = run_scan(s, q_dev, q_dtype, nq, k, mode, safe_run, nullptr, nullptr, st, &xd);
    } else {
      // an empty shard (more ranks than row blocks) contributes an all-padding list
      VODB_CUDA_CHECK(cudaMemsetAsync(w.cnt, 0, (size_t)nq * kCntStride * sizeof(int), st));
      rc = launch_select(w.cand_s, w.cand_i, w.cnt, w.tau, w.cap, nq, k, true, nullptr, nullptr, s->row_offset, st, &xd);
    }
    if (rc != VODB_OK) return rc;
    const uint64_t* gll = reinterpret_cast<const uint64_t*>(x->local) + (size_t)parity * x->world * x->slot_words;
    rc = launch_merge_exchange(gll, x->epoch, x->world, x->slot_words, x->slot * 3, nq, k, o_s, o_i, w.overflow + 1, st);
    if (rc != VODB_OK) return rc;
    if (out_on_device) return VODB_OK;  // asynchronous: the caller polls vodb_search_check (same answer on every rank)
    // host outputs: results + the all-shard overflow flag in ONE copy; the flag is the OR over every rank's shard, so
    // all ranks take the same branch here and the re-run on the overflow-proof schedule stays in lockstep
    VODB_CUDA_CHECK(cudaMemcpyAsync(w.out_pack + nqk * 12, w.overflow + 1, sizeof(int), cudaMemcpyDeviceToDevice, st));
    VODB_CUDA_CHECK(cudaMemcpyAsync(w.out_host, w.out_pack, nqk * 12 + sizeof(int), cudaMemcpyDeviceToHost, st));
    VODB_CUDA_CHECK(cudaStreamSynchronize(st));
    int flag;
    std::memcpy(&flag, w.out_host + nqk * 12, sizeof(int));
    if (flag == 0) break;
    VODB_CUDA_CHECK(cudaMemsetAsync(w.overflow, 0, 2 * sizeof(int), st));
    if (flag & 2) {
      set_error("vodb_search_sharded: the ranks did not pass the same batch shape (nq=%d, k=%d here): every rank must call "
                "with the same queries, in the same order", nq, k);
      return VODB_ESTATE;
    }
    if (safe_run) {
      set_error("vodb_search_sharded: candidate list overflow in safe mode (internal error)");
      return VODB_ESTATE;
    }
    safe_run = true;
  }
  std::memcpy(out_idx, w.out_host, nqk * 8);
  std::memcpy(out_scores, w.out_host + nqk * 8, nqk * 4);
  return VODB_OK;
}

int vodb_store_prepare_tensor(vodb_store* s, void* stream) {
  VODB_REQUIRE(s != nullptr, "vodb_store_prepare_tensor: store is NULL");
  std::lock_guard<std::mutex> store_lock(s->mu);
  if (!tensor_path_supported(s)) {
    set_error("vodb_store_prepare_tensor: the driver does not export cuTensorMapEncodeTiled");
    return VODB_EUNSUPPORTED;
  }
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = order_streams(s, st);
  if (rc != VODB_OK) return rc;
  return ensure_planes(s, st);  // no-op for 16-bit stores; VODB_ENOMEM when the planes do not fit
}

int vodb_search_check(vodb_store* s, void* stream) {
  VODB_REQUIRE(s != nullptr, "vodb_search_check: store is NULL");
  std::lock_guard<std::mutex> store_lock(s->mu);
  Workspace& w = s->ws;
  if (!w.overflow) return 0;
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  VODB_CUDA_CHECK(cudaMemcpyAsync(w.overflow_host, w.overflow, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VODB_CUDA_CHECK(cudaMemsetAsync(w.overflow, 0, 2 * sizeof(int), st));
  VODB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (w.overflow_host[1] & 2) {
    set_error("vodb_search_check: a sharded search ran with different batch shapes (nq, k) on different ranks");
    return VODB_ESTATE;
  }
  return (w.overflow_host[0] | w.overflow_host[1]) != 0 ? 1 : 0;
}

int vodb_store_set_profiling(vodb_store* s, int enable) {
  VODB_REQUIRE(s != nullptr, "vodb_store_set_profiling: store is NULL");
  std::lock_guard<std::mutex> store_lock(s->mu);
  if (!s->prof) s->prof = new ProfileState();
  static_cast<ProfileState*>(s->prof)->reset();
  s->profiling = enable != 0;
  return VODB_OK;
}

int vodb_store_profile(vodb_store* s, double out[4]) {
  VODB_REQUIRE(s != nullptr && out != nullptr, "vodb_store_profile: NULL argument");
  std::lock_guard<std::mutex> store_lock(s->mu);
  out[0] = out[1] = out[2] = out[3] = 0.0;
  if (!s->prof) return VODB_OK;
  DeviceGuard guard(s->device);
  VODB_CUDA_CHECK(cudaDeviceSynchronize());
  ProfileState* p = static_cast<ProfileState*>(s->prof);
  for (size_t j = 0; j < p->kinds.size(); ++j) {
    float ms = 0.f;
    VODB_CUDA_CHECK(cudaEventElapsedTime(&ms, p->pool[2 * j], p->pool[2 * j + 1]));
    if (p->kinds[j] == 0) {
      out[0] += ms;
      out[2] += 1.0;
    } else {
      out[1] += ms;
    }
  }
  out[3] = p->bytes;
  p->reset();
  return VODB_OK;
}

int vodb_plan_scan(int64_t n_rows, int nq, int k, int safe, int* out_cap, int64_t* out_bounds, int max_bounds) {
  VODB_REQUIRE(n_rows >= 1 && nq >= 1 && k >= 1 && k <= VODB_MAX_K, "vodb_plan_scan: bad sizes");
  VODB_REQUIRE(out_cap != nullptr && out_bounds != nullptr && max_bounds >= 2, "vodb_plan_scan: bad output arguments");
  const int cap = choose_cap(nq, k);
  const std::vector<int64_t> b = plan_segments(n_rows, cap, k, nq, safe != 0);
  *out_cap = cap;
  const int n = (int)std::min<size_t>(b.size(), (size_t)max_bounds);
  std::memcpy(out_bounds, b.data(), (size_t)n * sizeof(int64_t));
  return (int)b.size();
}

int vodb_search_stats(const vodb_store* s, int64_t out[8]) {
  VODB_REQUIRE(s != nullptr && out != nullptr, "vodb_search_stats: NULL argument");
  std::memcpy(out, s->stats, sizeof(s->stats));
  return VODB_OK;
}

int vodb_merge_topk(int device, const float* scores, const int64_t* idx, int n_lists, int nq, int k_in, int k_out,
                    float* out_scores, int64_t* out_idx, int on_device, void* stream) {
  VODB_REQUIRE(n_lists >= 1 && nq >= 0 && k_in >= 1 && k_out >= 1 && k_out <= VODB_MAX_K, "vodb_merge_topk: bad sizes");
  VODB_REQUIRE((int64_t)n_lists * k_in <= (1 << 20), "vodb_merge_topk: n_lists*k_in too large");
  if (nq == 0) return VODB_OK;
  VODB_REQUIRE(scores && idx && out_scores && out_idx, "vodb_merge_topk: NULL pointer");
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("cudaSetDevice(%d) failed", device);
    return VODB_ECUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (on_device) return launch_merge(scores, idx, n_lists, nq, k_in, k_out, out_scores, out_idx, st);
  size_t n_in = (size_t)n_lists * nq * k_in, n_out = (size_t)nq * k_out;
  CallScratch& cs = call_scratch(device);
  std::lock_guard<std::mutex> lock(cs.mu);
  char* d = scratch_reserve(cs, (n_in + n_out) * 12 + 64, "vodb_merge_topk");
  if (!d) return VODB_ENOMEM;
  int64_t* d_i = reinterpret_cast<int64_t*>(d);
  int64_t* d_oi = d_i + n_in;
  float* d_s = reinterpret_cast<float*>(d_oi + n_out);
  float* d_os = d_s + n_in;
  int rc = VODB_OK;
  cudaError_t e = cudaMemcpyAsync(d_s, scores, n_in * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_i, idx, n_in * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) rc = launch_merge(d_s, d_i, n_lists, nq, k_in, k_out, d_os, d_oi, st);
  if (e == cudaSuccess && rc == VODB_OK) e = cudaMemcpyAsync(out_scores, d_os, n_out * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == VODB_OK) e = cudaMemcpyAsync(out_idx, d_oi, n_out * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("vodb_merge_topk: %s", cudaGetErrorString(e));
    return VODB_ECUDA;
  }
  return rc;
}

int vodb_merge_results(int device, int n_engines, const void* const* scores, const int64_t* const* indices,
                       const int64_t* const* labels, const int* widths, const double* weights, const int* zero_scores,
                       int B, int score_dtype, int normalize, double offset, int label_engine, int out_width,
                       void* out_scores, int64_t* out_indices, int64_t* out_labels, void* out_raw, int* out_counts,
                       int on_device, void* stream) {
  VODB_REQUIRE(n_engines >= 1 && n_engines <= 8, "vodb_merge_results: n_engines=%d outside [1, 8]", n_engines);
  VODB_REQUIRE(B >= 0 && out_width >= 1, "vodb_merge_results: bad B / out_width");
  VODB_REQUIRE(score_dtype == VODB_F32 || score_dtype == 3 /* VODB_F64 */, "vodb_merge_results: score dtype must be f32 (0) or f64 (3)");
  VODB_REQUIRE(scores && indices && widths && weights, "vodb_merge_results: NULL pointer table");
  VODB_REQUIRE(label_engine >= -1 && label_engine < n_engines, "vodb_merge_results: bad label_engine");
  VODB_REQUIRE(label_engine < 0 || (labels && labels[label_engine] && out_labels), "vodb_merge_results: label engine without labels");
  if (B == 0) return VODB_OK;
  VODB_REQUIRE(out_scores && out_indices && out_raw && out_counts, "vodb_merge_results: NULL output");
  const int is_f64 = score_dtype != VODB_F32;
  const size_t fsz = is_f64 ? 8 : 4;
  int64_t M = 0;
  for (int e = 0; e < n_engines; ++e) {
    VODB_REQUIRE(widths[e] >= 0 && scores[e] && indices[e], "vodb_merge_results: engine %d has NULL arrays", e);
    M += widths[e];
  }
  VODB_REQUIRE(M >= 1 && M <= 8192, "vodb_merge_results: %lld entries per row outside [1, 8192]", (long long)M);
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("cudaSetDevice(%d) failed", device);
    return VODB_ECUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (on_device)
    return launch_merge_results(n_engines, scores, indices, labels, widths, weights, zero_scores, B, is_f64, normalize,
                                offset, label_engine, out_width, out_scores, out_indices, out_labels, out_raw,
                                out_counts, st);
  // host buffers: one packed device allocation for inputs and outputs
  size_t in_bytes = 0;
  for (int e = 0; e < n_engines; ++e) in_bytes += (size_t)B * widths[e] * (fsz + 8 + 8);
  const size_t nout = (size_t)B * out_width;
  size_t out_bytes = nout * (fsz + 8 + 8) + (size_t)n_engines * nout * fsz + (size_t)B * 4 + 64;
  CallScratch& cs = call_scratch(device);
  std::lock_guard<std::mutex> lock(cs.mu);
  char* d = scratch_reserve(cs, in_bytes + out_bytes + 256, "vodb_merge_results");
  if (!d) return VODB_ENOMEM;
  const void* d_scores[8];
  const int64_t* d_idx[8];
  const int64_t* d_lab[8];
  char* cur = d;
  cudaError_t e2 = cudaSuccess;
  for (int e = 0; e < n_engines && e2 == cudaSuccess; ++e) {
    size_t n = (size_t)B * widths[e];
    d_idx[e] = reinterpret_cast<int64_t*>(cur);
    e2 = cudaMemcpyAsync(cur, indices[e], n * 8, cudaMemcpyHostToDevice, st);
    cur += n * 8;
    d_lab[e] = nullptr;
    if (e2 == cudaSuccess && labels && labels[e]) {
      d_lab[e] = reinterpret_cast<int64_t*>(cur);
      e2 = cudaMemcpyAsync(cur, labels[e], n * 8, cudaMemcpyHostToDevice, st);
    }
    cur += n * 8;
    d_scores[e] = cur;
    if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(cur, scores[e], n * fsz, cudaMemcpyHostToDevice, st);
    cur += (n * fsz + 7) / 8 * 8;
  }
  int64_t* o_i = reinterpret_cast<int64_t*>(cur); cur += nout * 8;
  int64_t* o_l = reinterpret_cast<int64_t*>(cur); cur += nout * 8;
  char* o_s = cur; cur += (nout * fsz + 7) / 8 * 8;
  char* o_r = cur; cur += ((size_t)n_engines * nout * fsz + 7) / 8 * 8;
  int* o_c = reinterpret_cast<int*>(cur);
  int rc = VODB_OK;
  if (e2 == cudaSuccess)
    rc = launch_merge_results(n_engines, d_scores, d_idx, d_lab, widths, weights, zero_scores, B, is_f64, normalize,
                              offset, label_engine, out_width, o_s, o_i, out_labels ? o_l : nullptr, o_r, o_c, st);
  if (e2 == cudaSuccess && rc == VODB_OK) e2 = cudaMemcpyAsync(out_scores, o_s, nout * fsz, cudaMemcpyDeviceToHost, st);
  if (e2 == cudaSuccess && rc == VODB_OK) e2 = cudaMemcpyAsync(out_indices, o_i, nout * 8, cudaMemcpyDeviceToHost, st);
  if (e2 == cudaSuccess && rc == VODB_OK && out_labels) e2 = cudaMemcpyAsync(out_labels, o_l, nout * 8, cudaMemcpyDeviceToHost, st);
  if (e2 == cudaSuccess && rc == VODB_OK) e2 = cudaMemcpyAsync(out_raw, o_r, (size_t)n_engines * nout * fsz, cudaMemcpyDeviceToHost, st);
  if (e2 == cudaSuccess && rc == VODB_OK) e2 = cudaMemcpyAsync(out_counts, o_c, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
  if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
  if (e2 != cudaSuccess) {
    set_error("vodb_merge_results: %s", cudaGetErrorString(e2));
    return VODB_ECUDA;
  }
  return rc;
}

int vodb_sample(int device, const float* scores, const uint8_t* labels, const float* noise, int B, int K,
                int k_positive, int k_total, int normalized, float temperature, int max_support, int quirks,
                uint64_t seed, uint64_t offset, int64_t* out_ids, float* out_logw, uint8_t* out_labels,
                float* out_lse, int on_device, void* stream) {
  VODB_REQUIRE(B >= 0 && K >= 0, "vodb_sample: negative shape");
  VODB_REQUIRE(K <= 8192, "vodb_sample: K=%d > 8192", K);
  VODB_REQUIRE(k_total >= 0 && k_positive >= 0 && k_positive <= k_total, "vodb_sample: need 0 <= k_positive <= k_total (got %d, %d)", k_positive, k_total);
  if (B == 0) return VODB_OK;
  VODB_REQUIRE(scores && out_ids && out_logw && out_labels && out_lse, "vodb_sample: NULL pointer");
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("cudaSetDevice(%d) failed", device);
    return VODB_ECUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (on_device)
    return launch_sample(scores, labels, noise, B, K, k_positive, k_total, normalized, temperature, max_support,
                         quirks, seed, offset, out_ids, out_logw, out_labels, out_lse, st);
  // host buffers: one packed device allocation
  size_t nBK = (size_t)B * K, nBk = (size_t)B * k_total;
  size_t off_scores = 0, off_noise = off_scores + nBK * 4, off_logw = off_noise + (noise ? nBK * 4 : 0);
  size_t off_lse = off_logw + nBk * 4, off_ids = (off_lse + (size_t)B * 8 + 7) / 8 * 8;
  size_t off_labels = off_ids + nBk * 8, off_olab = off_labels + (labels ? nBK : 0), total = off_olab + nBk + 16;
  CallScratch& cs = call_scratch(device);
  std::lock_guard<std::mutex> lock(cs.mu);
  char* d = scratch_reserve(cs, total, "vodb_sample");
  if (!d) return VODB_ENOMEM;
  int rc = VODB_OK;
  cudaError_t e = cudaMemcpyAsync(d + off_scores, scores, nBK * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && noise) e = cudaMemcpyAsync(d + off_noise, noise, nBK * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && labels) e = cudaMemcpyAsync(d + off_labels, labels, nBK, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    rc = launch_sample((const float*)(d + off_scores), labels ? (const uint8_t*)(d + off_labels) : nullptr,
                       noise ? (const float*)(d + off_noise) : nullptr, B, K, k_positive, k_total, normalized,
                       temperature, max_support, quirks, seed, offset, (int64_t*)(d + off_ids), (float*)(d + off_logw),
                       (uint8_t*)(d + off_olab), (float*)(d + off_lse), st);
  if (e == cudaSuccess && rc == VODB_OK && nBk) {
    e = cudaMemcpyAsync(out_ids, d + off_ids, nBk * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_logw, d + off_logw, nBk * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_labels, d + off_olab, nBk, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess && rc == VODB_OK) e = cudaMemcpyAsync(out_lse, d + off_lse, (size_t)B * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("vodb_sample: %s", cudaGetErrorString(e));
    return VODB_ECUDA;
  }
  return rc;
}

// sample_search_results (core/sample.py:22-84) for host arrays in ONE call: labeled priority sampling of the [B,K]
// retrieved scores, then the gathers of ids / scores at the picks and max_sampling_id (sample.py:57-71) by the same
// gather kernel the retrieve->sample chain uses. One packed upload, one packed download.
int vodb_sample_results(int device, const float* scores, const int64_t* indices, const uint8_t* labels,
                        const float* noise, int B, int K, int k_positive, int k_total, float temperature, int max_support, int quirks, uint64_t seed,
                        uint64_t offset, int64_t* out_idx, float* out_scores, float* out_logw, uint8_t* out_labels,
                        float* out_lse, float* out_msid, int64_t* out_local, void* stream) {
  VODB_REQUIRE(B >= 0 && K >= 1 && K <= 8192, "vodb_sample_results: bad shape [%d, %d] (1 <= K <= 8192)", B, K);
  VODB_REQUIRE(k_total >= 0 && k_positive >= 0 && k_positive <= k_total,
               "vodb_sample_results: need 0 <= k_positive <= k_total (got %d, %d)", k_positive, k_total);
  if (B == 0) return VODB_OK;
  VODB_REQUIRE(scores && indices && out_idx && out_scores && out_logw && out_labels && out_lse && out_msid,
               "vodb_sample_results: NULL pointer");
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("cudaSetDevice(%d) failed", device);
    return VODB_ECUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t nBK = (size_t)B * K, nBk = (size_t)B * k_total;
  // inputs [indices i64 | scores f32 | labels u8], outputs [idx i64 | local i64 | scores f32 | logw f32 | lse | msid | labels u8]
  const size_t i_idx = 0, i_sc = i_idx + nBK * 8, i_noise = i_sc + nBK * 4, i_lab = i_noise + (noise ? nBK * 4 : 0);
  const size_t in_bytes = (i_lab + (labels ? nBK : 0) + 7) / 8 * 8;
  const size_t o_idx = in_bytes, o_local = o_idx + nBk * 8, o_sc = o_local + nBk * 8, o_logw = o_sc + nBk * 4;
  const size_t o_lse = o_logw + nBk * 4, o_msid = o_lse + (size_t)B * 8, o_olab = o_msid + (size_t)B * 4;
  const size_t total = o_olab + nBk + 16;
  CallScratch& cs = call_scratch(device);
  std::lock_guard<std::mutex> lock(cs.mu);
  char* d = scratch_reserve(cs, total, "vodb_sample_results");
  if (!d) return VODB_ENOMEM;
  int max_sup = max_support;
  if (max_sup >= 0 && max_sup < k_total) max_sup = k_total;  // sample.py:133-135
  int rc = VODB_OK;
  cudaError_t e = cudaMemcpyAsync(d + i_idx, indices, nBK * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + i_sc, scores, nBK * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && labels) e = cudaMemcpyAsync(d + i_lab, labels, nBK, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && noise) e = cudaMemcpyAsync(d + i_noise, noise, nBK * 4, cudaMemcpyHostToDevice, st);
  const uint8_t* d_lab = labels ? reinterpret_cast<const uint8_t*>(d + i_lab) : nullptr;
  const float* d_noise = noise ? reinterpret_cast<const float*>(d + i_noise) : nullptr;
  if (e == cudaSuccess)
    rc = launch_sample(reinterpret_cast<const float*>(d + i_sc), d_lab, d_noise, B, K, k_positive, k_total,
                       /*normalized=*/1, temperature, max_sup, quirks, seed, offset, reinterpret_cast<int64_t*>(d + o_local),
                       reinterpret_cast<float*>(d + o_logw), reinterpret_cast<uint8_t*>(d + o_olab),
                       reinterpret_cast<float*>(d + o_lse), st);
  if (e == cudaSuccess && rc == VODB_OK)
    rc = launch_gather_picks(reinterpret_cast<const float*>(d + i_sc), reinterpret_cast<const int64_t*>(d + i_idx), d_lab, B,
                             K, k_total, reinterpret_cast<const int64_t*>(d + o_local),
                             reinterpret_cast<const uint8_t*>(d + o_olab), reinterpret_cast<int64_t*>(d + o_idx),
                             reinterpret_cast<float*>(d + o_sc), reinterpret_cast<float*>(d + o_msid), st);
  // outputs are contiguous on the device: one copy into the pinned-or-pageable host block would need a host staging
  // buffer; the seven arrays are small ([B,k_total]), so they are copied one by one and synchronised once
  if (e == cudaSuccess && rc == VODB_OK && nBk) {
    e = cudaMemcpyAsync(out_idx, d + o_idx, nBk * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out_local) e = cudaMemcpyAsync(out_local, d + o_local, nBk * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, d + o_sc, nBk * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_logw, d + o_logw, nBk * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_labels, d + o_olab, nBk, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess && rc == VODB_OK) e = cudaMemcpyAsync(out_lse, d + o_lse, (size_t)B * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == VODB_OK) e = cudaMemcpyAsync(out_msid, d + o_msid, (size_t)B * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("vodb_sample_results: %s", cudaGetErrorString(e));
    return VODB_ECUDA;
  }
  return rc;
}

// The RealmCollate chain for the dense-only flow in ONE call (realm_collate.py:101-122 -> core/sample.py:22-84):
// search top_k -> label the retrieved ids against the gold ids -> labeled priority sampling of k_total ->
// gather ids / scores at the picks + max_sampling_id. Only the [nq, k_total] result crosses PCIe (one packed copy).
int vodb_retrieve_sample(vodb_store* s, const void* queries, int q_dtype, int q_on_device, int nq, int top_k, int mode,
                         const int64_t* gold_ids, int n_gold, int k_positive, int k_total, float temperature,
                         int max_support, int quirks, uint64_t seed, uint64_t offset, int64_t* out_idx,
                         float* out_scores, float* out_logw, uint8_t* out_labels, float* out_lse, float* out_msid,
                         int64_t* out_local, void* stream) {
  int rc = check_search_args(s, queries, q_dtype, nq, top_k, mode, out_scores, out_idx, "vodb_retrieve_sample");
  if (rc != VODB_OK) return rc;
  std::lock_guard<std::mutex> store_lock(s->mu);
  VODB_REQUIRE(top_k <= 8192, "vodb_retrieve_sample: top_k=%d > 8192", top_k);
  VODB_REQUIRE(k_total >= 0 && k_positive >= 0 && k_positive <= k_total,
               "vodb_retrieve_sample: need 0 <= k_positive <= k_total (got %d, %d)", k_positive, k_total);
  VODB_REQUIRE(n_gold >= 0 && (n_gold == 0 || gold_ids != nullptr), "vodb_retrieve_sample: gold_ids is NULL");
  if (nq == 0) return VODB_OK;
  VODB_REQUIRE(out_logw && out_labels && out_lse && out_msid, "vodb_retrieve_sample: output pointer is NULL");
  if (s->n_added <= 0) {
    set_error("vodb_retrieve_sample: the store is empty");
    return VODB_ESTATE;
  }
  if (is_tensor_mode(mode) && !tensor_path_supported(s)) {
    set_error("vodb_retrieve_sample: VODB_MODE_TENSOR* needs a driver exporting cuTensorMapEncodeTiled");
    return VODB_EUNSUPPORTED;
  }
  DeviceGuard guard(s->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  rc = ensure_workspace(s, nq, top_k, dtype_size(q_dtype), st);
  if (rc != VODB_OK) return rc;
  Workspace& w = s->ws;

  // chain scratch: [result pack: idx i64 | local i64 | scores f32 | logw f32 | lse f32 x2 | msid f32 | overflow | labels u8]
  // followed by device-only parts [gold ids i64 | retrieved labels u8]
  const size_t nkt = (size_t)nq * k_total, nK = (size_t)nq * top_k;
  const size_t o_idx = 0, o_local = o_idx + nkt * 8, o_scores = o_local + nkt * 8, o_logw = o_scores + nkt * 4;
  const size_t o_lse = o_logw + nkt * 4, o_msid = o_lse + (size_t)nq * 8, o_flag = o_msid + (size_t)nq * 4;
  const size_t o_olab = o_flag + 4, pack_bytes = o_olab + nkt;
  const size_t o_gold = (pack_bytes + 7) / 8 * 8, o_labels = o_gold + (size_t)nq * n_gold * 8, need = o_labels + nK + 16;
  if (need > w.chain_bytes) {
    if (w.chain_dev) cudaFree(w.chain_dev);
    if (w.chain_host) cudaFreeHost(w.chain_host);
    w.chain_dev = nullptr; w.chain_host = nullptr; w.chain_bytes = 0;
    VODB_CUDA_CHECK(cudaMalloc(&w.chain_dev, need));
    VODB_CUDA_CHECK(cudaMallocHost(&w.chain_host, need));
    w.chain_bytes = need;
  }
  char* d = w.chain_dev;
  const void* q_dev = nullptr;
  mode = effective_mode(s, queries, q_dtype, q_on_device, nq, mode);
  rc = upload_queries(s, queries, q_dtype, q_on_device, nq, st, &q_dev);
  if (rc != VODB_OK) return rc;
  uint8_t* labels = nullptr;
  if (n_gold > 0) {
    std::memcpy(w.chain_host + o_gold, gold_ids, (size_t)nq * n_gold * 8);  // pinned staging: the copy below is async
    VODB_CUDA_CHECK(cudaMemcpyAsync(d + o_gold, w.chain_host + o_gold, (size_t)nq * n_gold * 8, cudaMemcpyHostToDevice, st));
    labels = reinterpret_cast<uint8_t*>(d + o_labels);
  }
  int64_t* top_i = reinterpret_cast<int64_t*>(w.out_pack);
  float* top_s = reinterpret_cast<float*>(w.out_pack + nK * 8);
  int max_sup = max_support;
  if (max_sup >= 0 && max_sup < k_total) max_sup = k_total;  // sample.py:133-135

  bool safe = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    rc = run_scan(s, q_dev, q_dtype, nq, top_k, mode, safe, top_s, top_i, st);
    if (rc != VODB_OK) return rc;
    if (labels) {
      rc = launch_match_labels(top_i, reinterpret_cast<const int64_t*>(d + o_gold), n_gold, nq, top_k, labels, st);
      if (rc != VODB_OK) return rc;
    }
    rc = launch_sample(top_s, labels, nullptr, nq, top_k, k_positive, k_total, /*normalized=*/1, temperature, max_sup,
                       quirks, seed, offset, reinterpret_cast<int64_t*>(d + o_local), reinterpret_cast<float*>(d + o_logw),
                       reinterpret_cast<uint8_t*>(d + o_olab), reinterpret_cast<float*>(d + o_lse), st);
    if (rc != VODB_OK) return rc;
    rc = launch_gather_picks(top_s, top_i, labels, nq, top_k, k_total, reinterpret_cast<const int64_t*>(d + o_local),
                             reinterpret_cast<const uint8_t*>(d + o_olab), reinterpret_cast<int64_t*>(d + o_idx),
                             reinterpret_cast<float*>(d + o_scores), reinterpret_cast<float*>(d + o_msid), st);
    if (rc != VODB_OK) return rc;
    VODB_CUDA_CHECK(cudaMemcpyAsync(d + o_flag, w.overflow, sizeof(int), cudaMemcpyDeviceToDevice, st));
    VODB_CUDA_CHECK(cudaMemcpyAsync(w.chain_host, d, pack_bytes, cudaMemcpyDeviceToHost, st));
    VODB_CUDA_CHECK(cudaStreamSynchronize(st));
    int flag;
    std::memcpy(&flag, w.chain_host + o_flag, sizeof(int));
    if (flag == 0) break;
    VODB_CUDA_CHECK(cudaMemsetAsync(w.overflow, 0, 2 * sizeof(int), st));
    if (safe) {
      set_error("vodb_retrieve_sample: candidate list overflow in safe mode (internal error)");
      return VODB_ESTATE;
    }
    safe = true;
  }
  const char* h = w.chain_host;
  std::memcpy(out_idx, h + o_idx, nkt * 8);
  if (out_local) std::memcpy(out_local, h + o_local, nkt * 8);
  std::memcpy(out_scores, h + o_scores, nkt * 4);
  std::memcpy(out_logw, h + o_logw, nkt * 4);
  std::memcpy(out_lse, h + o_lse, (size_t)nq * 8);
  std::memcpy(out_msid, h + o_msid, (size_t)nq * 4);
  std::memcpy(out_labels, h + o_olab, nkt);
  return VODB_OK;
}

}  // extern "C"
