// merge_results.cu — score normalisation + weighted union of several engines' results + raw-score / label gather.
//
// The step between search and sampling in the reference's collate, one CTA per query row:
//   _subtract_min_score          src/vod_dataloaders/core/normalize.py:17-20   s - min_finite_row(s) + offset
//   result * weight              src/vod_types/retrieval.py:222-233
//   _nopy_merge_two_search_results / _write_1d_arr / _search_1d_arr
//                                src/vod_dataloaders/core/merge.py:71-105, 108-164   union by index in insertion
//                                order, duplicate -> scores added in entry order, negative ids skipped
//   gather_values_by_indices     src/vod_dataloaders/core/numpy_ops.py:24-143  per-engine raw scores (NaN if absent)
//                                and labels (-1 if absent) of every output slot, first match wins
// The reference does this with O(K^2) linear probes per row in numba; here the entries (engine-major, then column)
// are sorted by (id, position) with a bitonic network in shared memory, each run of equal ids is summed by its head
// thread in position order (the reference's left-to-right float additions, so the result is bit-identical), and
// the output slot of a run is the rank of its first position (block prefix sum), i.e. first-occurrence order.
#include "common.cuh"

namespace vodb {

namespace {

constexpr int kThreads = 512;
constexpr int kMaxEngines = 8;

struct MergeArgs {
  int n_engines;
  int B;
  int out_width;
  int normalize;
  int label_engine;  // engine whose labels are gathered (-1: none)
  int width[kMaxEngines];
  int base[kMaxEngines + 1];  // prefix sums of widths (entry positions)
  int zero_scores[kMaxEngines];  // scores treated as 0 (the lookup engine, core/search.py:92)
  double weight[kMaxEngines];
  double offset;
  const void* scores[kMaxEngines];
  const int64_t* indices[kMaxEngines];
  const int64_t* labels[kMaxEngines];
};

template <typename F>
__device__ __forceinline__ F f_inf();
template <>
__device__ __forceinline__ float f_inf<float>() { return __int_as_float(0x7f800000); }
template <>
__device__ __forceinline__ double f_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <typename F>
__device__ __forceinline__ F f_nan();
template <>
__device__ __forceinline__ float f_nan<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double f_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// ascending bitonic sort of (key, pos) pairs; P power of two
__device__ void sort_pairs(uint64_t* key, int32_t* pos, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = (lo & size) == 0;
        uint64_t ka = key[lo], kb = key[hi];
        int32_t pa = pos[lo], pb = pos[hi];
        bool a_gt_b = (ka > kb) || (ka == kb && pa > pb);
        if (a_gt_b == up) {
          key[lo] = kb; key[hi] = ka;
          pos[lo] = pb; pos[hi] = pa;
        }
      }
    }
  }
  __syncthreads();
}

template <typename F>
__global__ void __launch_bounds__(kThreads)
merge_results_kernel(const MergeArgs a, int M, int P, F* __restrict__ out_scores, int64_t* __restrict__ out_indices,
                     int64_t* __restrict__ out_labels, F* __restrict__ out_raw, int* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char dyn[];
  uint64_t* key = reinterpret_cast<uint64_t*>(dyn);               // [P]
  F* norm = reinterpret_cast<F*>(key + P);                        // [M]  normalised (unweighted) scores by position
  int32_t* pos = reinterpret_cast<int32_t*>(norm + M + (M & 1));  // [P]
  int32_t* slot = pos + P;                                        // [M]  flag -> output slot of a run's first position
  __shared__ F row_min[kMaxEngines];
  __shared__ int scan_carry;
  __shared__ int warp_sums[kThreads / 32];

  const int row = blockIdx.x;
  const int tid = threadIdx.x;

  // 1. per-engine minimum over the finite scores of this row (normalize.py:17-20)
  if (tid < kMaxEngines) row_min[tid] = f_inf<F>();
  __syncthreads();
  if (a.normalize) {
    for (int e = 0; e < a.n_engines; ++e) {
      if (a.zero_scores[e]) {
        if (tid == 0) row_min[e] = a.width[e] > 0 ? (F)0 : f_inf<F>();
        continue;
      }
      const F* s = reinterpret_cast<const F*>(a.scores[e]) + (size_t)row * a.width[e];
      F m = f_inf<F>();
      for (int j = tid; j < a.width[e]; j += blockDim.x) {
        F v = s[j];
        if (!(isinf(v) || isnan(v))) m = fmin(m, v);
      }
      // block min: every finite value compares exactly, so a shuffle tree + one shared slot per warp is enough
      for (int off = 16; off > 0; off >>= 1) {
        F o = __shfl_xor_sync(0xffffffffu, m, off);
        m = fmin(m, o);
      }
      __shared__ F warp_min[kThreads / 32];
      if ((tid & 31) == 0) warp_min[tid >> 5] = m;
      __syncthreads();
      if (tid == 0) {
        F mm = f_inf<F>();
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmin(mm, warp_min[w]);
        row_min[e] = mm;
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // 2. entries: normalised score and (id, position) key, engine-major
  for (int p = tid; p < P; p += blockDim.x) {
    if (p < M) {
      int e = 0;
      while (p >= a.base[e + 1]) ++e;
      const int j = p - a.base[e];
      F v = a.zero_scores[e] ? (F)0 : reinterpret_cast<const F*>(a.scores[e])[(size_t)row * a.width[e] + j];
      if (a.normalize) v = add_rn(sub_rn(v, row_min[e]), (F)a.offset);
      norm[p] = v;
      key[p] = (uint64_t)a.indices[e][(size_t)row * a.width[e] + j];  // negative ids sort last, -1 very last
      pos[p] = p;
      slot[p] = 0;
    } else {
      key[p] = ~0ull;
      pos[p] = 0x7fffffff;
    }
  }
  sort_pairs(key, pos, P);

  // 3. mark the first position of every run of a non-negative id
  for (int t = tid; t < M; t += blockDim.x) {
    const bool head = (t == 0) || (key[t] != key[t - 1]);
    if (head && (int64_t)key[t] >= 0) slot[pos[t]] = 1;
  }
  __syncthreads();

  // 4. exclusive prefix sum of the flags over positions -> output slots (first-occurrence order, merge.py:135-156)
  if (tid == 0) scan_carry = 0;
  __syncthreads();
  for (int base = 0; base < M; base += blockDim.x) {
    const int p = base + tid;
    const int f = (p < M) ? slot[p] : 0;
    int incl = f;
    for (int off = 1; off < 32; off <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, off);
      if ((tid & 31) >= off) incl += v;
    }
    if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
    __syncthreads();
    int warp_off = 0;
    for (int w = 0; w < (tid >> 5); ++w) warp_off += warp_sums[w];
    const int carry = scan_carry;
    if (p < M) slot[p] = f ? (carry + warp_off + incl - 1) : -1;
    __syncthreads();
    if (tid == blockDim.x - 1) scan_carry = carry + warp_off + incl;
    __syncthreads();
  }
  const int count = scan_carry;
  if (tid == 0) out_counts[row] = count;

  // 5. pad values: slots past `count` hold id -1 / score -inf; their raw score / label follow the reference's gather
  //    semantics (first entry of the engine whose id is exactly -1, else NaN / -1)
  F* o_s = out_scores + (size_t)row * a.out_width;
  int64_t* o_i = out_indices + (size_t)row * a.out_width;
  for (int c = count + tid; c < a.out_width; c += blockDim.x) {
    o_s[c] = -f_inf<F>();
    o_i[c] = -1;
    if (out_labels) out_labels[(size_t)row * a.out_width + c] = -1;
    for (int e = 0; e < a.n_engines; ++e) out_raw[((size_t)e * a.B + row) * a.out_width + c] = f_nan<F>();
  }
  __syncthreads();

  // 6. one thread per run: left-to-right sum of the weighted scores, first raw score per engine, label
  for (int t = tid; t < M; t += blockDim.x) {
    const bool head = (t == 0) || (key[t] != key[t - 1]);
    if (!head) continue;
    const int64_t id = (int64_t)key[t];
    if (id < 0 && id != -1) continue;
    F raw[kMaxEngines];
    bool have[kMaxEngines];
    for (int e = 0; e < a.n_engines; ++e) { raw[e] = f_nan<F>(); have[e] = false; }
    int64_t label = -1;
    bool have_label = false;
    F acc = 0;
    bool first = true;
    for (int u = t; u < M && key[u] == key[t]; ++u) {
      const int p = pos[u];
      int e = 0;
      while (p >= a.base[e + 1]) ++e;
      const F w = mul_rn(norm[p], (F)a.weight[e]);  // `result * weight`, then `score + scores[found]`
      acc = first ? w : add_rn(w, acc);
      first = false;
      if (!have[e]) { raw[e] = norm[p]; have[e] = true; }
      if (e == a.label_engine && !have_label) {
        label = a.labels[e][(size_t)row * a.width[e] + (p - a.base[e])];
        have_label = true;
      }
    }
    if (id >= 0) {
      const int c = slot[pos[t]];
      o_s[c] = acc;
      o_i[c] = id;
      if (out_labels) out_labels[(size_t)row * a.out_width + c] = label;
      for (int e = 0; e < a.n_engines; ++e) out_raw[((size_t)e * a.B + row) * a.out_width + c] = raw[e];
    } else {
      // the run of id == -1: supplies the gathered values of the padding slots
      for (int c = count; c < a.out_width; ++c) {
        if (out_labels && have_label) out_labels[(size_t)row * a.out_width + c] = label;
        for (int e = 0; e < a.n_engines; ++e)
          if (have[e]) out_raw[((size_t)e * a.B + row) * a.out_width + c] = raw[e];
      }
    }
  }
}

int pow2ceil_i(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

template <typename F>
int launch_typed(const MergeArgs& a, int M, void* out_scores, int64_t* out_indices, int64_t* out_labels, void* out_raw,
                 int* out_counts, cudaStream_t st) {
  const int P = pow2ceil_i(M > 1 ? M : 2);
  size_t smem = (size_t)P * 8 + (size_t)(M + (M & 1)) * sizeof(F) + (size_t)P * 4 + (size_t)M * 4 + 64;
  if (smem > 220 * 1024) {
    set_error("vodb_merge_results: %d entries per row need %zu bytes of shared memory (limit 220 KB)", M, smem);
    return VODB_EUNSUPPORTED;
  }
  if (smem > 48 * 1024) VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&merge_results_kernel<F>), 220 * 1024));
  merge_results_kernel<F><<<a.B, kThreads, smem, st>>>(a, M, P, reinterpret_cast<F*>(out_scores), out_indices,
                                                       out_labels, reinterpret_cast<F*>(out_raw), out_counts);
  VODB_CUDA_CHECK(cudaGetLastError());
  return VODB_OK;
}

}  // namespace

int launch_merge_results(int n_engines, const void* const* scores, const int64_t* const* indices,
                         const int64_t* const* labels, const int* widths, const double* weights, const int* zero_scores,
                         int B, int is_f64, int normalize, double offset, int label_engine, int out_width,
                         void* out_scores, int64_t* out_indices, int64_t* out_labels, void* out_raw, int* out_counts,
                         cudaStream_t st) {
  MergeArgs a{};
  a.n_engines = n_engines;
  a.B = B;
  a.out_width = out_width;
  a.normalize = normalize;
  a.label_engine = label_engine;
  a.offset = offset;
  int M = 0;
  for (int e = 0; e < n_engines; ++e) {
    a.width[e] = widths[e];
    a.base[e] = M;
    M += widths[e];
    a.zero_scores[e] = zero_scores ? zero_scores[e] : 0;
    a.weight[e] = weights[e];
    a.scores[e] = scores[e];
    a.indices[e] = indices[e];
    a.labels[e] = labels ? labels[e] : nullptr;
  }
  for (int e = n_engines; e <= kMaxEngines; ++e) a.base[e] = M;
  if (B == 0) return VODB_OK;
  return is_f64 ? launch_typed<double>(a, M, out_scores, out_indices, out_labels, out_raw, out_counts, st)
                : launch_typed<float>(a, M, out_scores, out_indices, out_labels, out_raw, out_counts, st);
}

}  // namespace vodb
