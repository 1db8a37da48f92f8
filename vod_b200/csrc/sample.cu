// sample.cu — fused labeled priority sampling (softmax / log-sum-exp, Gumbel-style priority keys,
// top-(k+1) selection, importance log-weights, per-label self-normalisation).
//
// Replaces the numba path `_labeled_priority_sampling_2d_` -> `_labeled_priority_sampling_1d_` ->
// `_priority_sampling_1d` (reference src/vod_dataloaders/core/sample.py:323-352, :245-320, :160-219) and the
// helpers it calls (`log_softmax_1d_`, `max_1d`, `_logsumexp_1d`: numpy_ops.py:162-216). One CTA per
// query row; everything lives in shared memory; the only HBM traffic is the [K] score row in and the
// [k_total] picks out. The operation order is the specification shared with oracle/sample_twin.c, which
// makes results bit-identical to the CPU twin:
//   * exp/log/log1p: vodb_math.h (IEEE-only, no FMA contraction);
//   * sums: thread t accumulates i = t, t+256, ... in increasing i, then a pairwise tree over 256 lanes;
//   * order: bitonic sort of (group, key desc [NaN last], index asc) packed in 64 bits (a total order).
// Noise: Exp(1) from Philox-4x32-10(seed, offset, row, col) unless an explicit noise matrix is passed.
#include "common.cuh"

namespace vodb {

namespace {

constexpr int NT = 256;

__device__ __forceinline__ uint64_t sort_key(int group, float v, uint32_t i) {
  return ((uint64_t)(group & 1) << 63) | ((uint64_t)(~ord_u32(v)) << 31) | (uint64_t)i;
}

// ascending bitonic sort of skey[0..P), P power of two
__device__ void block_sort_u64(uint64_t* skey, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (P >> 1); t += NT) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = (lo & size) == 0;
        uint64_t a = skey[lo], b = skey[hi];
        if ((a > b) == up) {
          skey[lo] = b;
          skey[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// spec-ordered sums of f[i] per group (sel==nullptr: single group 0). Results in out[0..1], all threads.
__device__ void block_sum2(const float* f, const uint8_t* sel, int n, float* red /*[2*NT]*/, float out[2]) {
  float a0 = 0.0f, a1 = 0.0f;
  for (int i = threadIdx.x; i < n; i += NT) {
    int g = sel ? sel[i] : 0;
    if (g == 0) a0 = VM_ADD(a0, f[i]);
    else a1 = VM_ADD(a1, f[i]);
  }
  red[threadIdx.x] = a0;
  red[NT + threadIdx.x] = a1;
  __syncthreads();
  for (int off = NT / 2; off >= 1; off >>= 1) {
    if ((int)threadIdx.x < off) {
      red[threadIdx.x] = VM_ADD(red[threadIdx.x], red[threadIdx.x + off]);
      red[NT + threadIdx.x] = VM_ADD(red[NT + threadIdx.x], red[NT + threadIdx.x + off]);
    }
    __syncthreads();
  }
  out[0] = red[0];
  out[1] = red[NT];
  __syncthreads();
}

struct SampleSmem {
  int m[2];
  int n_neg_finite;
  uint32_t mx_ord[2];
  float thr[2];
  int active[2];
};

__global__ void __launch_bounds__(NT)
sample_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ labels, const float* __restrict__ noise,
              int K, int P, int k_positive, int k_total, int normalized, float temperature, int max_support,
              int quirks, uint64_t seed, uint64_t offset, int64_t* __restrict__ out_ids,
              float* __restrict__ out_logw, uint8_t* __restrict__ out_labels, float* __restrict__ out_lse) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ SampleSmem sm;
  __shared__ float red[2 * NT];
  uint64_t* skey = reinterpret_cast<uint64_t*>(dyn);                 // [P]
  float* lp = reinterpret_cast<float*>(dyn + (size_t)P * 8);         // [K]
  float* ex = lp + K;                                                // [K] scratch: exp values, then keys, then weights
  uint8_t* grp = reinterpret_cast<uint8_t*>(ex + K);                 // [K]

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* s = scores + (size_t)b * K;
  const uint8_t* lab = labels ? labels + (size_t)b * K : nullptr;
  int64_t* o_ids = out_ids + (size_t)b * k_total;
  float* o_w = out_logw + (size_t)b * k_total;
  uint8_t* o_lab = out_labels + (size_t)b * k_total;
  const float NINF = vm_ninf();

  for (int j = tid; j < k_total; j += NT) {
    o_ids[j] = -1;
    o_w[j] = NINF;
    o_lab[j] = 0;
  }
  if (tid == 0) {
    sm.m[0] = sm.m[1] = 0;
    sm.n_neg_finite = 0;
    sm.mx_ord[0] = sm.mx_ord[1] = ord_u32(NINF);
    sm.active[0] = sm.active[1] = 0;
  }
  __syncthreads();

  // group membership and counts (sample.py:258-264)
  const float tinv = temperature > 0.0f ? temperature : 1.0f;
  {
    int c0 = 0, c1 = 0, cf = 0;
    for (int i = tid; i < K; i += NT) {
      float v = s[i];
      int g = (lab != nullptr && lab[i] > 0) ? 0 : 1;
      grp[i] = (uint8_t)g;
      if (g == 0) c0++;
      else {
        c1++;
        if (!vm_isinf(v)) cf++;
      }
      float l = VM_MUL(v, tinv);
      lp[i] = vm_isnan(l) ? NINF : l;
    }
    if (c0) atomicAdd(&sm.m[0], c0);
    if (c1) atomicAdd(&sm.m[1], c1);
    if (cf) atomicAdd(&sm.n_neg_finite, cf);
  }
  __syncthreads();
  const int m0 = sm.m[0], m1 = sm.m[1];
  const int kt = k_total < K ? k_total : K;
  int kp = k_positive;
  if (sm.n_neg_finite < kt - kp) kp = kt - sm.n_neg_finite;

  // truncation of the support (sample.py:176-178)
  if (max_support > 0 && (m0 > max_support || m1 > max_support)) {
    for (int i = tid; i < P; i += NT) skey[i] = (i < K) ? sort_key(grp[i], lp[i], (uint32_t)i) : ~0ull;
    block_sort_u64(skey, P);
    if (tid < 2) {
      int g = tid;
      int mg = g == 0 ? m0 : m1;
      int start = g == 0 ? 0 : m0;
      sm.active[g] = mg > max_support;
      if (sm.active[g]) sm.thr[g] = lp[(uint32_t)(skey[start + max_support - 1] & 0x7fffffffu)];
    }
    __syncthreads();
    for (int i = tid; i < K; i += NT) {
      int g = grp[i];
      if (!sm.active[g]) continue;
      bool mask = (quirks & 1) ? (lp[i] >= sm.thr[g]) : (lp[i] < sm.thr[g]);
      if (mask) lp[i] = NINF;
    }
    __syncthreads();
  }

  // log-softmax per group (numpy_ops.py:207-216) + log normaliser (sample.py:184)
  {
    uint32_t mo0 = ord_u32(NINF), mo1 = mo0;
    for (int i = tid; i < K; i += NT) {
      uint32_t o = ord_u32(lp[i]);
      if (grp[i] == 0) mo0 = max(mo0, o);
      else mo1 = max(mo1, o);
    }
    atomicMax(&sm.mx_ord[0], mo0);
    atomicMax(&sm.mx_ord[1], mo1);
  }
  __syncthreads();
  float mx[2];
  for (int g = 0; g < 2; ++g) {
    mx[g] = ord_to_float(sm.mx_ord[g]);
    if (vm_f2u(mx[g]) == 0xff800000u) mx[g] = 0.0f;
  }
  for (int i = tid; i < K; i += NT) {
    float v = VM_SUB(lp[i], mx[grp[i]]);
    lp[i] = v;
    ex[i] = vodb_expf(v);
  }
  __syncthreads();
  float sums[2];
  block_sum2(ex, grp, K, red, sums);
  float lse[2] = {vodb_logf(sums[0]), vodb_logf(sums[1])};
  for (int i = tid; i < K; i += NT) {
    float v = VM_SUB(lp[i], lse[grp[i]]);
    lp[i] = v;
    ex[i] = vodb_expf(v);
  }
  __syncthreads();
  block_sum2(ex, grp, K, red, sums);
  if (tid == 0) {
    out_lse[(size_t)b * 2 + 0] = vm_canon_nan(vodb_logf(sums[0]));
    out_lse[(size_t)b * 2 + 1] = vm_canon_nan(vodb_logf(sums[1]));
  }

  // priority keys (sample.py:187-193) and per-group descending order (sample.py:196)
  float* key = ex;
  for (int i = tid; i < K; i += NT) {
    float kv = lp[i];
    if (temperature > 0.0f) {
      float e = noise ? noise[(size_t)b * K + i] : vodb_exp1_noise(seed, offset, (uint32_t)b, (uint32_t)i);
      kv = VM_SUB(kv, vodb_logf(e));
    }
    key[i] = kv;
  }
  __syncthreads();
  for (int i = tid; i < P; i += NT) skey[i] = (i < K) ? sort_key(grp[i], key[i], (uint32_t)i) : ~0ull;
  block_sort_u64(skey, P);

  // picks + importance weights (sample.py:199-216), positives first then negatives (sample.py:310-320)
  int written = 0;
  for (int g = 0; g < 2; ++g) {
    const int mg = g == 0 ? m0 : m1;
    const int start = g == 0 ? 0 : m0;
    int kg = g == 0 ? kp : kt - written;
    if (kg < 0) kg = 0;
    const int n_pick = kg < mg ? kg : mg;
    float log_tau = NINF;
    if (kg < mg) log_tau = key[(uint32_t)(skey[start + kg] & 0x7fffffffu)];
    float* w = o_w + written;
    for (int j = tid; j < n_pick; j += NT) {
      uint32_t i = (uint32_t)(skey[start + j] & 0x7fffffffu);
      float log_pi = lp[i];
      float lw;
      if (log_tau > NINF) {
        float d = VM_SUB(log_pi, log_tau);
        float q = vodb_log1pf(VM_SUB(0.0f, vodb_expf(VM_SUB(0.0f, vodb_expf(d)))));
        lw = VM_SUB(log_pi, q);
      } else {
        lw = log_pi;
      }
      w[j] = vm_canon_nan(lw);
      o_ids[written + j] = (int64_t)i;
      o_lab[written + j] = (uint8_t)(g == 0);
    }
    __syncthreads();
    if (normalized && n_pick > 0) {
      // log_softmax over the picked weights (sample.py:289-290, 301-302), same order as the twin;
      // the exp scratch [k_total] sits behind grp[] in dynamic shared memory.
      float* scratch = reinterpret_cast<float*>(grp + ((K + 15) / 16) * 16);  // [k_total] (allocated by the host)
      if (tid == 0) sm.mx_ord[0] = ord_u32(NINF);
      __syncthreads();
      uint32_t mo = ord_u32(NINF);
      for (int j = tid; j < n_pick; j += NT) {
        float v = w[j];
        if (vm_isnan(v)) { v = NINF; w[j] = v; }
        mo = max(mo, ord_u32(v));
      }
      atomicMax(&sm.mx_ord[0], mo);
      __syncthreads();
      float wmx = ord_to_float(sm.mx_ord[0]);
      if (vm_f2u(wmx) == 0xff800000u) wmx = 0.0f;
      for (int j = tid; j < n_pick; j += NT) {
        float v = VM_SUB(w[j], wmx);
        w[j] = v;
        scratch[j] = vodb_expf(v);
      }
      __syncthreads();
      float ws[2];
      block_sum2(scratch, nullptr, n_pick, red, ws);
      float wl = vodb_logf(ws[0]);
      for (int j = tid; j < n_pick; j += NT) w[j] = vm_canon_nan(VM_SUB(w[j], wl));
      __syncthreads();
    }
    written += n_pick;
  }
}

__host__ __device__ inline int pow2ceil_i(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

int launch_sample(const float* scores, const uint8_t* labels, const float* noise, int B, int K, int k_positive,
                  int k_total, int normalized, float temperature, int max_support, int quirks, uint64_t seed,
                  uint64_t offset, int64_t* out_ids, float* out_logw, uint8_t* out_labels, float* out_lse,
                  cudaStream_t stream) {
  if (B == 0) return VODB_OK;
  int P = pow2ceil_i(K > 1 ? K : 2);
  size_t smem = (size_t)P * 8 + (size_t)K * 8 + (size_t)((K + 15) / 16) * 16 + (size_t)k_total * 4 + 16;
  if (smem > 220 * 1024) {
    set_error("vodb_sample: K=%d needs %zu bytes of shared memory (limit 220 KB)", K, smem);
    return VODB_EUNSUPPORTED;
  }
  VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&sample_kernel), smem));
  sample_kernel<<<B, NT, smem, stream>>>(scores, labels, noise, K, P, k_positive, k_total, normalized, temperature,
                                         max_support, quirks, seed, offset, out_ids, out_logw, out_labels, out_lse);
  VODB_CUDA_CHECK(cudaGetLastError());
  return VODB_OK;
}

}  // namespace vodb
