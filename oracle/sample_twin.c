/* oracle/sample_twin.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Single-threaded CPU twin of the fused sampling kernel (vod_b200/csrc/sample.cuh).
 * It restates the reference's labeled priority sampling
 *     src/vod_dataloaders/core/sample.py:245-320  (_labeled_priority_sampling_1d_)
 *     src/vod_dataloaders/core/sample.py:160-219  (_priority_sampling_1d)
 *     src/vod_dataloaders/core/numpy_ops.py:162-216 (max_1d, _logsumexp_1d, log_softmax_1d_)
 * with two things pinned that the reference leaves open, so that the GPU result can
 * be compared bit for bit:
 *   (1) exp/log/log1p are the shared IEEE-only kernels of vodb_math.h (the reference
 *       uses numba fastmath libm calls, reproducible only to ~1e-6);
 *   (2) sums run in the fixed "256 strided lanes + pairwise tree" order the CUDA
 *       kernel uses (the reference sums sequentially); argsort ties (only possible
 *       among equal keys) break by lower index, NaN keys sort last.
 * Parity pinning: this twin is checked against the reference's own numba code
 * (loaded by oracle/ref_shim.py) in tests/test_sampling_oracle.py and against the
 * golden vectors under tests/golden/ that were generated from that code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../vod_b200/csrc/vodb_math.h"

#define NT 256 /* virtual lanes of the reduction order */

/* exported thin wrappers so that tests can probe the shared math directly */
float twin_logf(float x) { return vodb_logf(x); }
float twin_expf(float x) { return vodb_expf(x); }
float twin_log1pf(float x) { return vodb_log1pf(x); }
float twin_exp1_noise(uint64_t seed, uint64_t offset, uint32_t row, uint32_t col) {
  return vodb_exp1_noise(seed, offset, row, col);
}
float twin_synth_value(uint64_t seed, uint64_t row, uint32_t col) {
  return vodb_synth_value(seed, row, col);
}
uint16_t twin_f32_to_bf16(float x) { return vodb_f32_to_bf16(x); }
uint16_t twin_f32_to_f16(float x) { return vodb_f32_to_f16(x); }
float twin_bf16_to_f32(uint16_t h) { return vodb_bf16_to_f32(h); }
float twin_f16_to_f32(uint16_t h) { return vodb_f16_to_f32(h); }

void twin_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                 uint32_t out[4]) {
  vm_u32x4 r = vodb_philox4x32(c0, c1, c2, c3, k0, k1);
  memcpy(out, r.v, 16);
}

/* order-preserving map float -> uint32 (larger float = larger uint), NaN lowest, -0 == +0 */
static uint32_t ord_u32(float x) {
  if (vm_isnan(x)) return 0u;
  uint32_t u = vm_f2u(x);
  if ((u & 0x7fffffffu) == 0u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

static float ord_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return vm_u2f(u);
}

/* composite sort key: group asc, value desc, index asc */
static uint64_t sort_key(int group, float v, uint32_t i) {
  return ((uint64_t)(group & 1) << 63) | ((uint64_t)(~ord_u32(v)) << 31) | (uint64_t)i;
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}

/* sum over i in [0,n) of f[i] for which sel[i]==g (sel==NULL: all), in kernel order */
static float spec_sum(const float* f, const uint8_t* sel, int g, int n) {
  float p[NT];
  for (int t = 0; t < NT; ++t) {
    float acc = 0.0f;
    for (int i = t; i < n; i += NT)
      if (sel == NULL || sel[i] == g) acc = VM_ADD(acc, f[i]);
    p[t] = acc;
  }
  for (int off = NT / 2; off >= 1; off >>= 1)
    for (int t = 0; t < off; ++t) p[t] = VM_ADD(p[t], p[t + off]);
  return p[0];
}

/* in-place log-softmax of w[0..n) in kernel order (numpy_ops.py:207-216) */
static void spec_log_softmax(float* w, int n, float* scratch) {
  uint32_t mo = ord_u32(vm_ninf());
  for (int i = 0; i < n; ++i) {
    if (vm_isnan(w[i])) w[i] = vm_ninf();
    if (ord_u32(w[i]) > mo) mo = ord_u32(w[i]);
  }
  float mx = ord_to_float(mo);                 /* max through the ordered image: -0 and +0 are one value */
  if (vm_f2u(mx) == 0xff800000u) mx = 0.0f; /* max_1d: -inf -> 0 */
  for (int i = 0; i < n; ++i) {
    w[i] = VM_SUB(w[i], mx);
    scratch[i] = vodb_expf(w[i]);
  }
  float lse = vodb_logf(spec_sum(scratch, NULL, 0, n));
  for (int i = 0; i < n; ++i) w[i] = vm_canon_nan(VM_SUB(w[i], lse));
}

static void sample_row(const float* s, const uint8_t* lab, const float* noise, int K, int k_positive,
                       int k_total, int normalized, float temperature, int max_support, int quirks,
                       int64_t* out_ids, float* out_logw, uint8_t* out_labels, float* out_lse,
                       float* lp, float* ex, float* key, uint8_t* grp, uint64_t* skey) {
  int m[2] = {0, 0};
  int n_neg_finite = 0;
  for (int i = 0; i < K; ++i) {
    grp[i] = (lab != NULL && lab[i] > 0) ? 0 : 1; /* 0 = positive group, 1 = negative group */
    m[grp[i]]++;
    if (grp[i] == 1 && !vm_isinf(s[i])) n_neg_finite++; /* np.isinf: NaN counts as finite */
  }
  int kt = k_total < K ? k_total : K;                  /* sample.py:267 */
  int kp = k_positive;
  if (n_neg_finite < kt - kp) kp = kt - n_neg_finite;  /* sample.py:277-278 */

  float tinv = temperature > 0.0f ? temperature : 1.0f; /* sample.py:170 */
  for (int i = 0; i < K; ++i) {
    float v = VM_MUL(s[i], tinv);
    lp[i] = vm_isnan(v) ? vm_ninf() : v;
  }

  /* truncation, sample.py:176-178 */
  if (max_support > 0) {
    for (int i = 0; i < K; ++i) skey[i] = sort_key(grp[i], lp[i], (uint32_t)i);
    qsort(skey, (size_t)K, sizeof(uint64_t), cmp_u64);
    int start[2] = {0, m[0]};
    float thr[2];
    int active[2];
    for (int g = 0; g < 2; ++g) {
      active[g] = m[g] > max_support;
      if (active[g]) thr[g] = lp[(uint32_t)(skey[start[g] + max_support - 1] & 0x7fffffffu)];
    }
    for (int i = 0; i < K; ++i) {
      int g = grp[i];
      if (!active[g]) continue;
      int mask = (quirks & 1) ? (lp[i] >= thr[g]) : (lp[i] < thr[g]);
      if (mask) lp[i] = vm_ninf();
    }
  }

  /* log-softmax per group + log-normaliser (sample.py:180-184) */
  uint32_t mo[2] = {ord_u32(vm_ninf()), ord_u32(vm_ninf())};
  for (int i = 0; i < K; ++i)
    if (ord_u32(lp[i]) > mo[grp[i]]) mo[grp[i]] = ord_u32(lp[i]);
  float mx[2];
  for (int g = 0; g < 2; ++g) {
    mx[g] = ord_to_float(mo[g]);
    if (vm_f2u(mx[g]) == 0xff800000u) mx[g] = 0.0f;
  }
  for (int i = 0; i < K; ++i) {
    lp[i] = VM_SUB(lp[i], mx[grp[i]]);
    ex[i] = vodb_expf(lp[i]);
  }
  float lse[2];
  for (int g = 0; g < 2; ++g) lse[g] = vodb_logf(spec_sum(ex, grp, g, K));
  for (int i = 0; i < K; ++i) {
    lp[i] = VM_SUB(lp[i], lse[grp[i]]);
    ex[i] = vodb_expf(lp[i]);
  }
  for (int g = 0; g < 2; ++g) out_lse[g] = vm_canon_nan(vodb_logf(spec_sum(ex, grp, g, K)));

  /* keys (sample.py:187-193) and per-group descending order (sample.py:196) */
  for (int i = 0; i < K; ++i) {
    key[i] = temperature > 0.0f ? VM_SUB(lp[i], vodb_logf(noise[i])) : lp[i];
    skey[i] = sort_key(grp[i], key[i], (uint32_t)i);
  }
  qsort(skey, (size_t)K, sizeof(uint64_t), cmp_u64);

  int start[2] = {0, m[0]};
  int ksel[2];
  int written = 0;
  for (int g = 0; g < 2; ++g) {
    int kg = (g == 0) ? kp : kt - written; /* sample.py:296 */
    if (kg < 0) kg = 0;
    int n_pick = kg < m[g] ? kg : m[g];
    float log_tau = vm_ninf();
    if (kg < m[g]) log_tau = key[(uint32_t)(skey[start[g] + kg] & 0x7fffffffu)]; /* :199-203 */
    float* w = out_logw + written;
    for (int j = 0; j < n_pick; ++j) {
      uint32_t i = (uint32_t)(skey[start[g] + j] & 0x7fffffffu);
      float log_pi = lp[i];
      float lw;
      if (log_tau > vm_ninf()) { /* :210-216 */
        float d = VM_SUB(log_pi, log_tau);
        float q = vodb_log1pf(VM_SUB(0.0f, vodb_expf(VM_SUB(0.0f, vodb_expf(d)))));
        lw = VM_SUB(log_pi, q);
      } else {
        lw = log_pi;
      }
      w[j] = vm_canon_nan(lw);
      out_ids[written + j] = (int64_t)i;
      out_labels[written + j] = (uint8_t)(g == 0);
    }
    if (normalized && n_pick > 0) spec_log_softmax(w, n_pick, ex); /* :289-290, :301-302 */
    ksel[g] = n_pick;
    written += n_pick;
  }
  (void)ksel;
}

/* Same contract as vodb_sample (include/vodb.h), host pointers only.
 * Returns 0, or -1 on bad arguments / allocation failure. */
int twin_sample(const float* scores, const uint8_t* labels, const float* noise, int B, int K,
                int k_positive, int k_total, int normalized, float temperature, int max_support,
                int quirks, uint64_t seed, uint64_t offset, int64_t* out_ids, float* out_logw,
                uint8_t* out_labels, float* out_lse) {
  if (B < 0 || K < 0 || k_total < 0 || k_positive < 0 || k_positive > k_total) return -1;
  for (int64_t i = 0; i < (int64_t)B * k_total; ++i) {
    out_ids[i] = -1;
    out_logw[i] = vm_ninf();
    out_labels[i] = 0;
  }
  if (B == 0) return 0;
  size_t kk = (size_t)(K > 0 ? K : 1);
  float* lp = (float*)malloc(kk * sizeof(float));
  float* ex = (float*)malloc(kk * sizeof(float));
  float* key = (float*)malloc(kk * sizeof(float));
  float* nz = (float*)malloc(kk * sizeof(float));
  uint8_t* grp = (uint8_t*)malloc(kk);
  uint64_t* skey = (uint64_t*)malloc(kk * sizeof(uint64_t));
  if (!lp || !ex || !key || !nz || !grp || !skey) return -1;
  for (int b = 0; b < B; ++b) {
    const float* nrow;
    if (noise != NULL) {
      nrow = noise + (size_t)b * K;
    } else {
      for (int i = 0; i < K; ++i) nz[i] = vodb_exp1_noise(seed, offset, (uint32_t)b, (uint32_t)i);
      nrow = nz;
    }
    sample_row(scores + (size_t)b * K, labels ? labels + (size_t)b * K : NULL, nrow, K, k_positive,
               k_total, normalized, temperature, max_support, quirks, out_ids + (size_t)b * k_total,
               out_logw + (size_t)b * k_total, out_labels + (size_t)b * k_total, out_lse + (size_t)b * 2,
               lp, ex, key, grp, skey);
  }
  free(lp); free(ex); free(key); free(nz); free(grp); free(skey);
  return 0;
}

/* Synthetic corpus rows [row0,row0+n) x dim as float32 already rounded to `dtype`
 * (0=f32,1=bf16,2=f16) — the CPU side of vodb_store_fill_synthetic. */
void twin_synth_rows(uint64_t seed, int64_t row0, int64_t n, int dim, int dtype, int unit_norm,
                     float* out) {
  for (int64_t r = 0; r < n; ++r) {
    float* o = out + (size_t)r * dim;
    for (int c = 0; c < dim; ++c) o[c] = vodb_synth_value(seed, (uint64_t)(row0 + r), (uint32_t)c);
    if (unit_norm) {
      /* sequential fp32 sum of squares, then one division per element (matches the kernel) */
      float ss = 0.0f;
      for (int c = 0; c < dim; ++c) ss = VM_ADD(ss, VM_MUL(o[c], o[c]));
      float nrm = sqrtf(ss); /* IEEE correctly rounded on both sides */
      if (nrm > 0.0f)
        for (int c = 0; c < dim; ++c) o[c] = VM_DIV(o[c], nrm);
    }
    for (int c = 0; c < dim; ++c) {
      if (dtype == 1) o[c] = vodb_bf16_to_f32(vodb_f32_to_bf16(o[c]));
      else if (dtype == 2) o[c] = vodb_f16_to_f32(vodb_f32_to_f16(o[c]));
    }
  }
}
