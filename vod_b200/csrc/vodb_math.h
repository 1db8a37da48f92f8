/* vodb_math.h — bit-reproducible float32 math shared by the CUDA kernels and the
 * CPU twin (oracle/sample_twin.c).
 *
 * The sampling step that follows retrieval (reference:
 * src/vod_dataloaders/core/sample.py:160-219, numpy_ops.py:198-216) is made of
 * exp / log / log1p calls. libm (host) and libdevice (GPU) round those
 * differently, so "sampled indices and weights bit-identical to a CPU
 * reimplementation" is only reachable when both sides evaluate the SAME sequence
 * of IEEE-754 operations. Everything here is built from +, -, *, / (all
 * correctly rounded on both sides), integer bit manipulation and nothing else;
 * no operation may be contracted into an FMA, hence the explicit _rn intrinsics
 * on the device and -ffp-contract=off for the host twin.
 *
 * Algorithms: the classic fdlibm/musl float kernels (argument reduction +
 * minimax polynomial), re-expressed with explicit operation order.
 * Accuracy: < 1 ulp for expf/logf on their whole domain; log1pf uses Kahan's
 * log(1+x)*x/((1+x)-1) form (a few ulp).
 */
#ifndef VODB_MATH_H_
#define VODB_MATH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define VM_FN __host__ __device__ __forceinline__
#else
#define VM_FN static inline
#include <string.h>
#endif

#if defined(__CUDA_ARCH__)
#define VM_MUL(a, b) __fmul_rn((a), (b))
#define VM_ADD(a, b) __fadd_rn((a), (b))
#define VM_SUB(a, b) __fsub_rn((a), (b))
#define VM_DIV(a, b) __fdiv_rn((a), (b))
#else
/* host: plain IEEE ops; the translation unit must be built with -ffp-contract=off */
#define VM_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define VM_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define VM_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define VM_DIV(a, b) ((float)((float)(a) / (float)(b)))
#endif

VM_FN uint32_t vm_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
#endif
}

VM_FN float vm_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x;
  memcpy(&x, &u, 4);
  return x;
#endif
}

#define VM_INF_BITS 0x7f800000u
#define VM_NAN_BITS 0x7fc00000u

VM_FN float vm_inf(void) { return vm_u2f(VM_INF_BITS); }
VM_FN float vm_ninf(void) { return vm_u2f(0xff800000u); }
VM_FN float vm_nan(void) { return vm_u2f(VM_NAN_BITS); }
VM_FN int vm_isnan(float x) { return (vm_f2u(x) & 0x7fffffffu) > VM_INF_BITS; }
VM_FN int vm_isinf(float x) { return (vm_f2u(x) & 0x7fffffffu) == VM_INF_BITS; }
/* NaN sign/payload differs between x86 (default NaN 0xffc00000) and the GPU (0x7fffffff): results that may be
 * NaN are written through this so that CPU and GPU outputs compare bit for bit. */
VM_FN float vm_canon_nan(float x) { return vm_isnan(x) ? vm_u2f(VM_NAN_BITS) : x; }

/* natural logarithm */
VM_FN float vodb_logf(float x) {
  const float ln2_hi = 6.9313812256e-01f; /* 0x3f317180 */
  const float ln2_lo = 9.0580006145e-06f; /* 0x3717f7d1 */
  const float Lg1 = 0.66666662693f;       /* 0xaaaaaa.0p-24 */
  const float Lg2 = 0.40000972152f;       /* 0xccce13.0p-25 */
  const float Lg3 = 0.28498786688f;       /* 0x91e9ee.0p-25 */
  const float Lg4 = 0.24279078841f;       /* 0xf89e26.0p-26 */
  uint32_t ix = vm_f2u(x);
  int k = 0;
  if ((ix & 0x7fffffffu) == 0u) return vm_ninf(); /* log(+-0) = -inf */
  if (ix > VM_INF_BITS && ix < 0x80000000u) return x; /* +NaN */
  if (ix >> 31) return vm_nan();                  /* x < 0 or -NaN */
  if (ix == VM_INF_BITS) return x;                /* +inf */
  if (ix < 0x00800000u) {                         /* subnormal: scale up by 2^25 */
    x = VM_MUL(x, 33554432.0f);
    ix = vm_f2u(x);
    k -= 25;
  }
  /* reduce x into [sqrt(2)/2, sqrt(2)) */
  ix += 0x3f800000u - 0x3f3504f3u;
  k += (int)(ix >> 23) - 127;
  ix = (ix & 0x007fffffu) + 0x3f3504f3u;
  x = vm_u2f(ix);

  float f = VM_SUB(x, 1.0f);
  float s = VM_DIV(f, VM_ADD(2.0f, f));
  float z = VM_MUL(s, s);
  float w = VM_MUL(z, z);
  float t1 = VM_MUL(w, VM_ADD(Lg2, VM_MUL(w, Lg4)));
  float t2 = VM_MUL(z, VM_ADD(Lg1, VM_MUL(w, Lg3)));
  float R = VM_ADD(t2, t1);
  float hfsq = VM_MUL(VM_MUL(0.5f, f), f);
  float dk = (float)k;
  /* dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f) */
  float a = VM_ADD(VM_MUL(s, VM_ADD(hfsq, R)), VM_MUL(dk, ln2_lo));
  float b = VM_SUB(VM_SUB(hfsq, a), f);
  return VM_SUB(VM_MUL(dk, ln2_hi), b);
}

/* y * 2^k with a fixed, shared sequence of roundings (k in [-160, 128]) */
VM_FN float vm_scale2(float y, int k) {
  if (k > 127) { /* k == 128 */
    y = VM_MUL(y, 2.0f);
    k -= 1;
  }
  if (k >= -125) return VM_MUL(y, vm_u2f((uint32_t)(k + 127) << 23));
  /* result may be subnormal: scale in two exact-constant steps */
  y = VM_MUL(y, vm_u2f((uint32_t)(k + 100 + 127) << 23));
  return VM_MUL(y, 7.8886090522e-31f); /* 2^-100 */
}

/* e^x */
VM_FN float vodb_expf(float x) {
  const float ln2hi = 6.9314575195e-1f;  /* 0x3f317200 */
  const float ln2lo = 1.4286067653e-6f;  /* 0x35bfbe8e */
  const float invln2 = 1.4426950216e+0f; /* 0x3fb8aa3b */
  const float P1 = 1.6666625440e-1f;     /*  0xaaaa8f.0p-26 */
  const float P2 = -2.7667332906e-3f;    /* -0xb55215.0p-32 */
  uint32_t hx = vm_f2u(x);
  int sign = (int)(hx >> 31);
  float hi, lo;
  int k;
  hx &= 0x7fffffffu;
  if (hx >= 0x42aeac50u) { /* |x| >= 87.33655 or NaN */
    if (hx > VM_INF_BITS) return x; /* NaN */
    if (hx >= 0x42b17218u && !sign) return vm_inf(); /* x >= 88.722839: overflow */
    if (sign && hx >= 0x42cff1b5u) return 0.0f;      /* x <= -103.972084: underflow */
  }
  if (hx > 0x3eb17218u) { /* |x| > 0.5 ln2 */
    if (hx > 0x3f851592u) { /* |x| > 1.5 ln2 */
      float t = VM_ADD(VM_MUL(invln2, x), sign ? -0.5f : 0.5f);
      k = (int)t; /* truncation toward zero: exact, same on both sides */
    } else {
      k = 1 - sign - sign;
    }
    float fk = (float)k;
    hi = VM_SUB(x, VM_MUL(fk, ln2hi)); /* fk*ln2hi is exact */
    lo = VM_MUL(fk, ln2lo);
    x = VM_SUB(hi, lo);
  } else if (hx > 0x39000000u) { /* |x| > 2^-14 */
    k = 0;
    hi = x;
    lo = 0.0f;
  } else {
    return VM_ADD(1.0f, x);
  }
  float xx = VM_MUL(x, x);
  float c = VM_SUB(x, VM_MUL(xx, VM_ADD(P1, VM_MUL(xx, P2))));
  /* y = 1 + (x*c/(2-c) - lo + hi) */
  float q = VM_DIV(VM_MUL(x, c), VM_SUB(2.0f, c));
  float y = VM_ADD(1.0f, VM_ADD(VM_SUB(q, lo), hi));
  if (k == 0) return y;
  return vm_scale2(y, k);
}

/* log(1+x), Kahan's form */
VM_FN float vodb_log1pf(float x) {
  if (vm_isnan(x)) return x;
  if (vm_f2u(x) == VM_INF_BITS) return x; /* +inf */
  float u = VM_ADD(1.0f, x);
  if (u == 1.0f) return x;
  float l = vodb_logf(u); /* u<0 -> NaN, u==0 -> -inf */
  if (vm_isnan(l) || vm_isinf(l)) return l;
  return VM_MUL(l, VM_DIV(x, VM_SUB(u, 1.0f)));
}

/* ------------------------------------------------------------------------- */
/* Philox-4x32-10 counter-based generator (Salmon et al., SC'11).             */

typedef struct {
  uint32_t v[4];
} vm_u32x4;

VM_FN vm_u32x4 vodb_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                               uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  vm_u32x4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

/* uniform in (0,1): 24 random bits, centred so that 0 and 1 are excluded */
VM_FN float vodb_u01(uint32_t r) {
  return VM_MUL(VM_ADD((float)(r >> 8), 0.5f), 5.9604644775390625e-8f); /* 2^-24 */
}

/* Exp(1) noise for element (row, col) of a [B,K] score matrix: the counter-based
 * stand-in for np.random.exponential(size=[B,K]) (reference sample.py:398). */
VM_FN float vodb_exp1_noise(uint64_t seed, uint64_t offset, uint32_t row, uint32_t col) {
  vm_u32x4 r = vodb_philox4x32(col >> 2, row, (uint32_t)offset, (uint32_t)(offset >> 32),
                               (uint32_t)seed, (uint32_t)(seed >> 32));
  float u = vodb_u01(r.v[col & 3u]);
  return VM_SUB(0.0f, vodb_logf(u));
}

/* Synthetic embedding value for element (row, col): sum of four uniform bytes
 * (Irwin-Hall, n=4) centred and scaled to unit variance. Pure integer work plus
 * one exact int->float conversion and one multiplication, so the CPU and the
 * GPU produce the same float32 before rounding to the store dtype. */
VM_FN float vodb_synth_from_word(uint32_t w) {
  int s = (int)(w & 0xffu) + (int)((w >> 8) & 0xffu) + (int)((w >> 16) & 0xffu) + (int)(w >> 24);
  return VM_MUL((float)(s - 510), 6.7658743e-3f); /* 1/sqrt(4*(256^2-1)/12) */
}

VM_FN float vodb_synth_value(uint64_t seed, uint64_t row, uint32_t col) {
  vm_u32x4 r = vodb_philox4x32(col >> 2, (uint32_t)row, (uint32_t)(row >> 32), 0x53594e54u /* "SYNT" */,
                               (uint32_t)seed, (uint32_t)(seed >> 32));
  return vodb_synth_from_word(r.v[col & 3u]);
}

/* float32 -> bf16 / fp16 bit patterns, round-to-nearest-even, shared definition */
VM_FN uint16_t vodb_f32_to_bf16(float x) {
  uint32_t u = vm_f2u(x);
  if ((u & 0x7fffffffu) > VM_INF_BITS) return (uint16_t)((u >> 16) | 0x40u); /* quiet NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

VM_FN float vodb_bf16_to_f32(uint16_t h) { return vm_u2f((uint32_t)h << 16); }

VM_FN uint16_t vodb_f32_to_f16(float x) {
  uint32_t u = vm_f2u(x);
  uint32_t sign = (u >> 16) & 0x8000u;
  uint32_t a = u & 0x7fffffffu;
  if (a > VM_INF_BITS) return (uint16_t)(sign | 0x7e00u);  /* NaN */
  if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* >= 65520 -> inf */
  if (a < 0x33000001u) return (uint16_t)sign;              /* <= 2^-25 -> 0 (ties to even) */
  if (a < 0x38800000u) {                                   /* subnormal half */
    uint32_t e = a >> 23;                                  /* 102..112 */
    uint32_t m = (a & 0x007fffffu) | 0x00800000u;
    uint32_t shift = 126u - e;                             /* 14..24 */
    uint32_t r = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1u);
    if (rem > half || (rem == half && (r & 1u))) r += 1u;
    return (uint16_t)(sign | r);
  }
  uint32_t r = a - 0x38000000u; /* rebias exponent 127 -> 15 */
  r += 0x0fffu + ((r >> 13) & 1u);
  return (uint16_t)(sign | (r >> 13));
}

VM_FN float vodb_f16_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu;
  uint32_t m = h & 0x3ffu;
  if (e == 0) {
    if (m == 0) return vm_u2f(sign);
    /* subnormal: value = m * 2^-24 (exact) */
    float v = VM_MUL((float)m, 5.9604644775390625e-8f);
    return vm_u2f(vm_f2u(v) | sign);
  }
  if (e == 31) return vm_u2f(sign | VM_INF_BITS | (m << 13));
  return vm_u2f(sign | ((e + 112u) << 23) | (m << 13));
}

#endif /* VODB_MATH_H_ */
