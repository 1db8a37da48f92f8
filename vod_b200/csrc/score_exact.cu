// score_exact.cu — fp32-exact scoring fused with the top-k candidate filter (CUDA cores).
//
// Replaces the scoring half of `faiss_index.search(query_vec, k)` for an IndexFlatIP
// (reference src/vod_search/faiss_search/server.py:84; index built at build.py:60-73):
// scores[q, r] = sum_d queries[q, d] * corpus[r, d] with float32 FMA accumulation over the stored
// values (the store may be f32, bf16 or f16; elements are widened exactly). The score matrix never
// reaches HBM: each CTA computes a [128 corpus rows x BN queries] tile in registers and appends only
// entries with score >= tau[q] (the running k-th best, see select.cu) to the per-query candidate list.
//
// Roofline: fp32 FMA (2*nq*rows*D flop) — compute bound on CUDA cores for nq >~ 16; see DESIGN.md.
#include "common.cuh"

namespace vodb {

namespace {

constexpr int BM = 128;  // corpus rows per CTA tile
constexpr int BK = 32;   // K elements per smem stage
constexpr int APAD = 4;

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
  __device__ static void unpack(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <>
struct Vec4<__nv_bfloat16> {
  using type = uint2;
  __device__ static void unpack(const uint2& v, float (&o)[4]) {
    o[0] = __uint_as_float(v.x << 16); o[1] = __uint_as_float(v.x & 0xffff0000u);
    o[2] = __uint_as_float(v.y << 16); o[3] = __uint_as_float(v.y & 0xffff0000u);
  }
};
template <>
struct Vec4<__half> {
  using type = uint2;
  __device__ static void unpack(const uint2& v, float (&o)[4]) {
    __half2 a = *reinterpret_cast<const __half2*>(&v.x);
    __half2 b = *reinterpret_cast<const __half2*>(&v.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    o[0] = fa.x; o[1] = fa.y; o[2] = fb.x; o[3] = fb.y;
  }
};

// Shared-memory tiles are stored K-major ([k][row]) with the row index XOR-swizzled by the k group: a thread that
// loaded 4 consecutive k of one row writes them to 4 different tile rows, and without the swizzle the 8 lanes that
// share a corpus row land on 2 banks (4-way conflicts: 17% of the LSU pipe in the first ncu capture).
// position(k, row) = row ^ (((k >> 2) & 7) << 2): multiples of 4, so the float4 / float2 reads stay aligned.

// BN = 32 / 64 queries per tile; 8 corpus rows x TN queries per thread (TN = 8 for BN = 64, else 4); 128 threads
template <int BN>
struct ExactCfg {
  static constexpr int TN = BN >= 64 ? 8 : 4;
  static constexpr int kThreads = 16 * (BN / TN);   // 128
  static constexpr int kMinBlocks = BN == 64 ? 3 : 4;  // register caps 170 / 128
};

template <typename T, int BN>
__global__ void __launch_bounds__(ExactCfg<BN>::kThreads, ExactCfg<BN>::kMinBlocks)
score_exact_kernel(const T* __restrict__ corpus, int pitch, int64_t row_begin, int64_t row_end,
                   const float* __restrict__ queries, int nq, float* __restrict__ cand_s,
                   int32_t* __restrict__ cand_i, int* __restrict__ cnt, const float* __restrict__ tau,
                   int* __restrict__ overflow, int cap, int dump, int n_qtiles) {
  constexpr int TN = ExactCfg<BN>::TN;
  constexpr int NT = ExactCfg<BN>::kThreads;
  constexpr int TXN = BN / TN;                                   // threads along the query dimension
  constexpr int A_LOADS = BM * BK / 4 / NT;                      // vec4 loads per thread (8 / 4)
  constexpr int B_LOADS = (BN * BK / 4 + NT - 1) / NT;           // 2 / 4 / 4
  using V = typename Vec4<T>::type;

  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN + APAD];
  __shared__ float tau_s[BN];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  // query tile fastest: the CTAs that share a corpus tile are launched together and find it in L2
  const int64_t r0 = row_begin + (int64_t)(blockIdx.x / n_qtiles) * BM;
  const int q0 = (int)(blockIdx.x % n_qtiles) * BN;

  pdl_launch_dependents();
  pdl_wait();  // staged queries, tau and the lists come from the preceding kernels of the stream
  if (tid < BN) tau_s[tid] = (q0 + tid < nq) ? tau[q0 + tid] : INFINITY;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  V a_reg[A_LOADS];
  float4 b_reg[B_LOADS];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LOADS; ++i) {
      int idx = tid + i * NT;
      int row = idx >> 3, kq = idx & 7;
      int64_t r = r0 + row;
      if (r < row_end) a_reg[i] = *reinterpret_cast<const V*>(corpus + (size_t)r * pitch + k0 + kq * 4);
      else a_reg[i] = V{};
    }
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i) {
      int idx = tid + i * NT;
      int row = idx >> 3, kq = idx & 7;
      b_reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < BN && q0 + row < nq)
        b_reg[i] = *reinterpret_cast<const float4*>(queries + (size_t)(q0 + row) * pitch + k0 + kq * 4);
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < A_LOADS; ++i) {
      int idx = tid + i * NT;
      int row = idx >> 3, kq = idx & 7;
      float v[4];
      Vec4<T>::unpack(a_reg[i], v);
      const int pos = row ^ (kq << 2);  // position of this row for k = kq * 4 + j, j = 0..3
#pragma unroll
      for (int j = 0; j < 4; ++j) As[kq * 4 + j][pos] = v[j];
    }
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i) {
      int idx = tid + i * NT;
      int row = idx >> 3, kq = idx & 7;
      if (row < BN) {
        const int pos = row ^ (kq << 2);
        Bs[kq * 4 + 0][pos] = b_reg[i].x;
        Bs[kq * 4 + 1][pos] = b_reg[i].y;
        Bs[kq * 4 + 2][pos] = b_reg[i].z;
        Bs[kq * 4 + 3][pos] = b_reg[i].w;
      }
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < pitch; k0 += BK) {
    __syncthreads();  // previous tile fully consumed
    store_tiles();
    __syncthreads();
    if (k0 + BK < pitch) load_tiles(k0 + BK);  // prefetch next tile into registers
    // 8 groups of 4 k: the swizzle term is constant inside a group, so the group loop stays rolled (one XOR per
    // operand and group instead of 32 precomputed addresses) and the 4 x 8 x TN FMAs of a group give the ILP
#pragma unroll 1
    for (int kg = 0; kg < BK / 4; ++kg) {
      const int s4 = kg << 2;
      const int pa0 = (ty * 4) ^ s4, pa1 = (64 + ty * 4) ^ s4;
      const int pb0 = (tx * 4) ^ s4, pb1 = (BN / 2 + tx * 4) ^ s4;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = kg * 4 + kk;
        float a[8], b[TN];
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][pa0]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][pa1]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][pb0]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if constexpr (TN == 8) {
          float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][pb1]);
          b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // epilogue: threshold filter + append
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    int64_t r = r0 + m;
    if (r >= row_end) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n;
      if constexpr (TN == 8) n = (j < 4) ? (tx * 4 + j) : (BN / 2 + tx * 4 + (j - 4));
      else n = tx * TN + j;
      int q = q0 + n;
      float s = acc[i][j];
      if (dump) {  // first segment: every score is a candidate, slot = row - row_begin (no atomics)
        if (q < nq) {
          cand_s[(size_t)q * cap + (size_t)(r - row_begin)] = s;
          cand_i[(size_t)q * cap + (size_t)(r - row_begin)] = (int32_t)r;
        }
      } else if (q < nq && s >= tau_s[n]) {
        int pos = atomicAdd(&cnt[(size_t)q * kCntStride], 1);
        if (pos < cap) {
          cand_s[(size_t)q * cap + pos] = s;
          cand_i[(size_t)q * cap + pos] = (int32_t)r;
        } else {
          *overflow = 1;
        }
      }
    }
  }
}

template <typename T>
int launch_typed(const SegmentArgs& a, cudaStream_t stream) {
  int64_t rows = a.row_end - a.row_begin;
  if (rows <= 0) return VODB_OK;
  int64_t tiles = (rows + BM - 1) / BM;
  const T* corpus = reinterpret_cast<const T*>(a.corpus);
  const float* q = reinterpret_cast<const float*>(a.queries);
  cudaError_t e;
  // one CTA per (corpus tile, query tile); 64-query tiles (8 x 8 outputs per thread) above 32 queries. A 128-query
  // tile with the same thread tile needs 256 threads x 143 registers = one CTA per SM and measured 10% slower.
  if (a.nq > 32) {
    const int nqt = (a.nq + 63) / 64;
    if (tiles * nqt > 0x7fffffffLL) {
      set_error("launch_score_exact: %lld tiles x %d query tiles exceed the grid limit", (long long)tiles, nqt);
      return VODB_EUNSUPPORTED;
    }
    e = launch_pdl(score_exact_kernel<T, 64>, dim3((unsigned)(tiles * nqt)), dim3(ExactCfg<64>::kThreads), 0, stream, corpus,
                   a.pitch, a.row_begin, a.row_end, q, a.nq, a.cand_s, a.cand_i, a.cnt, a.tau, a.overflow, a.cap,
                   a.dump ? 1 : 0, nqt);
  } else {
    e = launch_pdl(score_exact_kernel<T, 32>, dim3((unsigned)tiles), dim3(ExactCfg<32>::kThreads), 0, stream, corpus,
                   a.pitch, a.row_begin, a.row_end, q, a.nq, a.cand_s, a.cand_i, a.cnt, a.tau, a.overflow, a.cap,
                   a.dump ? 1 : 0, 1);
  }
  VODB_CUDA_CHECK(e);
  return VODB_OK;
}

}  // namespace

int launch_score_exact(const SegmentArgs& a, int /*sm_count*/, cudaStream_t stream) {
  switch (a.dtype) {
    case VODB_F32: return launch_typed<float>(a, stream);
    case VODB_BF16: return launch_typed<__nv_bfloat16>(a, stream);
    case VODB_F16: return launch_typed<__half>(a, stream);
  }
  set_error("launch_score_exact: bad dtype %d", a.dtype);
  return VODB_EINVAL;
}

}  // namespace vodb
