// score_tc.cu — tensor-core scoring (tcgen05 + TMA + TMEM) fused with the top-k candidate filter.
//
// Replaces the scoring half of `faiss_index.search(query_vec, k)` when the index is served from GPUs
// (reference src/vod_search/faiss_search/server.py:51-54,84: GpuIndexFlat shards with useFloat16 storage,
// src/vod_configs/search.py:52,71 — cuBLAS GEMM + faiss k-select, score matrix through HBM). Here the
// [corpus rows x queries] score tile lives only in TMEM. Two kernel families share one structure:
//
//   warp 0 (1 thread)  TMA producer: streams 128-row x 64-col corpus boxes (and, unless the query tile is resident
//                      in shared memory, the query boxes; 128-byte swizzle) through an mbarrier ring;
//   warp 1 (1 thread)  MMA issuer: tcgen05.mma kind::f16, M = corpus rows, N = queries (x terms), K=16 per
//                      instruction, fp32 accumulation in one of two TMEM accumulator buffers;
//   warps 2..5         epilogue: tcgen05.ld 32 lanes x 32 columns, compare every score with the query's
//                      running threshold tau[q] (k-th best so far, select.cu) and append the rare survivors
//                      to the per-query candidate list (per-warp staging, deferred slot-reserving atomics).
//
//   score_tc2_kernel<BN, T, RES>  CTA pairs (cta_group::2): 256 corpus rows x BN queries x T query terms per item, each
//                      CTA staging its own 128 rows and its own half of the queries; RES keeps that half resident.
//                      Every 16-bit-store search runs here.
//   score_tc_kernel<BN, T, P>     one CTA per item (cta_group::1): fp32 stores through P bf16 planes, shapes whose query
//                      tile does not fit the resident variant, and the VODB_TC2=0 A/B path.
//
// D (accumulator) row i = TMEM lane i = corpus row; column j = query j of the tile. Persistent CTAs walk
// (corpus tile, query tile) items; with several query tiles the corpus is walked in L2-sized blocks with the query
// tile as the slow index (TcParams::raster_tiles).
//
// Algorithmic work per segment: rows*pitch*2 bytes from HBM, 2*nq*rows*pitch flop.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vodb {

namespace {

constexpr int BM = 128;       // corpus rows per MMA (M)
constexpr int KC = 64;        // K elements per pipeline stage (128 bytes = one swizzle atom)
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int kThreads = 192; // 6 warps
constexpr uint32_t kSmemBudget = 200 * 1024;

// L2 cache-policy descriptors (cute::TMA::CacheHintSm90 values)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (error code at the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("vodb: mbarrier timeout block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x, (void*)bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* smem_dst, int c0, int c1,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, 16-bit inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;               // LBO (unused for swizzled K-major; canonical value 1)
  d |= (uint64_t)(1024 >> 4) << 32;     // SBO: 8 rows x 128 B between 8-row core-matrix groups
  d |= (uint64_t)1 << 46;               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;               // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TcParams {
  int64_t row_begin, row_end;
  int nq;
  int n_ctiles, n_qtiles;
  int kchunks;  // pitch / KC
  int q_rows_pad;  // rows of one query term plane in the staged query tensor [terms][q_rows_pad][pitch]
  int plane_rows;  // rows of one corpus plane in the corpus tensor map (fp32 stores searched through bf16 planes)
  float* cand_s;
  int32_t* cand_i;
  int* cnt;
  const float* tau;
  int* overflow;
  int cap;
  uint32_t idesc;
  int dump;  // first segment: store every score at slot (row - row_begin) instead of filtering
  const int* term_any;  // per prepare-block masks of the query terms that are nonzero anywhere (api.cu prepare_kernel)
  int term_blocks;
  // Large multi-term batches are launched twice: the 2-CTA one-term kernel with kOnlyIfSingleTerm and the multi-term
  // kernel with kOnlyIfMultiTerm. Exactly one of them does the work, chosen on the device from term_any; the other
  // finds zero items and retires in a few microseconds.
  int term_policy;
  // Item order when a corpus tile meets several query tiles (n_qtiles > 1). 0: query tile fastest — every CTA works
  // on the same one or two corpus tiles at the same moment, so their first loads of a tile reach the L2 together,
  // before the line has arrived from HBM, and are fetched more than once (ncu: 2.9x the algorithmic DRAM bytes).
  // > 0: the corpus is walked in blocks of `raster_tiles` tiles (one per CTA / CTA pair, ~30 MB, L2 resident); inside
  // a block the QUERY tile is the slow index: during the first pass every CTA streams a different corpus tile from
  // HBM (each line is requested once), the remaining n_qtiles - 1 passes find the block in L2.
  int raster_tiles;
  // The last query tile of a one-term batch is staged through its own tensor map whose box is only as tall as the
  // tile's MMA width (item_columns): a batch of 160 queries moves 160, not 256, query rows per K chunk over the
  // L2 -> SM path, which is the path that limits these kernels. Bytes of that box (per CTA, per term).
  uint32_t last_box_bytes;
  int n_stages;  // resident-query pair kernel: depth of the corpus ring (what is left of shared memory), else unused
};
constexpr int kTermAlways = 0, kOnlyIfSingleTerm = 1, kOnlyIfMultiTerm = 2;

// MMA N of an item = its query columns rounded up to 32 (not the full tile): the last query tile of a batch costs
// tensor-pipe time in proportion to the queries it holds. 32 because the epilogue reads the accumulator 32 columns
// at a time and must only see columns this item's MMAs wrote (the staged query rows up to the tile end are zero, so
// those columns hold 0 and fail against tau = +inf; never-written TMEM could hold a +inf bit pattern).
__device__ __forceinline__ int item_columns(const TcParams& p, int q0, int bn) {
  const int rem = p.nq - q0;
  return rem >= bn ? bn : ((rem + 31) & ~31);
}
__device__ __forceinline__ uint32_t idesc_with_n(uint32_t idesc, int n) {
  return (idesc & ~(0x3fu << 17)) | ((uint32_t)(n >> 3) << 17);
}

__device__ __forceinline__ void item_to_tiles(const TcParams& p, int item, int& ct, int& qt) {
  if (p.raster_tiles <= 0) {
    ct = item / p.n_qtiles;
    qt = item - ct * p.n_qtiles;
    return;
  }
  const int per_block = p.raster_tiles * p.n_qtiles;
  const int blk = item / per_block;
  const int r = item - blk * per_block;
  const int c0 = blk * p.raster_tiles;
  const int bc = min(p.raster_tiles, p.n_ctiles - c0);  // the last block may be short
  qt = r / bc;
  ct = c0 + (r - qt * bc);
}

// Per-warp staging of filter survivors (shared memory). Each epilogue warp owns two small buffers: survivors of
// item i are pushed with shared-memory atomics into buffer i&1 and appended to the global per-query lists at the
// start of item i+1 by the same warp, 32 lanes wide, so the L2 atomic round trip (microseconds when the L2 is
// saturated by the operand stream) is paid once per batch of entries and overlaps the next item's work. Nothing is
// shared between the four epilogue warps: they never synchronise with each other.
constexpr int kWarpStageCap = 128;
struct WarpStage {
  int count[2];
  int pad[2];
  float s[2][kWarpStageCap];
  int32_t row[2][kWarpStageCap];
  int32_t q[2][kWarpStageCap];
};

__device__ __noinline__ void append_global(int* cnt, float* cand_s, int32_t* cand_i, int* overflow, int cap, int q,
                                           float s, int32_t row) {
  const int pos = atomicAdd(&cnt[(size_t)q * kCntStride], 1);
  if (pos < cap) {
    cand_s[(size_t)q * cap + pos] = s;
    cand_i[(size_t)q * cap + pos] = row;
  } else {
    *overflow = 1;
  }
}

// `issue` copies the staged entries of buffer `b` to registers and fires the slot-reserving atomics; `complete`
// (called after the item's score columns have been processed) consumes the slots.
constexpr int kFlushPerLane = kWarpStageCap / 32;
struct PendingFlush {
  float s[kFlushPerLane];
  int32_t row[kFlushPerLane];
  int32_t q[kFlushPerLane];
  int pos[kFlushPerLane];
};

__device__ __forceinline__ void flush_issue(const TcParams& p, const WarpStage& ws, int b, int lane, PendingFlush& f) {
  const int n = min(ws.count[b], kWarpStageCap);
#pragma unroll
  for (int u = 0; u < kFlushPerLane; ++u) {
    const int e = lane + u * 32;
    f.pos[u] = -1;
    if (e < n) {
      f.s[u] = ws.s[b][e];
      f.row[u] = ws.row[b][e];
      f.q[u] = ws.q[b][e];
      f.pos[u] = atomicAdd(&p.cnt[(size_t)f.q[u] * kCntStride], 1);
    }
  }
}

__device__ __forceinline__ void flush_complete(const TcParams& p, PendingFlush& f) {
#pragma unroll
  for (int u = 0; u < kFlushPerLane; ++u) {
    if (f.pos[u] >= 0) {
      if (f.pos[u] < p.cap) {
        p.cand_s[(size_t)f.q[u] * p.cap + f.pos[u]] = f.s[u];
        p.cand_i[(size_t)f.q[u] * p.cap + f.pos[u]] = f.row[u];
      } else {
        *p.overflow = 1;
      }
      f.pos[u] = -1;
    }
  }
}

// v[j] for a run-time j (registers cannot be indexed dynamically): 5-level select tree, 31 SEL, no branches
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

// ---- epilogue building blocks shared by the 1-CTA and the 2-CTA kernel ---------------------------------------

// start of an item: fire the slot-reserving atomics for the previous item's survivors, load this query tile's
// thresholds into the warp's own shared-memory copy (queries past nq never pass), recycle the flushed buffer
template <int BN>
__device__ __forceinline__ void epilogue_begin_item(const TcParams& p, WarpStage& ws, float* tau_cur, int q0, int sb,
                                                    int local, int lane, PendingFlush& pend) {
  if (!p.dump) {
    // thresholds first: their loads must not queue behind the flush's atomics in the memory pipeline
    if (q0 + BN <= p.nq) {
      for (int c = lane * 4; c < BN; c += 128)
        *reinterpret_cast<float4*>(tau_cur + c) = *reinterpret_cast<const float4*>(p.tau + q0 + c);
    } else {
      for (int c = lane; c < BN; c += 32) tau_cur[c] = (q0 + c < p.nq) ? p.tau[q0 + c] : INFINITY;
    }
    if (local > 0) flush_issue(p, ws, sb ^ 1, lane, pend);
  }
  __syncwarp();
  if (lane == 0) ws.count[sb ^ 1] = 0;  // every lane has read it; next pushed to two items from now
}

// the BN score columns of this thread's corpus row: dump them (first segment) or filter against tau and push the
// rare survivors into the warp's staging buffer
// GROUPS > 1: the score is the sum of GROUPS accumulator column groups BN apart — [0, BN) holds the leading product,
// the others the corrections, which are summed first and added to the leading product last
// `groups` (<= GROUPS, uniform over the CTA) is how many of them this launch actually wrote: correction terms that are
// zero for the whole query batch are skipped by the producer and the MMA warp, and their columns hold nothing.
// HALF > 0 (CTA-pair kernels with the terms stacked in one MMA): the accumulator holds, for each CTA's half of the
// item's queries (HALF = BN/2 of them), `groups` column groups HALF apart — [half h][term t][query j] — instead of
// `groups` groups BN apart.
template <int BN, int GROUPS = 1, int HALF = 0>
__device__ __forceinline__ void epilogue_columns(const TcParams& p, WarpStage& ws, const float* tau_cur,
                                                 uint32_t taddr0, int64_t row, bool valid, int q0, int sb, int groups,
                                                 int n_cols = BN) {
#pragma unroll 1
  for (int c0 = 0; c0 < n_cols; c0 += 32) {
    uint32_t v[32];
    const int gs = HALF > 0 ? HALF : BN;                                                   // distance between groups
    const int cb = HALF > 0 ? (c0 / HALF) * (groups * HALF) + (c0 % HALF) : c0;             // column of the leading group
    tmem_ld_32x32b_x32(taddr0 + (uint32_t)cb, v);
    if (GROUPS >= 3 && groups >= 3) {
      uint32_t w[32], x[32];
      tmem_ld_32x32b_x32(taddr0 + (uint32_t)(cb + gs), w);
      tmem_ld_32x32b_x32(taddr0 + (uint32_t)(cb + 2 * gs), x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), __fadd_rn(__uint_as_float(w[j]), __uint_as_float(x[j]))));
    } else if (GROUPS >= 2 && groups >= 2) {
      uint32_t w[32];
      tmem_ld_32x32b_x32(taddr0 + (uint32_t)(cb + gs), w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), __uint_as_float(w[j])));
    } else {
      tmem_ld_wait();
    }
    if (p.dump) {
      // first segment: every score is a candidate; slot = row - row_begin, no atomics, coalesced over lanes
      if (valid) {
        const size_t slot = (size_t)(row - p.row_begin);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int q = q0 + c0 + j;
          if (q < p.nq) {
            p.cand_s[(size_t)q * p.cap + slot] = __uint_as_float(v[j]);
            p.cand_i[(size_t)q * p.cap + slot] = (int32_t)row;
          }
        }
      }
      continue;
    }
    bool any = false;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 t = *reinterpret_cast<const float4*>(tau_cur + c0 + j4 * 4);
      any |= (__uint_as_float(v[j4 * 4 + 0]) >= t.x) | (__uint_as_float(v[j4 * 4 + 1]) >= t.y) |
             (__uint_as_float(v[j4 * 4 + 2]) >= t.z) | (__uint_as_float(v[j4 * 4 + 3]) >= t.w);
    }
    if (any && valid) {
      // rare, divergent: this lane (= corpus row) has survivors among the 32 query columns. Build the column
      // bitmask from registers, reserve staging slots with ONE shared-memory atomic, then push each set bit.
      uint32_t m = 0;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 t = *reinterpret_cast<const float4*>(tau_cur + c0 + j4 * 4);
        m |= (__uint_as_float(v[j4 * 4 + 0]) >= t.x ? 1u : 0u) << (j4 * 4 + 0);
        m |= (__uint_as_float(v[j4 * 4 + 1]) >= t.y ? 1u : 0u) << (j4 * 4 + 1);
        m |= (__uint_as_float(v[j4 * 4 + 2]) >= t.z ? 1u : 0u) << (j4 * 4 + 2);
        m |= (__uint_as_float(v[j4 * 4 + 3]) >= t.w ? 1u : 0u) << (j4 * 4 + 3);
      }
      int idx = atomicAdd(&ws.count[sb], __popc(m));
#pragma unroll 1
      while (m != 0u) {
        const int j = __ffs(m) - 1;
        m &= m - 1u;
        const float sc = __uint_as_float(pick32(v, j));
        const int q = q0 + c0 + j;
        if (idx < kWarpStageCap) {
          ws.s[sb][idx] = sc;
          ws.row[sb][idx] = (int32_t)row;
          ws.q[sb][idx] = q;
        } else {
          append_global(p.cnt, p.cand_s, p.cand_i, p.overflow, p.cap, q, sc, (int32_t)row);  // staging full: slow, correct
        }
        ++idx;
      }
    }
  }
}

// BN = queries per tile (MMA N); T = query terms: a float32 query is split into T 16-bit terms (hi, lo, lo2), so T=2
// keeps ~16 and T=3 all 24 mantissa bits of the query while the corpus tile is loaded from HBM/L2 only once per
// stage. With T > 1 an item has TWO accumulators: the leading product (hi x hi) and the sum of the correction
// products, added once in the epilogue. The tensor core truncates (does not round) every accumulation, a bias of up
// to one ulp of the accumulator per MMA; keeping the small corrections apart leaves kchunks*4 truncations at full
// magnitude instead of kchunks*4*T (or *6 with corpus planes): 1.5e-6 instead of 1e-5 relative at dim 768.
//
// P > 1 (fp32 store, api.cu `ensure_planes`): the corpus is held as P bf16 planes c = c_0 + c_1 + c_2 (each the bf16
// rounding of what the previous ones left, so 3 planes carry all 24 mantissa bits) and the queries as T = P terms; the
// kernel accumulates the products c_p * q_t with p + t < P (6 of 9 for P = 3: the dropped ones are below 2^-24 of
// the leading product), every one exact in the fp32 accumulator. A pipeline stage is one (K chunk, plane) pair:
// the plane's corpus box plus the P - p query boxes it is multiplied with.
template <int BN, int T, int P = 1>
struct TcConfig {
  static_assert(P == 1 || P == T, "corpus planes come with as many query terms");
  static constexpr uint32_t kABytes = BM * KC * 2;
  static constexpr uint32_t kBBytes = BN * KC * 2;          // one term
  static constexpr uint32_t kStageBytes = kABytes + T * kBBytes;
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  // kConcat: the T query boxes of a stage sit back to back in shared memory, i.e. they ARE one K-major tile of T*BN
  // rows, so ONE MMA with N = T*BN multiplies the corpus tile with all terms at once and every term gets its own
  // accumulator column group: a third of the MMAs and of the corpus-tile reads of the one-MMA-per-term form, which
  // is what keeps 3-term searches of <= 64 queries HBM-bound. Needs T*BN <= 256 (UMMA N) and two such buffers in TMEM.
  static constexpr bool kConcat = (T > 1) && (T * BN <= 256);
  static constexpr bool kDual = (T > 1) && !kConcat;          // one MMA per term, leading / correction accumulators
  static constexpr int kGroups = kConcat ? T : (kDual ? 2 : 1);
  static constexpr uint32_t kAccCols = kGroups * BN;          // TMEM columns of one accumulator buffer
  static_assert(2 * kAccCols <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  static constexpr uint32_t kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                        : (2 * kAccCols <= 256) ? 256 : 512;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                                         4 * BN * sizeof(float) /*tau, one copy per epilogue warp*/ + 4 * sizeof(WarpStage);
};

template <int BN, int T, int P = 1>
__global__ void __launch_bounds__(kThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tmap_corpus, const __grid_constant__ CUtensorMap tmap_query,
                const __grid_constant__ CUtensorMap tmap_query_last, const TcParams p) {
  using Cfg = TcConfig<BN, T, P>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle (TMA writes and UMMA reads XOR address bits [4,7) with [7,10))
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;                                   // STAGES x [128 x 128B]
  unsigned char* smem_b = smem + STAGES * Cfg::kABytes;           // STAGES x T x [BN x 128B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* tau_s = reinterpret_cast<float*>(bars + 2 * STAGES + 6);  // [4 epilogue warps][BN]
  WarpStage* wst = reinterpret_cast<WarpStage*>(tau_s + 4 * BN);   // [4 epilogue warps]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    for (int w = 0; w < 4; ++w) wst[w].count[0] = wst[w].count[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation) overlapped the previous kernel's tail; from here on the kernel
  // reads what its predecessors wrote (staged queries, tau, cnt) and appends to the lists
  pdl_launch_dependents();
  pdl_wait();

  int n_items = p.n_ctiles * p.n_qtiles;
  // query terms in use: all T, unless the prepare kernel found the trailing correction terms empty for the whole
  // batch (float32 queries that are exact in the store dtype) — then they are neither loaded nor multiplied, and
  // the result is bit-identical to the full computation. Corpus planes (P > 1) always run in full.
  int nt_run = T;
  if constexpr (T > 1 && P == 1) {
    int m = 0;
    for (int i = threadIdx.x; i < p.term_blocks; i += kThreads) m |= p.term_any[i];
    const int any1 = __syncthreads_or(m & 2), any2 = __syncthreads_or(m & 4);  // logical ORs over the CTA
    nt_run = (T >= 3 && any2) ? 3 : any1 ? 2 : 1;
    if (p.term_policy == kOnlyIfMultiTerm && nt_run == 1) n_items = 0;  // the one-term pair kernel has this batch
  }

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_corpus) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_query) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_query_last) : "memory");
      const uint64_t corpus_policy = (p.n_qtiles == 1) ? kEvictFirst : kEvictLast;
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int ct, qt;
        item_to_tiles(p, item, ct, qt);
        const int row0 = (int)(p.row_begin + (int64_t)ct * BM);
        const int q0 = qt * BN;
        const bool last_q = (T == 1 && P == 1 && qt == p.n_qtiles - 1);  // narrower box for the last query tile
        const CUtensorMap* qmap = last_q ? &tmap_query_last : &tmap_query;
        const uint32_t b_bytes = last_q ? p.last_box_bytes : Cfg::kBBytes;
        for (int kc = 0; kc < p.kchunks; ++kc) {
#pragma unroll
          for (int pl = 0; pl < P; ++pl) {
            const int nt = (P == 1) ? nt_run : (P - pl);  // query terms multiplied with this plane
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], Cfg::kABytes + (uint32_t)nt * b_bytes);
            // blocked order: the last pass over a block lets its lines go (the next block needs the room)
            tma_load_2d(&tmap_corpus, &full_bar[stage], smem_a + stage * Cfg::kABytes, kc * KC,
                        pl * p.plane_rows + row0,
                        (p.raster_tiles > 0 && qt == p.n_qtiles - 1) ? kEvictFirst : corpus_policy);
#pragma unroll
            for (int t = 0; t < T; ++t)
              if (t < nt)
                tma_load_2d(qmap, &full_bar[stage], smem_b + (stage * T + t) * Cfg::kBBytes, kc * KC,
                            t * p.q_rows_pad + q0, kEvictLast);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        uint32_t idesc_item = p.idesc;
        if constexpr (T == 1 && P == 1) {  // one term: the MMA is as wide as the item's queries (multiple of 32)
          int ct, qt;
          item_to_tiles(p, item, ct, qt);
          idesc_item = idesc_with_n(p.idesc, item_columns(p, qt * BN, BN));
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * Cfg::kAccCols;  // leading product; corrections at + BN
        for (int kc = 0; kc < p.kchunks; ++kc) {
#pragma unroll
          for (int pl = 0; pl < P; ++pl) {
            const int nt = (P == 1) ? nt_run : (P - pl);
            mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
            tcgen05_fence_after();
            const uint64_t da = make_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
            if constexpr (Cfg::kConcat) {
              // one MMA per K step over the nt stacked query boxes (N = nt * BN). Plane 0 fills every column group
              // (group t = c_0 * q_t); the later planes land one group further right, so that group 0 stays the
              // leading product c_0 * q_0 alone (truncation bias, see TcConfig) and the corrections share groups 1..
              const uint64_t db = make_desc_sw128(smem_u32(smem_b + (stage * T) * Cfg::kBBytes));
              const uint32_t idesc = (p.idesc & ~(0x3fu << 17)) | ((uint32_t)((BN * nt) >> 3) << 17);
              const uint32_t d = tmem_d + (pl == 0 ? 0u : (uint32_t)BN);
#pragma unroll
              for (int k = 0; k < KC / UMMA_K; ++k)
                umma_f16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | pl | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int t = 0; t < T; ++t) {
                if (t >= nt) break;
                const uint64_t db = make_desc_sw128(smem_u32(smem_b + (stage * T + t) * Cfg::kBBytes));
#pragma unroll
                for (int k = 0; k < KC / UMMA_K; ++k) {
                  // advance 32 bytes (16 elements) inside the 128-byte swizzle atom: +2 in the >>4 address field
                  const bool lead = (pl == 0 && t == 0);
                  const bool first = (kc | k) == 0 && (lead || (pl == 0 && t == 1));  // first MMA into its accumulator
                  umma_f16(lead || !Cfg::kDual ? tmem_d : tmem_d + (uint32_t)BN, da + (uint64_t)(2 * k),
                           db + (uint64_t)(2 * k), idesc_item, first ? 0u : 1u);
                }
              }
            }
            umma_commit(&empty_bar[stage]);  // frees the smem stage once the MMAs above have read it
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ---------------- epilogue warps (2..5): independent of each other ----------------
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int ew = warp - 2;
    WarpStage& ws = wst[ew];
    float* tau_cur = tau_s + ew * BN;
    int local = 0;
    PendingFlush pend;
#pragma unroll
    for (int u = 0; u < kFlushPerLane; ++u) pend.pos[u] = -1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
      int ct, qt;
      item_to_tiles(p, item, ct, qt);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int q0 = qt * BN;
      const int sb = local & 1;  // staging buffer that receives this item's survivors
      epilogue_begin_item<BN>(p, ws, tau_cur, q0, sb, local, lane, pend);

      const int64_t row = p.row_begin + (int64_t)ct * BM + quarter * 32 + lane;
      const bool valid = row < p.row_end;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * Cfg::kAccCols;
      epilogue_columns<BN, Cfg::kGroups>(p, ws, tau_cur, taddr0, row, valid, q0, sb,
                                          P == 1 ? (Cfg::kConcat ? nt_run : (nt_run > 1 ? 2 : 1)) : Cfg::kGroups,
                                          (T == 1 && P == 1) ? item_columns(p, q0, BN) : BN);
      // all TMEM reads of this buffer are complete (wait::ld above): hand it back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      flush_complete(p, pend);  // slots reserved at the top of this item have arrived by now
    }
    __syncwarp();
    if (!p.dump && local > 0) {
      flush_issue(p, ws, (local - 1) & 1, lane, pend);
      flush_complete(p, pend);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

// ---- 2-CTA variant (cta_group::2) for large batches ------------------------------------------------------------
//
// A CTA pair (cluster of 2, one TPC) computes a [256 corpus rows x 256 queries] item with ONE tcgen05.mma issued by
// the leader: each CTA stages only its own 128 corpus rows (A half) and its own 128 queries (B half) per K chunk
// — 32 KB per stage instead of 48 KB for the same MMA work, i.e. 1/3 less L2->SM traffic and shared-memory write
// bandwidth per flop than the 1-CTA 128x256 tile. Both CTAs keep their own 128 x 256 fp32 accumulator (two buffers,
// all 512 TMEM columns) and run the same epilogue as the 1-CTA kernel on it.
//   leader (cluster rank 0): TMA producer for its halves, MMA issuer for the pair, epilogue for rows 0..127
//   follower (rank 1):       TMA producer for its halves (completes on the LEADER's full barrier), epilogue 128..255
//   barriers: full[s] (leader) <- tx bytes of both CTAs; empty[s], tfull[a] (both CTAs) <- tcgen05.commit multicast;
//             tempty[a] (leader) <- the 8 epilogue warps of the pair (follower warps arrive remotely).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (-> leader CTA)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const void* tmap, uint64_t* leader_bar, void* smem_dst, int c0, int c1,
                                                uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1),
        "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair once the MMAs issued so far finish
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

constexpr uint32_t kSmemMax = 227 * 1024;  // dynamic shared memory one CTA may opt into on sm_100

// BN = queries per pair item, T = query terms (see TcConfig). Each CTA stages its own 128 corpus rows and, per term,
// its own half (BN/2) of the item's queries: kABytes + T * kBBytes per stage against kABytes + T * 2 * kBBytes in the
// 1-CTA kernel for the same MMA work per SM, i.e. more stages in flight (7 instead of 5 at <64,3>, 5 instead of 3 at
// <128,3>) and half the query bytes per SM on the L2 -> SM path. Stacked terms (T * BN <= 256): ONE MMA of N = nt * BN per
// K step; CTA h's B tile holds [term][query of half h], so the accumulator columns are [half][term][query] (epilogue
// HALF = BN/2). Otherwise (<128,3>): one MMA of N = BN per term, leading product and corrections in two accumulators.
template <int BN, int T>
struct Tc2Config {
  static constexpr uint32_t kABytes = BM * KC * 2;          // this CTA's 128 corpus rows
  static constexpr uint32_t kBBytes = (BN / 2) * KC * 2;    // one term, this CTA's half of the item's queries
  static constexpr uint32_t kStageBytes = kABytes + T * kBBytes;
  static constexpr uint32_t kExtra = 1024 + 256 + 4 * BN * sizeof(float) + 4 * sizeof(WarpStage);
  static constexpr int kFit = (kSmemMax - kExtra) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr bool kConcat = (T > 1) && (T * BN <= 256);
  static constexpr bool kDual = (T > 1) && !kConcat;
  static constexpr int kGroups = kConcat ? T : (kDual ? 2 : 1);
  static constexpr int kHalf = kConcat ? BN / 2 : 0;
  static constexpr uint32_t kAccCols = kGroups * BN;
  static_assert(2 * kAccCols <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  static constexpr uint32_t kTmemCols = (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128 : (2 * kAccCols <= 256) ? 256 : 512;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kExtra;
  static_assert(kStages >= 3, "pipeline too shallow");
  static constexpr int kMaxStagesRes = 12;  // resident variant: barrier slots (2 * 12 + 6 words fit the 256-byte area)
};

// RES (one query tile, BN <= 128; one term, or stacked terms): the CTA's half of the query tile — all K chunks of it,
// kchunks x T x BN/2 rows x 128 B (48 KB at 64 one-term queries x 768 dims, 96 KB at 128, 144 KB at 64 three-term
// queries) — is loaded ONCE into shared memory and stays there; the ring then carries corpus boxes only (10 / 7 / 4
// stages of 16 KB). The L2 -> SM path, which caps these kernels at ~6300 B/clk for the whole chip (B300_MICROARCH.md
// "LTS throughput cap"), carries the corpus bytes and nothing else: 1.0x the HBM stream instead of 1.5x (64 queries,
// 1-CTA kernel), 2x (128 queries) or 1.75x (64 three-term queries streamed with the corpus).
template <int BN, int T, bool RES = false>
__global__ void __launch_bounds__(kThreads, 1)
score_tc2_kernel(const __grid_constant__ CUtensorMap tmap_corpus, const __grid_constant__ CUtensorMap tmap_query,
                 const __grid_constant__ CUtensorMap tmap_query_last, const TcParams p) {
  using Cfg = Tc2Config<BN, T>;
  static_assert(!RES || (BN <= 128 && (T == 1 || Cfg::kConcat)), "resident queries: one tile of at most 128 queries, stacked terms");
  const int STAGES = RES ? p.n_stages : Cfg::kStages;
  constexpr int kBarSlots = RES ? Cfg::kMaxStagesRes : Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // RES: [kchunks x T resident query boxes][STAGES corpus boxes]; else [STAGES corpus boxes][STAGES x T query boxes]
  unsigned char* smem_res = smem;
  unsigned char* smem_a = RES ? smem + (size_t)p.kchunks * T * Cfg::kBBytes : smem;
  unsigned char* smem_b = smem + STAGES * Cfg::kABytes;   // STAGES x T x [BN/2 x 128B] (unused when RES)
  uint64_t* bars = reinterpret_cast<uint64_t*>(RES ? smem_a + (size_t)STAGES * Cfg::kABytes : smem + STAGES * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kBarSlots;
  uint64_t* tfull_bar = bars + 2 * kBarSlots;
  uint64_t* tempty_bar = bars + 2 * kBarSlots + 2;
  uint64_t* bfull_bar = bars + 2 * kBarSlots + 4;          // RES: the resident query boxes have landed (both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kBarSlots + 5);
  float* tau_s = reinterpret_cast<float*>(bars + 2 * kBarSlots + 6);
  WarpStage* wst = reinterpret_cast<WarpStage*>(tau_s + 4 * BN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's expect_tx arrival; the bytes of both CTAs complete on it
      mbar_init(&empty_bar[s], 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);  // 4 epilogue warps x 2 CTAs
    }
    mbar_init(bfull_bar, 1);
    for (int w = 0; w < 4; ++w) wst[w].count[0] = wst[w].count[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // the peer's barriers exist before anything can signal them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int n_pairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  int n_items = p.n_ctiles * p.n_qtiles;  // n_ctiles counts 256-row pair tiles here
  // which query terms carry anything (prepare kernel's masks; the same answer in every CTA of the grid)
  int nt_run = T;
  if (T > 1 || p.term_policy != kTermAlways) {
    int m = 0;
    for (int i = threadIdx.x; i < p.term_blocks; i += kThreads) m |= p.term_any[i];
    const int any1 = __syncthreads_or(m & 2), any2 = __syncthreads_or(m & 4);
    const int nt_needed = any2 ? 3 : any1 ? 2 : 1;
    if (T > 1) nt_run = nt_needed < T ? nt_needed : T;
    if (p.term_policy == kOnlyIfSingleTerm && nt_needed > 1) n_items = 0;  // the multi-term launch has this batch
    if (p.term_policy == kOnlyIfMultiTerm && nt_needed == 1) n_items = 0;  // the one-term launch has this batch
  }

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs) ----------------
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_corpus) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_query) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap_query_last) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      if (RES && n_items > 0) {
        // the CTA's half of the (only) query tile, every K chunk of it, once
        if constexpr (T == 1) {
          const int q_half = (int)rank * (item_columns(p, 0, BN) / 2);
          if (leader) mbar_expect_tx(bfull_bar, 2u * (uint32_t)p.kchunks * p.last_box_bytes);
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_2d_2sm(&tmap_query_last, bfull_bar, smem_res + (size_t)kc * Cfg::kBBytes, kc * KC, q_half, kEvictLast);
        } else {
          // stacked terms: [K chunk][term][BN/2 rows], the layout one ring stage has in the streaming variant
          const int q_half = (int)rank * (BN / 2);
          if (leader) mbar_expect_tx(bfull_bar, 2u * (uint32_t)p.kchunks * (uint32_t)nt_run * Cfg::kBBytes);
          for (int kc = 0; kc < p.kchunks; ++kc)
            for (int t = 0; t < nt_run; ++t)
              tma_load_2d_2sm(&tmap_query, bfull_bar, smem_res + ((size_t)kc * T + t) * Cfg::kBBytes, kc * KC,
                              t * p.q_rows_pad + q_half, kEvictLast);
        }
      }
      for (int item = pair; item < n_items; item += n_pairs) {
        int ct, qt;
        item_to_tiles(p, item, ct, qt);
        const int row0 = (int)(p.row_begin + ((int64_t)ct * 2 + rank) * BM);
        // this CTA supplies half of the item's query columns: rows [rank * n/2, (rank + 1) * n/2) of the tile (the box is
        // always BN/2 rows; the MMA reads the first n/2 of them). Only one-term items are narrower than BN.
        const int n_item = (T == 1) ? item_columns(p, qt * BN, BN) : BN;
        const int q0 = qt * BN + (int)rank * (n_item / 2);
        const uint64_t corpus_policy =
            (p.n_qtiles == 1 || (p.raster_tiles > 0 && qt == p.n_qtiles - 1)) ? kEvictFirst : kEvictLast;
        const bool last_q = (T == 1 && qt == p.n_qtiles - 1);  // narrower box for the last query tile
        const CUtensorMap* qmap = last_q ? &tmap_query_last : &tmap_query;
        const uint32_t b_bytes = last_q ? p.last_box_bytes : Cfg::kBBytes;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * (Cfg::kABytes + (RES ? 0u : (uint32_t)nt_run * b_bytes)));
          tma_load_2d_2sm(&tmap_corpus, &full_bar[stage], smem_a + stage * Cfg::kABytes, kc * KC, row0, corpus_policy);
#pragma unroll
          for (int t = 0; t < T; ++t)
            if (!RES && t < nt_run)
              tma_load_2d_2sm(qmap, &full_bar[stage], smem_b + (stage * T + t) * Cfg::kBBytes, kc * KC,
                              t * p.q_rows_pad + q0, kEvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ---------------- MMA issuer (leader only, for the pair) ----------------
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      if (RES && n_items > 0) {
        mbar_wait(bfull_bar, 0);  // resident query boxes of both CTAs
        tcgen05_fence_after();
      }
      for (int item = pair; item < n_items; item += n_pairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        int ct, qt;
        item_to_tiles(p, item, ct, qt);
        // MMA N: one term -> the item's query columns; stacked terms -> nt * BN; one MMA per term -> BN
        const uint32_t idesc_item =
            idesc_with_n(p.idesc, T == 1 ? item_columns(p, qt * BN, BN) : (Cfg::kConcat ? nt_run * BN : BN));
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * Cfg::kAccCols;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t da = make_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          if constexpr (!Cfg::kDual) {
            const uint64_t db = make_desc_sw128(smem_u32(RES ? smem_res + (size_t)kc * T * Cfg::kBBytes
                                                             : smem_b + (stage * T) * Cfg::kBBytes));
#pragma unroll
            for (int k = 0; k < KC / UMMA_K; ++k)
              umma_f16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_item, (kc | k) != 0 ? 1u : 0u);
          } else {
#pragma unroll
            for (int t = 0; t < T; ++t) {
              if (t >= nt_run) break;
              const uint64_t db = make_desc_sw128(smem_u32(smem_b + (stage * T + t) * Cfg::kBBytes));
#pragma unroll
              for (int k = 0; k < KC / UMMA_K; ++k) {
                const bool first = (kc | k) == 0 && t <= 1;  // first MMA into the leading / the correction accumulator
                umma_f16_2sm(t == 0 ? tmem_d : tmem_d + (uint32_t)BN, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k),
                             idesc_item, first ? 0u : 1u);
              }
            }
          }
          umma_commit_2sm(&empty_bar[stage]);  // frees the stage in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[acc]);  // accumulators of both CTAs are complete
      }
    }
  } else {
    // ---------------- epilogue warps (2..5) of both CTAs ----------------
    const int quarter = warp & 3;
    const int ew = warp - 2;
    WarpStage& ws = wst[ew];
    float* tau_cur = tau_s + ew * BN;
    int local = 0;
    PendingFlush pend;
#pragma unroll
    for (int u = 0; u < kFlushPerLane; ++u) pend.pos[u] = -1;
    for (int item = pair; item < n_items; item += n_pairs, ++local) {
      int ct, qt;
      item_to_tiles(p, item, ct, qt);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int q0 = qt * BN;
      const int sb = local & 1;
      epilogue_begin_item<BN>(p, ws, tau_cur, q0, sb, local, lane, pend);

      const int64_t row = p.row_begin + ((int64_t)ct * 2 + rank) * BM + quarter * 32 + lane;
      const bool valid = row < p.row_end;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * Cfg::kAccCols;
      epilogue_columns<BN, Cfg::kGroups, Cfg::kHalf>(p, ws, tau_cur, taddr0, row, valid, q0, sb,
                                                      Cfg::kConcat ? nt_run : (nt_run > 1 ? 2 : 1),
                                                      T == 1 ? item_columns(p, q0, BN) : BN);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty_bar[acc], 0);  // the leader's barrier collects both CTAs' epilogues
      flush_complete(p, pend);
    }
    __syncwarp();
    if (!p.dump && local > 0) {
      flush_issue(p, ws, (local - 1) & 1, lane, pend);
      flush_complete(p, pend);
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the peer can still signal its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols)
                 : "memory");
  }
}

// ---- host side -------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn resolve_encode_fn() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    return reinterpret_cast<EncodeTiledFn>(ptr);
  return nullptr;
}
EncodeTiledFn get_encode_fn() {
  static const EncodeTiledFn fn = resolve_encode_fn();  // C++11 magic static: initialised once, thread safe
  return fn;
}

// 2D row-major [rows, pitch] tensor of 16-bit elements, box = [box_rows, 64 cols], 128B swizzle
int encode_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int pitch, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return VODB_EUNSUPPORTED;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = dtype == VODB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld pitch=%d box_rows=%d)", (int)r,
              (long long)rows, pitch, box_rows);
    return VODB_ECUDA;
  }
  return VODB_OK;
}

// the 2-CTA kernel serves large single-term batches; VODB_TC2=0 forces the 1-CTA kernel (A/B comparisons)
bool pair_kernel_enabled() {
  static const char* env = std::getenv("VODB_TC2");
  return env ? (env[0] != '0') : true;
}
// VODB_RASTER=0 restores the query-tile-fastest item order (A/B comparisons)
bool raster_enabled() {
  static const char* env = std::getenv("VODB_RASTER");
  return env ? (env[0] != '0') : true;
}
// VODB_RESIDENT=0 keeps the queries streaming through the ring (A/B comparisons)
bool resident_enabled() {
  static const char* env = std::getenv("VODB_RESIDENT");
  return env ? (env[0] != '0') : true;
}
bool use_pair_kernel(const SegmentArgs& a) { return pair_kernel_enabled() && a.terms == 1 && a.planes == 1 && a.nq > 128; }

// corpus tensor map (cached in the store): the rows themselves (bf16 / fp16 store), or the bf16 planes of an fp32
// store stacked along the rows ([3 * n_rows, pitch]; plane p of row r is row p * n_rows + r)
int corpus_tensor_map(vodb_store* s, const CUtensorMap** out) {
  const bool planes = (s->dtype == VODB_F32);
  unsigned char* storage = planes ? s->tmap_planes : s->tmap_corpus;
  bool& valid = planes ? s->tmap_planes_valid : s->tmap_corpus_valid;
  if (!valid) {
    int rc = planes ? encode_2d(reinterpret_cast<CUtensorMap*>(storage), s->planes, VODB_BF16, 3 * s->n_rows, s->pitch, BM)
                    : encode_2d(reinterpret_cast<CUtensorMap*>(storage), s->data, s->dtype, s->n_rows, s->pitch, BM);
    if (rc != VODB_OK) return rc;
    valid = true;
  }
  *out = reinterpret_cast<const CUtensorMap*>(storage);
  return VODB_OK;
}

template <int BN, int T, int P = 1>
int launch_bn(vodb_store* s, const SegmentArgs& a, cudaStream_t stream) {
  using Cfg = TcConfig<BN, T, P>;
  static_assert(Cfg::kStages >= 2, "pipeline needs at least two stages");
  VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&score_tc_kernel<BN, T, P>), Cfg::kSmemBytes));
  const CUtensorMap* tmap_store = nullptr;
  int rc = corpus_tensor_map(s, &tmap_store);
  if (rc != VODB_OK) return rc;
  alignas(64) CUtensorMap tmap_q;
  // the staged query buffer is zero padded to a multiple of 256 rows (api.cu), so every BN-row box is in bounds
  const int64_t q_rows_pad = ((int64_t)a.nq + 255) / 256 * 256;
  rc = encode_2d(&tmap_q, a.queries, a.dtype, q_rows_pad * T, s->pitch, BN);
  if (rc != VODB_OK) return rc;
  // the last query tile's own box: as tall as its MMA is wide (multiple of 32 rows)
  alignas(64) CUtensorMap tmap_q_last;
  const int last_cols = std::min(BN, ((a.nq - ((a.nq - 1) / BN) * BN) + 31) / 32 * 32);
  rc = encode_2d(&tmap_q_last, a.queries, a.dtype, q_rows_pad * T, s->pitch, last_cols);
  if (rc != VODB_OK) return rc;

  TcParams p;
  p.last_box_bytes = (uint32_t)last_cols * KC * 2;
  p.n_stages = 0;
  p.row_begin = a.row_begin;
  p.row_end = a.row_end;
  p.nq = a.nq;
  p.n_ctiles = (int)((a.row_end - a.row_begin + BM - 1) / BM);
  p.n_qtiles = (a.nq + BN - 1) / BN;
  p.kchunks = (s->dim + KC - 1) / KC;  // the pitch may carry one more, unscanned chunk (api.cu vodb_store_create)
  p.q_rows_pad = (int)q_rows_pad;
  p.plane_rows = (int)s->n_rows;
  p.cand_s = a.cand_s;
  p.cand_i = a.cand_i;
  p.cnt = a.cnt;
  p.tau = a.tau;
  p.overflow = a.overflow;
  p.cap = a.cap;
  p.dump = a.dump ? 1 : 0;
  p.term_any = a.term_any;
  p.term_blocks = a.term_blocks;
  p.term_policy = kTermAlways;
  const uint32_t fmt = (a.dtype == VODB_BF16) ? 1u : 0u;  // UMMA F16F32Format: F16=0, BF16=1
  p.idesc = (1u << 4) /*D=f32*/ | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  int64_t items = (int64_t)p.n_ctiles * p.n_qtiles;
  if (items <= 0) return VODB_OK;
  int grid = (int)(items < s->sm_count ? items : s->sm_count);
  p.raster_tiles = (p.n_qtiles > 1 && raster_enabled()) ? grid : 0;
  VODB_CUDA_CHECK(launch_pdl(score_tc_kernel<BN, T, P>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream,
                             *tmap_store, tmap_q, tmap_q_last, p));
  return VODB_OK;
}

template <int BN, int T, bool RES = false>
int launch_pair(vodb_store* s, const SegmentArgs& a, int term_policy, cudaStream_t stream, int res_stages = 0) {
  using Cfg = Tc2Config<BN, T>;
  const int kchunks_all = (s->dim + KC - 1) / KC;
  const size_t smem_bytes = RES ? (size_t)kchunks_all * T * Cfg::kBBytes + (size_t)res_stages * Cfg::kABytes + Cfg::kExtra
                                : (size_t)Cfg::kSmemBytes;
  VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&score_tc2_kernel<BN, T, RES>), smem_bytes));
  const CUtensorMap* tmap_store = nullptr;
  int rc = corpus_tensor_map(s, &tmap_store);
  if (rc != VODB_OK) return rc;
  alignas(64) CUtensorMap tmap_q;
  const int64_t q_rows_pad = ((int64_t)a.nq + 255) / 256 * 256;
  rc = encode_2d(&tmap_q, a.queries, a.dtype, q_rows_pad * T, s->pitch, BN / 2);  // box = one CTA's half of a query tile
  if (rc != VODB_OK) return rc;
  alignas(64) CUtensorMap tmap_q_last;  // the last query tile's own box: half of its MMA width
  const int last_cols = std::min(BN, ((a.nq - ((a.nq - 1) / BN) * BN) + 31) / 32 * 32);
  rc = encode_2d(&tmap_q_last, a.queries, a.dtype, q_rows_pad * T, s->pitch, last_cols / 2);
  if (rc != VODB_OK) return rc;
  TcParams p;
  p.last_box_bytes = (uint32_t)(last_cols / 2) * KC * 2;
  p.n_stages = res_stages;
  p.row_begin = a.row_begin;
  p.row_end = a.row_end;
  p.nq = a.nq;
  p.n_ctiles = (int)((a.row_end - a.row_begin + 2 * BM - 1) / (2 * BM));  // 256-row pair tiles
  p.n_qtiles = (a.nq + BN - 1) / BN;
  p.kchunks = (s->dim + KC - 1) / KC;  // the pitch may carry one more, unscanned chunk (api.cu vodb_store_create)
  p.q_rows_pad = (int)q_rows_pad;
  p.plane_rows = 0;
  p.cand_s = a.cand_s;
  p.cand_i = a.cand_i;
  p.cnt = a.cnt;
  p.tau = a.tau;
  p.overflow = a.overflow;
  p.cap = a.cap;
  p.dump = a.dump ? 1 : 0;
  p.term_any = a.term_any;
  p.term_blocks = a.term_blocks;
  p.term_policy = term_policy;
  const uint32_t fmt = (a.dtype == VODB_BF16) ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
  int64_t items = (int64_t)p.n_ctiles * p.n_qtiles;
  if (items <= 0) return VODB_OK;
  int pairs = (int)std::min<int64_t>(items, s->sm_count / 2);
  p.raster_tiles = (p.n_qtiles > 1 && raster_enabled()) ? pairs : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  VODB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, score_tc2_kernel<BN, T, RES>, *tmap_store, tmap_q, tmap_q_last, p));
  return VODB_OK;
}

// corpus stages left for the resident-query variant (0 = the CTA's half of the query tile does not leave `min_stages`)
template <int BN, int T = 1>
int resident_stages(const vodb_store* s, int min_stages = 6) {
  using C = Tc2Config<BN, T>;
  const int kchunks = (s->dim + KC - 1) / KC;
  const int st = std::min<int>(C::kMaxStagesRes, ((int)kSmemMax - (int)C::kExtra - kchunks * T * (int)C::kBBytes) / (int)C::kABytes);
  return (resident_enabled() && st >= min_stages) ? st : 0;
}
// VODB_RESIDENT_TERMS=0 streams the query terms of small multi-term batches with the corpus (A/B comparisons)
bool resident_terms_enabled() {
  static const char* env = std::getenv("VODB_RESIDENT_TERMS");
  return env ? (env[0] != '0') : true;
}

// one-term scan of up to 128 queries: resident-query pair kernel when it fits, else the 1-CTA kernel
template <int BN>
int launch_one_term_small(vodb_store* s, const SegmentArgs& a, int term_policy, cudaStream_t stream) {
  const int st = resident_stages<BN>(s);
  if (st > 0) return launch_pair<BN, 1, true>(s, a, term_policy, stream, st);
  return launch_pair<BN, 1>(s, a, term_policy, stream);
}

// multi-term scan of a 16-bit store on CTA pairs. Batches above 128 queries are launched twice — the wide one-term
// kernel, which works only if the correction terms turn out empty on the device (float32 queries that are exact in
// the store dtype), and the multi-term kernel, which works only if they do not: exactly one of them finds items.
// (Up to 128 queries the multi-term kernel skips the empty terms itself: a second launch per segment costs as much as
// the resident one-term kernel would gain, measured. Host-resident queries are classified on the host instead,
// api.cu `queries_fit_store_dtype`.)
template <int T>
int launch_pair_terms(vodb_store* s, const SegmentArgs& a, cudaStream_t stream) {
  if (a.nq <= 64) {
    // all terms of the 64 queries resident (144 KB at three terms x 768 dims) when that leaves four corpus stages
    const int st = resident_terms_enabled() ? resident_stages<64, T>(s, 4) : 0;
    if (st > 0) return launch_pair<64, T, true>(s, a, kTermAlways, stream, st);
    return launch_pair<64, T>(s, a, kTermAlways, stream);
  }
  if (a.nq <= 128) return launch_pair<128, T>(s, a, kTermAlways, stream);
  int rc = launch_pair<256, 1>(s, a, kOnlyIfSingleTerm, stream);
  if (rc != VODB_OK) return rc;
  return launch_pair<128, T>(s, a, kOnlyIfMultiTerm, stream);
}

}  // namespace

bool tensor_path_supported(const vodb_store* s) {
  // bf16 / fp16 stores natively; fp32 stores through their bf16 planes (api.cu ensure_planes)
  return get_encode_fn() != nullptr;
}

int launch_score_tensor(vodb_store* s, const SegmentArgs& a, cudaStream_t stream) {
  if (a.row_begin % BM != 0) {
    set_error("launch_score_tensor: segment start %lld is not a multiple of %d", (long long)a.row_begin, BM);
    return VODB_EINVAL;
  }
  if ((int64_t)s->n_rows * (a.planes > 1 ? 3 : 1) > 0x7fffffffLL) {
    set_error("launch_score_tensor: shard too large for 32-bit TMA coordinates");
    return VODB_EUNSUPPORTED;
  }
  if (a.planes > 1) {  // fp32 store: P bf16 corpus planes x P query terms
    if (a.planes != a.terms) {
      set_error("launch_score_tensor: %d corpus planes need as many query terms (got %d)", a.planes, a.terms);
      return VODB_EINVAL;
    }
    if (a.planes == 2) return a.nq <= 64 ? launch_bn<64, 2, 2>(s, a, stream) : launch_bn<128, 2, 2>(s, a, stream);
    return a.nq <= 64 ? launch_bn<64, 3, 3>(s, a, stream) : launch_bn<128, 3, 3>(s, a, stream);
  }
  if (a.terms == 1 && a.nq <= 128 && pair_kernel_enabled()) {
    // resident-query pair kernel when the CTA's half of the query tile (all K chunks) leaves >= 6 corpus stages
    if (a.nq <= 64 && resident_stages<64>(s) > 0) return launch_one_term_small<64>(s, a, kTermAlways, stream);
    if (a.nq > 64 && resident_stages<128>(s) > 0) return launch_one_term_small<128>(s, a, kTermAlways, stream);
  }
  if (use_pair_kernel(a)) return launch_pair<256, 1>(s, a, kTermAlways, stream);
  if (pair_kernel_enabled() && a.terms == 2) return launch_pair_terms<2>(s, a, stream);
  if (pair_kernel_enabled() && a.terms == 3) return launch_pair_terms<3>(s, a, stream);
  switch (a.terms) {  // 1-CTA kernels: one term up to 128 queries; everything when VODB_TC2=0
    case 1:
      if (a.nq <= 64) return launch_bn<64, 1>(s, a, stream);
      if (a.nq <= 128) return launch_bn<128, 1>(s, a, stream);
      return launch_bn<256, 1>(s, a, stream);
    case 2:
      if (a.nq <= 64) return launch_bn<64, 2>(s, a, stream);
      return launch_bn<128, 2>(s, a, stream);
    case 3:
      if (a.nq <= 64) return launch_bn<64, 3>(s, a, stream);
      return launch_bn<128, 3>(s, a, stream);
  }
  set_error("launch_score_tensor: bad number of query terms %d", a.terms);
  return VODB_EINVAL;
}

}  // namespace vodb
