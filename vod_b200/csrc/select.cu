// select.cu — per-query exact top-k selection over candidate lists (thread-maximum bound + rank sort, radix select +
// bitonic sort as the general path).
//
// Replaces the k-selection half of faiss' IndexFlat::search (heap / reservoir collection behind
// `faiss_index.search`, reference src/vod_search/faiss_search/server.py:84) and, as `merge`, the
// host-side IndexShards merge behind `faiss.index_cpu_to_all_gpus(..., co.shard=True)` (server.py:51-54).
//
// One CTA per query. The list is a bag of (score, id) pairs appended by the scoring kernels. Fast path: see
// block_select (k <= threads/2). General path:
//   1. 4-pass (8 bits each) MSB radix select on the order-preserving uint32 image of the score finds
//      v* = the k-th largest score and how many entries tied at v* are needed;
//   2. if the tie group is larger than needed, a second radix select picks the smallest ids;
//   3. the selected entries are compacted into shared memory; when `final`, a bitonic network sorts
//      them by (score desc, id asc) and writes scores / global ids, else they are written back to the
//      front of the list and tau[q] := v* becomes the filter threshold of the next scan segment.
// Exact (no approximation): the order is total because ids are unique within a list.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vodb {

namespace {

constexpr int kSelThreads = 256;

template <typename IdxT>
struct UIdx;
template <>
struct UIdx<int32_t> { using type = uint32_t; };
template <>
struct UIdx<int64_t> { using type = uint64_t; };

template <typename IdxT>
struct SelectSmem {
  int hist[512];     // radix histogram; the thread-maximum bound uses both halves (one per pass)
  int bin;
  int need;
  int n_eq;
  int sel_count;
  int dup_taken;
  uint32_t min_ord;
  uint32_t max_ord;  // first pass: largest key of the list
  int n_top;         // first pass: entries whose top byte equals the top byte of max_ord
};

// (o desc, uidx asc) "a before b"
template <typename U>
__device__ __forceinline__ bool before(uint32_t oa, U ia, uint32_t ob, U ib) {
  return (oa > ob) || (oa == ob && ia < ib);
}

// One warp locates the histogram bin holding the `need`-th entry, scanning bins from the top (kDescending) or the
// bottom: lane l owns 8 consecutive bins in scan order, a warp prefix sum over the lane totals finds the lane, the
// lane walks its 8 bins. Writes sm.bin, sm.need (rank inside the bin, 1-based) and sm.n_eq (size of the bin).
template <bool kDescending, typename IdxT>
__device__ __forceinline__ void find_bin(const int* hist, SelectSmem<IdxT>& sm, int need, int lane) {
  int h[8];
  int total = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int pos = lane * 8 + j;  // position in scan order
    const int bin = kDescending ? 255 - pos : pos;
    h[j] = hist[bin];
    total += h[j];
  }
  int incl = total;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  const int excl = incl - total;
  // the first lane whose inclusive count reaches `need` owns the bin (the last lane takes it if none does)
  const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
  const int owner = hit ? (__ffs(hit) - 1) : 31;
  if (lane == owner) {
    int cum = excl, j = 0;
    for (; j < 7; ++j) {
      if (cum + h[j] >= need) break;
      cum += h[j];
    }
    const int pos = lane * 8 + j;
    sm.bin = kDescending ? 255 - pos : pos;
    sm.need = need - cum;
    sm.n_eq = h[j];
  }
}

// The same search done by EVERY warp for itself (descending), result in registers: no shared-memory hand-over and no
// barrier after it — the thread-maximum bound runs it twice on a 32-warp CTA whose other warps would only wait.
__device__ __forceinline__ void find_bin_every_warp(const int* hist, int need, int lane, int& bin_out, int& need_out) {
  // scan position p = j * 32 + lane (bin 255 - p): eight conflict-free loads; block j = 32 consecutive positions
  int v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = hist[255 - (j * 32 + lane)];
  int cum = 0, jb = 7, before = 0;
  bool found = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = __reduce_add_sync(0xffffffffu, v[j]);
    if (!found && (cum + t >= need || j == 7)) {
      found = true;
      jb = j;
      before = cum;
    }
    cum += t;
  }
  int x = v[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) x = (j == jb) ? v[j] : x;
  int incl = x;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += u;
  }
  const unsigned hit = __ballot_sync(0xffffffffu, before + incl >= need);
  const int owner = hit ? (__ffs(hit) - 1) : 31;
  bin_out = 255 - (jb * 32 + owner);
  need_out = need - (before + __shfl_sync(0xffffffffu, incl - x, owner));
}

// Rank sort of ns <= blockDim.x entries: entry i goes to position #{j : j before i} in (score desc, id asc) order; equal
// (score, id) pairs exist only as padding entries of the merges and are ordered by position, so the ranks are a
// permutation. g consecutive lanes (g | 32) share one entry's ns comparisons. Entries ranked below k_out are dropped.
// One pass of broadcast shared-memory reads and one or two barriers, against log2(P)*(log2(P)+1)/2 barrier-separated
// stages of a bitonic network — but ns^2 comparisons: the callers use it up to kRankSortMax entries.
constexpr int kRankSortMax = 384;
template <typename IdxT>
__device__ __forceinline__ void rank_sort(const uint32_t* src_o, const IdxT* src_i, int ns, int k_out, uint32_t* dst_o,
                                          IdxT* dst_i, bool in_place) {
  using U = typename UIdx<IdxT>::type;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int g = (ns * 8 <= nt) ? 8 : (ns * 4 <= nt) ? 4 : (ns * 2 <= nt) ? 2 : 1;
  const int sv = tid / g, part = tid - sv * g;
  const bool live = sv < ns;
  const uint32_t o = live ? src_o[sv] : 0u;
  const U id = live ? (U)src_i[sv] : (U)0;
  int rank = 0;
  if (live) {
#pragma unroll 4
    for (int j = part; j < ns; j += g) {
      const uint32_t oj = src_o[j];
      const U ij = (U)src_i[j];
      rank += (before<U>(oj, ij, o, id) || (oj == o && ij == id && j < sv)) ? 1 : 0;
    }
  }
  for (int off = g >> 1; off > 0; off >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, off);
  if (in_place) __syncthreads();
  if (live && part == 0 && rank < k_out) {
    dst_o[rank] = o;
    dst_i[rank] = (IdxT)id;
  }
  __syncthreads();
}

// Core routine. load_s(i) / load_i(i) read entry i in [0,n). Results: sel_o/sel_i[0..n_sel) in shared memory
// (sorted if do_sort), returns n_sel = min(n,k); *vstar_out = ord image of the k-th best score (0 if n < k).
// `cache` (optional, capacity cache_n >= n required to be used) keeps the ordered score images in shared memory
// after the first pass, so that the remaining radix passes and the compaction never go back to L2.
//
// Thread-maximum bound (surv_o / surv_i given, 2k <= threads, keys cheap to re-read: cached or `smem_src`): every
// thread keeps the maximum of the keys it read in the first sweep. The k-th largest of those `threads` maxima is a
// lower bound of the k-th best key of the list (k different entries reach it), and a tight one: for a list in random
// order about -threads * ln(1 - k/threads) entries pass it whatever the list length (115 of a 16384-entry list at
// k = 100 with 1024 threads). The bound is found with two 8-bit radix passes over ONE value per thread (its low 16 bits
// are left zero: ~5% more survivors), the survivors are compacted and rank-sorted, and the first k are the result —
// six block barriers and two sweeps instead of the ~20 barriers and six sweeps of the radix select below, which stays
// as the general path (k > threads/2, or an ordering of the list that leaves more than `threads` survivors).
template <typename IdxT, typename LoadS, typename LoadI>
__device__ int block_select(LoadS load_s, LoadI load_i, int n, int k, bool do_sort, SelectSmem<IdxT>& sm,
                            uint32_t* sel_o, IdxT* sel_i, int P, uint32_t* cache, int cache_n, uint32_t* vstar_out,
                            const float* flat_s = nullptr, uint32_t* surv_o = nullptr, IdxT* surv_i = nullptr,
                            bool smem_src = false) {
  using U = typename UIdx<IdxT>::type;
  const int tid = threadIdx.x;
  const int nt = blockDim.x;
  int n_sel;
  uint32_t vstar = 0u;
  bool sorted = false;  // the selection below already left sel_o / sel_i in output order
  const bool cached = cache != nullptr && n <= cache_n;
  auto key = [&](int i) -> uint32_t { return cached ? cache[i] : ord_u32(load_s(i)); };

  if (n <= k) {
    if (tid == 0) sm.min_ord = 0xffffffffu;
    __syncthreads();
    uint32_t local_min = 0xffffffffu;
    for (int i = tid; i < n; i += nt) {
      uint32_t o = ord_u32(load_s(i));
      sel_o[i] = o;
      sel_i[i] = load_i(i);
      local_min = min(local_min, o);
    }
    atomicMin(&sm.min_ord, local_min);
    __syncthreads();
    n_sel = n;
    vstar = (n == k && n > 0) ? sm.min_ord : 0u;
  } else {
    // First sweep: read the scores (4 independent loads in flight per thread), cache the ordered images, keep the
    // thread's maximum.
    uint32_t lmax = 0u;
    {
      int i = tid;
      if (flat_s != nullptr && cached && ((reinterpret_cast<uintptr_t>(flat_s) | reinterpret_cast<uintptr_t>(cache)) & 15) == 0) {
        // contiguous list (16-byte aligned): four 16-byte loads in flight per thread, 16 scores each round — the
        // pass is latency bound (64 CTAs on 148 SMs), so memory-level parallelism is what shortens it
        const int n4 = n >> 2;
        const float4* src4 = reinterpret_cast<const float4*>(flat_s);
        uint4* dst4 = reinterpret_cast<uint4*>(cache);
        int v = tid;
        for (; v + 3 * nt < n4; v += 4 * nt) {
          const float4 a = src4[v], b = src4[v + nt], c = src4[v + 2 * nt], d = src4[v + 3 * nt];
          const uint4 oa = make_uint4(ord_u32(a.x), ord_u32(a.y), ord_u32(a.z), ord_u32(a.w));
          const uint4 ob = make_uint4(ord_u32(b.x), ord_u32(b.y), ord_u32(b.z), ord_u32(b.w));
          const uint4 oc = make_uint4(ord_u32(c.x), ord_u32(c.y), ord_u32(c.z), ord_u32(c.w));
          const uint4 od = make_uint4(ord_u32(d.x), ord_u32(d.y), ord_u32(d.z), ord_u32(d.w));
          dst4[v] = oa; dst4[v + nt] = ob; dst4[v + 2 * nt] = oc; dst4[v + 3 * nt] = od;
          lmax = max(lmax, max(max(max(oa.x, oa.y), max(oa.z, oa.w)), max(max(ob.x, ob.y), max(ob.z, ob.w))));
          lmax = max(lmax, max(max(max(oc.x, oc.y), max(oc.z, oc.w)), max(max(od.x, od.y), max(od.z, od.w))));
        }
        for (; v < n4; v += nt) {
          const float4 a = src4[v];
          const uint4 oa = make_uint4(ord_u32(a.x), ord_u32(a.y), ord_u32(a.z), ord_u32(a.w));
          dst4[v] = oa;
          lmax = max(lmax, max(max(oa.x, oa.y), max(oa.z, oa.w)));
        }
        i = (n4 << 2) + tid;  // the last n % 4 entries go through the scalar tail below
      }
      for (; i + 3 * nt < n; i += 4 * nt) {
        float s0 = load_s(i), s1 = load_s(i + nt), s2 = load_s(i + 2 * nt), s3 = load_s(i + 3 * nt);
        uint32_t o0 = ord_u32(s0), o1 = ord_u32(s1), o2 = ord_u32(s2), o3 = ord_u32(s3);
        if (cached) { cache[i] = o0; cache[i + nt] = o1; cache[i + 2 * nt] = o2; cache[i + 3 * nt] = o3; }
        lmax = max(max(lmax, o0), max(o1, max(o2, o3)));
      }
      for (; i < n; i += nt) {
        uint32_t o = ord_u32(load_s(i));
        if (cached) cache[i] = o;
        lmax = max(lmax, o);
      }
    }
    if (surv_o != nullptr && (cached || smem_src) && 2 * k <= nt) {
      for (int t = tid; t < 512; t += nt) sm.hist[t] = 0;
      if (tid == 0) sm.sel_count = 0;
      __syncthreads();  // also publishes the cached keys
      const int lane = tid & 31;
      {  // pass A: top byte of the thread maxima (equal bins of a warp share one atomic)
        const uint32_t b = lmax >> 24;
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        if (lane == __ffs(peers) - 1) atomicAdd(&sm.hist[b], __popc(peers));
      }
      __syncthreads();
      int bin_a_i, need_b;
      find_bin_every_warp(sm.hist, k, lane, bin_a_i, need_b);
      const uint32_t bin_a = (uint32_t)bin_a_i;
      // pass B: second byte, among the maxima inside bin A (the values differ from lane to lane here: plain atomics —
      // a match over 32 distinct values costs more than the 32 conflict-free atomics it would save)
      if ((lmax >> 24) == bin_a) atomicAdd(&sm.hist[256 + ((lmax >> 16) & 255u)], 1);
      __syncthreads();
      int bin_b, need_c;
      find_bin_every_warp(sm.hist + 256, need_b, lane, bin_b, need_c);
      const uint32_t bound = (bin_a << 24) | ((uint32_t)bin_b << 16);
      // compaction: key and list POSITION of every survivor (the id is fetched afterwards, one load per survivor and
      // all of them in flight together: loading it here would serialise an L2 round trip per survivor of a thread)
      for (int i0 = tid; i0 < n; i0 += 4 * nt) {  // four keys in flight, survivors are rare
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = (i0 + u * nt < n) ? key(i0 + u * nt) : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (o[u] >= bound && i0 + u * nt < n) {
            const int pos = atomicAdd(&sm.sel_count, 1);
            if (pos < nt) {
              surv_o[pos] = o[u];
              surv_i[pos] = (IdxT)(i0 + u * nt);
            }
          }
        }
      }
      __syncthreads();
      const int ns = sm.sel_count;  // >= k: the k largest thread maxima are different entries
      if (ns <= nt) {
        // The survivors' ids are fetched now (one load each, own slot) and needed only to order equal scores: the
        // ranking below runs on the keys alone while the loads are in flight, and is final unless two survivors tie.
        IdxT my_id = 0;
        if (tid < ns) my_id = load_i((int)surv_i[tid]);
        const int g = (ns * 8 <= nt) ? 8 : (ns * 4 <= nt) ? 4 : (ns * 2 <= nt) ? 2 : 1;
        const int sv = tid / g, part = tid - sv * g;
        const bool live = sv < ns;
        const uint32_t o = live ? surv_o[sv] : 0u;
        int rank = 0, tie = 0;
        if (live) {
#pragma unroll 4
          for (int j = part; j < ns; j += g) {
            const uint32_t oj = surv_o[j];
            rank += (oj > o || (oj == o && j < sv)) ? 1 : 0;
            tie |= (oj == o && j != sv) ? 1 : 0;
          }
        }
        for (int off = g >> 1; off > 0; off >>= 1) {
          rank += __shfl_xor_sync(0xffffffffu, rank, off);
          tie |= __shfl_xor_sync(0xffffffffu, tie, off);
        }
        if (tid < ns) surv_i[tid] = my_id;
        if (__syncthreads_or(tie) == 0) {
          if (live && part == 0 && rank < k) {
            sel_o[rank] = o;
            sel_i[rank] = surv_i[sv];
          }
          __syncthreads();
        } else {
          rank_sort<IdxT>(surv_o, surv_i, ns, k, sel_o, sel_i, false);  // (score desc, id asc) with the ids
        }
        sorted = true;
        n_sel = k;
        vstar = sel_o[k - 1];
      }
    }
    if (!sorted) {
      uint32_t prefix = 0u, mask = 0u;
      int need = k;
      for (int shift = 24; shift >= 0; shift -= 8) {
        for (int t = tid; t < 256; t += nt) sm.hist[t] = 0;
        __syncthreads();
        if (shift == 24) {
          // First radix pass. A histogram of the top byte (sign + 7 exponent bits) would send most of the list to three
          // or four bins, and shared-memory atomics on one address cost a cycle per lane: ~8 us for a 16k-entry dump
          // list. Scores of one query nearly always have >= k entries in the top byte of their maximum, so the pass
          // first tries exactly that bin with ballots instead of atomics: reduce the maximum, count the entries that
          // share its top byte. If there are at least `need` of them the k-th best lies in that bin and the pass is
          // done; otherwise (top bin too small, e.g. one outlier score) the general histogram below runs.
          if (tid == 0) { sm.max_ord = 0u; sm.n_top = 0; }
          __syncthreads();
          lmax = __reduce_max_sync(0xffffffffu, lmax);
          if ((tid & 31) == 0) atomicMax(&sm.max_ord, lmax);
          __syncthreads();
          const uint32_t top = sm.max_ord >> 24;
          int c = 0;
          for (int j = tid; j < n; j += nt) c += (key(j) >> 24) == top ? 1 : 0;
          c = __reduce_add_sync(0xffffffffu, c);
          if ((tid & 31) == 0 && c) atomicAdd(&sm.n_top, c);
          __syncthreads();
          if (sm.n_top >= need) {
            if (tid == 0) { sm.bin = (int)top; sm.need = need; sm.n_eq = sm.n_top; }
            __syncthreads();
            prefix |= top << 24;
            mask |= 0xffu << 24;
            __syncthreads();
            continue;
          }
          for (int j = tid; j < n; j += nt) atomicAdd(&sm.hist[key(j) >> 24], 1);
        } else {
          for (int i = tid; i < n; i += nt) {
            uint32_t o = key(i);
            if ((o & mask) == prefix) atomicAdd(&sm.hist[(o >> shift) & 255u], 1);
          }
        }
        __syncthreads();
        if (tid < 32) find_bin<true>(sm.hist, sm, need, tid);
        __syncthreads();
        prefix |= (uint32_t)sm.bin << shift;
        mask |= 0xffu << shift;
        need = sm.need;
        __syncthreads();
      }
      vstar = prefix;
      const int n_eq = sm.n_eq;
      // ties at v*: take the `need` smallest ids
      U istar = ~(U)0;
      int ineed = 0x7fffffff;  // how many entries with (o==v*, id==istar) to take
      if (n_eq > need) {
        U iprefix = 0, imask = 0;
        ineed = need;
        for (int shift = (int)sizeof(U) * 8 - 8; shift >= 0; shift -= 8) {
          for (int t = tid; t < 256; t += nt) sm.hist[t] = 0;
          __syncthreads();
          for (int i = tid; i < n; i += nt) {
            if (key(i) == vstar) {
              U u = (U)load_i(i);
              if ((u & imask) == iprefix) atomicAdd(&sm.hist[(int)((u >> shift) & 255u)], 1);
            }
          }
          __syncthreads();
          if (tid < 32) find_bin<false>(sm.hist, sm, ineed, tid);
          __syncthreads();
          iprefix |= (U)sm.bin << shift;
          imask |= (U)0xff << shift;
          ineed = sm.need;
          __syncthreads();
        }
        istar = iprefix;
      }
      if (tid == 0) {
        sm.sel_count = 0;
        sm.dup_taken = 0;
      }
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        uint32_t o = key(i);
        if (o < vstar) continue;
        IdxT id = load_i(i);
        bool take = o > vstar;
        if (!take) {
          U u = (U)id;
          if (u < istar) take = true;
          else if (u == istar) take = atomicAdd(&sm.dup_taken, 1) < ineed;
        }
        if (take) {
          int pos = atomicAdd(&sm.sel_count, 1);
          if (pos < P) {
            sel_o[pos] = o;
            sel_i[pos] = id;
          }
        }
      }
      __syncthreads();
      n_sel = min(sm.sel_count, k);
    }  // !sorted
  }

  if (sorted) {
    // nothing to do
  } else if (do_sort && n_sel <= nt && n_sel <= kRankSortMax) {
    rank_sort<IdxT>(sel_o, sel_i, n_sel, n_sel, sel_o, sel_i, true);
  } else if (do_sort) {
    for (int i = n_sel + tid; i < P; i += nt) {  // padding sorts last
      sel_o[i] = 0u;
      sel_i[i] = (IdxT)(~(U)0 >> 1);
    }
    for (int size = 2; size <= P; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        __syncthreads();
        for (int t = tid; t < (P >> 1); t += nt) {
          int lo = 2 * t - (t & (stride - 1));
          int hi = lo + stride;
          bool first_half = (lo & size) == 0;  // ascending ("before" order) in first half of each bitonic block
          uint32_t oa = sel_o[lo], ob = sel_o[hi];
          U ia = (U)sel_i[lo], ib = (U)sel_i[hi];
          bool a_before_b = before<U>(oa, ia, ob, ib);
          if (a_before_b != first_half) {
            sel_o[lo] = ob; sel_o[hi] = oa;
            sel_i[lo] = (IdxT)ib; sel_i[hi] = (IdxT)ia;
          }
        }
      }
    }
    __syncthreads();
  }
  *vstar_out = vstar;
  return n_sel;
}

__host__ __device__ inline int pow2ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// dynamic smem layout: sel_i[P] | sel_o[P] | cache[cache_n] | (fast) surv_o[threads] | surv_i[threads]
template <typename IdxT>
__global__ void __launch_bounds__(1024)
select_kernel(float* __restrict__ cand_s, int32_t* __restrict__ cand_i, int* __restrict__ cnt, float* __restrict__ tau,
              int cap, int k, int P, int final_pass, float* __restrict__ out_s, int64_t* __restrict__ out_i,
              int64_t row_offset, const ExchangeDst xd, int use_xd, int cache_n, int fast) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ SelectSmem<int32_t> sm;
  int32_t* sel_i = reinterpret_cast<int32_t*>(dyn);
  uint32_t* sel_o = reinterpret_cast<uint32_t*>(dyn + (size_t)P * sizeof(int32_t));
  uint32_t* cache = cache_n > 0 ? sel_o + P : nullptr;
  uint32_t* surv_o = fast ? sel_o + P + cache_n : nullptr;  // survivors of the thread-maximum bound (block_select)
  int32_t* surv_i = fast ? reinterpret_cast<int32_t*>(surv_o + blockDim.x) : nullptr;
  const int q = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();  // the scoring kernel's appends must be complete and visible
  const int n = min(cnt[(size_t)q * kCntStride], cap);
  float* ls = cand_s + (size_t)q * cap;
  int32_t* li = cand_i + (size_t)q * cap;
  auto load_s = [&](int i) -> float { return ls[i]; };
  auto load_i = [&](int i) -> int32_t { return li[i]; };
  uint32_t vstar;
  // lists start at multiples of cap (>= 2048) floats: 16-byte aligned for the vectorised first pass
  int n_sel = block_select<int32_t>(load_s, load_i, n, k, final_pass != 0, sm, sel_o, sel_i, P, cache, cache_n, &vstar, ls,
                                    surv_o, surv_i);
  __syncthreads();
  if (final_pass && use_xd) {
    // fused exchange: store this shard's result into every rank's gather buffer (own rank included) as
    // tagged 8-byte words; nothing else is needed to publish it
    const uint64_t tag = (uint64_t)xd.epoch << 32;
    if (q == 0 && threadIdx.x == 0) {
      // every scoring kernel of this search has completed (stream order): the shard's overflow flag is final
      // payload: bit 0 = this shard's overflow flag, bits 1.. = fingerprint of the batch shape (nq, k): ranks that call
      // with different shapes would read each other's slots at the wrong offsets without anyone noticing
      const uint32_t shape = ((uint32_t)gridDim.x * 2654435761u) ^ ((uint32_t)k * 40503u);
      const uint64_t f = tag | (uint64_t)(((shape & 0x7fffffffu) << 1) |
                                          (*reinterpret_cast<const volatile int*>(xd.overflow) != 0 ? 1u : 0u));
      for (int r = 0; r < xd.world; ++r)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(xd.peer_flag[r]), "l"(f) : "memory");
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const float sv = (j < n_sel) ? ord_to_float(sel_o[j]) : VODB_NEG_FLT_MAX;
      const int64_t iv = (j < n_sel) ? (int64_t)sel_i[j] + row_offset : (int64_t)-1;
      const uint64_t w0 = tag | (uint64_t)__float_as_uint(sv);
      const uint64_t w1 = tag | (uint64_t)(uint32_t)iv;
      const uint64_t w2 = tag | (uint64_t)(uint32_t)((uint64_t)iv >> 32);
      for (int r = 0; r < xd.world; ++r) {
        uint64_t* dst = xd.peer_ll[r] + (size_t)q * k + j;  // consecutive threads -> consecutive words of a plane
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + xd.slot_elems), "l"(w1) : "memory");
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 2 * xd.slot_elems), "l"(w2) : "memory");
      }
    }
  } else if (final_pass) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      if (j < n_sel) {
        out_s[(size_t)q * k + j] = ord_to_float(sel_o[j]);
        out_i[(size_t)q * k + j] = (int64_t)sel_i[j] + row_offset;
      } else {
        out_s[(size_t)q * k + j] = VODB_NEG_FLT_MAX;
        out_i[(size_t)q * k + j] = -1;
      }
    }
  } else {
    if (n > k) {  // compact the survivors to the front of the list
      for (int j = threadIdx.x; j < n_sel; j += blockDim.x) {
        ls[j] = ord_to_float(sel_o[j]);
        li[j] = sel_i[j];
      }
      if (threadIdx.x == 0) cnt[(size_t)q * kCntStride] = n_sel;
    }
    if (threadIdx.x == 0) tau[q] = (n >= k) ? ord_to_float(vstar) : -INFINITY;
  }
}

// Both merge kernels first copy the world*k_in (score, id) entries of their query into shared memory — one pass over
// the gathered lists (for the fused exchange: one tagged sys-scope load per word, spinning until the peer's store has
// landed) — and run the radix select + sort there. `staged` == 0 is the fallback for inputs that do not fit
// (world*k_in*12 bytes > ~190 KB): every pass then re-reads global memory as the first version did.
__global__ void __launch_bounds__(1024)
merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int n_lists, int nq, int k_in,
             int k_out, int P, int staged, int fast, float* __restrict__ out_s, int64_t* __restrict__ out_i) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ SelectSmem<int64_t> sm;
  int64_t* sel_i = reinterpret_cast<int64_t*>(dyn);
  uint32_t* sel_o = reinterpret_cast<uint32_t*>(dyn + (size_t)P * sizeof(int64_t));
  const int q = blockIdx.x;
  const int n = n_lists * k_in;
  auto off_of = [&](int i) -> size_t {
    int l = i / k_in, j = i - l * k_in;
    return ((size_t)l * nq + q) * k_in + j;
  };
  uint32_t vstar;
  int n_sel;
  if (staged) {
    int64_t* all_i = reinterpret_cast<int64_t*>(dyn + (((size_t)P * 12 + 15) & ~(size_t)15));
    float* all_s = reinterpret_cast<float*>(all_i + n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const size_t o = off_of(i);
      all_s[i] = scores[o];
      all_i[i] = idx[o];
    }
    __syncthreads();
    auto load_s = [&](int i) -> float { return all_s[i]; };
    auto load_i = [&](int i) -> int64_t { return all_i[i]; };
    int64_t* surv_i = fast ? all_i + n + (n + 1) / 2 : nullptr;  // behind all_s, 8-byte aligned
    uint32_t* surv_o = fast ? reinterpret_cast<uint32_t*>(surv_i + blockDim.x) : nullptr;
    n_sel = block_select<int64_t>(load_s, load_i, n, k_out, true, sm, sel_o, sel_i, P, nullptr, 0, &vstar, nullptr, surv_o,
                                  surv_i, true);
  } else {
    auto load_s = [&](int i) -> float { return scores[off_of(i)]; };
    auto load_i = [&](int i) -> int64_t { return idx[off_of(i)]; };
    n_sel = block_select<int64_t>(load_s, load_i, n, k_out, true, sm, sel_o, sel_i, P, nullptr, 0, &vstar);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
    bool ok = j < n_sel && sel_i[j] >= 0;
    out_s[(size_t)q * k_out + j] = ok ? ord_to_float(sel_o[j]) : VODB_NEG_FLT_MAX;
    out_i[(size_t)q * k_out + j] = ok ? sel_i[j] : -1;
  }
}

// Merge after the fused exchange: every entry is read with a spin on its epoch tag (entries of this rank's own
// select are already there; a peer's arrive as its final select runs), then world*k -> k as in merge_kernel.
__device__ __forceinline__ uint64_t ll_wait_word(const uint64_t* p, uint32_t epoch) {
  uint64_t v;
  unsigned long long spins = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if ((uint32_t)(v >> 32) == epoch) return v;
    if (++spins > (1ull << 31)) {  // a peer never arrived: fail loudly instead of hanging the GPU
      printf("vodb: exchange timeout (epoch %u, word tag %u)\n", epoch, (uint32_t)(v >> 32));
      __trap();
    }
  }
}

__global__ void __launch_bounds__(1024)
merge_exchange_kernel(const uint64_t* __restrict__ gather_ll, uint32_t epoch, int world, size_t slot_words,
                      size_t flag_word, int nq, int k, int P, int staged, int fast, float* __restrict__ out_s,
                      int64_t* __restrict__ out_i, int* __restrict__ overflow_any) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ SelectSmem<int64_t> sm;
  int64_t* sel_i = reinterpret_cast<int64_t*>(dyn);
  uint32_t* sel_o = reinterpret_cast<uint32_t*>(dyn + (size_t)P * sizeof(int64_t));
  pdl_launch_dependents();
  pdl_wait();  // this rank's final select (which also filled slot `rank` of the local gather buffer) is complete
  const int q = blockIdx.x;
  const int n = world * k;
  // every block first looks at the peers' flag words (they arrive with the first block of a peer's final select):
  // bit 0 = that shard overflowed (sticky, read back by the host with the results), the rest = the peer's batch shape.
  // A peer that runs this epoch with another (nq, k) never writes the slots this block would wait for, so the block
  // reports the mismatch and returns padding instead of spinning.
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if (threadIdx.x < world) {
    const uint32_t f = (uint32_t)ll_wait_word(gather_ll + (size_t)threadIdx.x * slot_words + flag_word, epoch);
    const uint32_t shape = (((uint32_t)nq * 2654435761u) ^ ((uint32_t)k * 40503u)) & 0x7fffffffu;
    if (q == 0 && (f & 1u)) atomicOr(overflow_any, 1);
    if ((f >> 1) != shape) {
      s_bad = 1;
      if (q == 0) atomicOr(overflow_any, 2);
    }
  }
  __syncthreads();
  if (s_bad) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      out_s[(size_t)q * k + j] = VODB_NEG_FLT_MAX;
      out_i[(size_t)q * k + j] = -1;
    }
    return;
  }
  const size_t plane = flag_word / 3;  // entries per plane of a slot: [3][plane] tagged words, then the flag word
  auto entry = [&](int i) -> const uint64_t* {
    int l = i / k, j = i - l * k;
    return gather_ll + (size_t)l * slot_words + (size_t)q * k + j;
  };
  uint32_t vstar;
  int n_sel;
  if (staged) {
    int64_t* all_i = reinterpret_cast<int64_t*>(dyn + (((size_t)P * 12 + 15) & ~(size_t)15));
    float* all_s = reinterpret_cast<float*>(all_i + n);
    // the three words of an entry are loaded back to back (independent loads in flight together); only a word whose
    // tag is not there yet is polled
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t* e = entry(i);
      uint64_t w[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[j]) : "l"(e + j * plane) : "memory");
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if ((uint32_t)(w[j] >> 32) != epoch) w[j] = ll_wait_word(e + j * plane, epoch);
      all_s[i] = __uint_as_float((uint32_t)w[0]);
      all_i[i] = (int64_t)((w[1] & 0xffffffffull) | (w[2] << 32));
    }
    __syncthreads();
    auto load_s = [&](int i) -> float { return all_s[i]; };
    auto load_i = [&](int i) -> int64_t { return all_i[i]; };
    int64_t* surv_i = fast ? all_i + n + (n + 1) / 2 : nullptr;  // behind all_s, 8-byte aligned
    uint32_t* surv_o = fast ? reinterpret_cast<uint32_t*>(surv_i + blockDim.x) : nullptr;
    n_sel = block_select<int64_t>(load_s, load_i, n, k, true, sm, sel_o, sel_i, P, nullptr, 0, &vstar, nullptr, surv_o,
                                  surv_i, true);
  } else {
    auto load_s = [&](int i) -> float { return __uint_as_float((uint32_t)ll_wait_word(entry(i), epoch)); };
    auto load_i = [&](int i) -> int64_t {
      const uint64_t* e = entry(i);
      uint64_t lo = ll_wait_word(e + plane, epoch), hi = ll_wait_word(e + 2 * plane, epoch);
      return (int64_t)((lo & 0xffffffffull) | (hi << 32));
    };
    n_sel = block_select<int64_t>(load_s, load_i, n, k, true, sm, sel_o, sel_i, P, nullptr, 0, &vstar);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    bool ok = j < n_sel && sel_i[j] >= 0;
    out_s[(size_t)q * k + j] = ok ? ord_to_float(sel_o[j]) : VODB_NEG_FLT_MAX;
    out_i[(size_t)q * k + j] = ok ? sel_i[j] : -1;
  }
}

// VODB_FAST_SELECT=0 keeps every selection on the radix path (A/B comparisons, tests of the general path)
inline bool fast_select_enabled() {
  static const char* env = std::getenv("VODB_FAST_SELECT");
  return env ? (env[0] != '0') : true;
}

inline int threads_for(int n) {
  static const char* env = std::getenv("VODB_SEL_THREADS");  // development knob (scripts/r02_sweep_select.py)
  if (env) return std::max(64, std::min(1024, std::atoi(env) / 32 * 32));
  return n >= 8192 ? 1024 : n >= 2048 ? 512 : kSelThreads;
}

}  // namespace

int launch_select(float* cand_s, int32_t* cand_i, int* cnt, float* tau, int cap, int nq, int k, bool final_pass,
                  float* out_s, int64_t* out_i, int64_t row_offset, cudaStream_t stream, const ExchangeDst* xd,
                  int expected_n) {
  int P = pow2ceil(k);
  // shared-memory cache of the ordered score images: sized for the list length the host expects (the first segment's
  // row count, or a few multiples of k afterwards); longer lists fall back to re-reading L2 on every pass
  int cache_n = std::min(std::max(expected_n, 0), std::min(cap, 32768));
  cache_n = (cache_n + 255) / 256 * 256;
  // few queries: threads by list length (long lists want memory-level parallelism, short ones cheap barriers);
  // many queries: small CTAs, several per SM
  const int threads = (nq > 512) ? kSelThreads : threads_for(expected_n > 0 ? expected_n : cap);
  const int fast = fast_select_enabled() && 2 * k <= threads ? 1 : 0;  // thread-maximum bound (block_select)
  size_t smem = (size_t)P * (sizeof(int32_t) + sizeof(uint32_t)) + (size_t)cache_n * sizeof(uint32_t) +
                (fast ? (size_t)threads * (sizeof(uint32_t) + sizeof(int32_t)) : 0);
  if (smem > 48 * 1024) VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&select_kernel<int32_t>), 170 * 1024));
  ExchangeDst none{};
  VODB_CUDA_CHECK(launch_pdl(select_kernel<int32_t>, dim3(nq), dim3(threads), smem, stream, cand_s, cand_i, cnt, tau, cap,
                             k, P, final_pass ? 1 : 0, out_s, out_i, row_offset, xd ? *xd : none,
                             (xd && final_pass) ? 1 : 0, cache_n, fast));
  return VODB_OK;
}

// dynamic shared memory of the merge kernels: selection buffers [P] + (if it fits) the staged world*k entries
// (+ the survivor buffers of the thread-maximum bound when it applies: staged input, 2k <= threads)
static size_t merge_smem(int P, int n, int k, int threads, int* staged, int* fast) {
  const size_t sel = ((size_t)P * (sizeof(int64_t) + sizeof(uint32_t)) + 15) & ~(size_t)15;
  const size_t all = (size_t)n * sizeof(int64_t) + (size_t)(n + 1) / 2 * 2 * sizeof(float);
  *staged = (sel + all <= 180 * 1024) ? 1 : 0;
  *fast = (*staged && fast_select_enabled() && 2 * k <= threads) ? 1 : 0;
  return (*staged ? sel + all : sel) + (*fast ? (size_t)threads * (sizeof(int64_t) + sizeof(uint32_t)) : 0);
}

int launch_merge_exchange(const uint64_t* gather_ll, uint32_t epoch, int world, size_t slot_words, size_t flag_word, int nq,
                          int k, float* out_s, int64_t* out_i, int* overflow_any, cudaStream_t stream) {
  int P = pow2ceil(k), staged = 0, fast = 0;
  const int threads = (nq > 512) ? kSelThreads : threads_for(world * k);
  const size_t smem = merge_smem(P, world * k, k, threads, &staged, &fast);
  VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&merge_exchange_kernel), smem));
  VODB_CUDA_CHECK(launch_pdl(merge_exchange_kernel, dim3(nq), dim3(threads), smem, stream, gather_ll, epoch, world,
                             slot_words, flag_word, nq, k, P, staged, fast, out_s, out_i, overflow_any));
  return VODB_OK;
}

int launch_merge(const float* scores, const int64_t* idx, int n_lists, int nq, int k_in, int k_out, float* out_s,
                 int64_t* out_i, cudaStream_t stream) {
  int P = pow2ceil(k_out), staged = 0, fast = 0;
  const int threads = (nq > 512) ? kSelThreads : threads_for(n_lists * k_in);
  const size_t smem = merge_smem(P, n_lists * k_in, k_out, threads, &staged, &fast);
  VODB_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(&merge_kernel), smem));
  merge_kernel<<<nq, threads, smem, stream>>>(scores, idx, n_lists, nq, k_in, k_out, P, staged, fast, out_s, out_i);
  VODB_CUDA_CHECK(cudaGetLastError());
  return VODB_OK;
}

}  // namespace vodb
