// Glue kernels of the retrieve -> sample chain (vodb_retrieve_sample, api.cu): what RealmCollate does between the
// dense search and the batch it returns, for the dense-only flow, without leaving the GPU.
//
//   match_labels_kernel   labels[b,j] = retrieved id (b,j) is one of row b's gold section ids
//                         (the reference gets these labels from the `lookup` engine and the union-merge,
//                          src/vod_dataloaders/core/search.py:79-125; with one engine the match is a direct compare)
//   gather_picks_kernel   src/vod_dataloaders/core/sample.py:57-71: take_along_axis of ids / scores at the sampled
//                         local positions (negative positions wrap around like numpy) and
//                         max_sampling_id[b] = #{j : not positive, finite, score_j >= min finite sampled negative score}
#include "common.cuh"

namespace vodb {
namespace {

__global__ void match_labels_kernel(const int64_t* __restrict__ idx, const int64_t* __restrict__ gold, int n_gold, int K,
                                    int B, uint8_t* __restrict__ labels) {
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * K) return;
  const int b = (int)(i / K);
  const int64_t id = idx[i];
  bool hit = false;
  if (id >= 0)  // -1 pads never match; negative gold ids are padding of ragged gold lists
    for (int p = 0; p < n_gold; ++p) hit |= gold[(size_t)b * n_gold + p] == id;
  labels[i] = hit ? 1 : 0;
}

constexpr int kGatherThreads = 128;

__global__ void __launch_bounds__(kGatherThreads)
gather_picks_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, const uint8_t* __restrict__ labels,
                    int K, int k_total, const int64_t* __restrict__ local, const uint8_t* __restrict__ picked_labels,
                    int64_t* __restrict__ out_idx, float* __restrict__ out_scores, float* __restrict__ out_msid) {
  __shared__ float s_min;
  __shared__ int s_count[kGatherThreads / 32];
  pdl_wait();
  const int b = blockIdx.x, t = threadIdx.x;
  const float* row_s = scores + (size_t)b * K;
  if (t == 0) s_min = vm_u2f(0x7f800000u);
  __syncthreads();
  // gathers; the minimum over <= k_total values is order-free, one thread folds it
  for (int j = t; j < k_total; j += kGatherThreads) {
    int64_t l = local[(size_t)b * k_total + j];
    if (l < 0) l += K;  // numpy take_along_axis wrap-around: the sampler's -1 (unused slot) reads the last column
    out_idx[(size_t)b * k_total + j] = idx[(size_t)b * K + l];
    out_scores[(size_t)b * k_total + j] = row_s[l];
  }
  __syncthreads();
  if (t == 0) {
    float m = vm_u2f(0x7f800000u);
    for (int j = 0; j < k_total; ++j) {
      const float v = out_scores[(size_t)b * k_total + j];
      const bool finite = (vm_f2u(v) & 0x7fffffffu) < 0x7f800000u;
      if (picked_labels[(size_t)b * k_total + j] == 0 && finite && v < m) m = v;
    }
    s_min = m;
  }
  __syncthreads();
  const float m = s_min;
  int c = 0;
  for (int j = t; j < K; j += kGatherThreads) {
    const float v = row_s[j];
    const bool finite = (vm_f2u(v) & 0x7fffffffu) < 0x7f800000u;
    const bool neg = labels == nullptr || labels[(size_t)b * K + j] == 0;
    c += (neg && finite && v >= m) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((t & 31) == 0) s_count[t >> 5] = c;
  __syncthreads();
  if (t == 0) {
    int total = 0;
    for (int w = 0; w < kGatherThreads / 32; ++w) total += s_count[w];
    out_msid[b] = (float)total;  // the reference sums float32 ones: exact for counts < 2^24
  }
}

}  // namespace

int launch_match_labels(const int64_t* idx, const int64_t* gold, int n_gold, int B, int K, uint8_t* labels,
                        cudaStream_t stream) {
  if ((size_t)B * K == 0) return VODB_OK;
  const unsigned grid = (unsigned)(((size_t)B * K + 255) / 256);
  cudaError_t e = launch_pdl(match_labels_kernel, dim3(grid), dim3(256), 0, stream, idx, gold, n_gold, K, B, labels);
  if (e != cudaSuccess) {
    set_error("match_labels launch: %s", cudaGetErrorString(e));
    return VODB_ECUDA;
  }
  return VODB_OK;
}

int launch_gather_picks(const float* scores, const int64_t* idx, const uint8_t* labels, int B, int K, int k_total,
                        const int64_t* local, const uint8_t* picked_labels, int64_t* out_idx, float* out_scores,
                        float* out_msid, cudaStream_t stream) {
  if (B == 0) return VODB_OK;
  cudaError_t e = launch_pdl(gather_picks_kernel, dim3(B), dim3(kGatherThreads), 0, stream, scores, idx, labels, K,
                             k_total, local, picked_labels, out_idx, out_scores, out_msid);
  if (e != cudaSuccess) {
    set_error("gather_picks launch: %s", cudaGetErrorString(e));
    return VODB_ECUDA;
  }
  return VODB_OK;
}

}  // namespace vodb
