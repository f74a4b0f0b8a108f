// Beyond-accuracy metrics on the top-K index tensor.
//
// Replaces, for all users at once, the per-user Python loops of NoveltyMetrics.calculate_metrics
// (reference src/evaluation/novelty.py:84-147: calculate_self_information :149-178, calculate_iif :180-206,
// calculate_coverage :208-226, calculate_personalized_novelty :343-377) and the O(users^2) sparse cosine of
// TopKRetrievalEvaluator._calculate_personalization (src/evaluation/tasks.py:402-427), as they are aggregated in
// TopKRetrievalEvaluator.evaluate (tasks.py:637-714).
//
// One pass over the lists (the K5 access pattern: a warp stages 32 consecutive lists with coalesced loads, lane t
// owns user t): per user the means of two per-item float64 tables over the listed items that have an entry, the
// number of distinct items, the fraction of items outside the user's history; and for personalization every list
// adds 1 / sqrt(|list|) to a per-item accumulator s_i, because for binary list vectors
//   sum_{u<v} cos(u, v) = ( sum_i s_i^2 - #non-empty lists ) / 2.
// s_i is accumulated in 2^-40 fixed point with integer atomics, so the result does not depend on the order of the adds.
#include <algorithm>

#include "pxr_common.cuh"

namespace nov {

#define NOV_WARPS 4
constexpr double FIX = 1099511627776.0;      // 2^40

__global__ void __launch_bounds__(32 * NOV_WARPS) novelty_kernel(
    const int32_t* __restrict__ topk, int k_stride, int64_t n_users, int64_t users_per_warp,
    const double* __restrict__ self_info, const double* __restrict__ iif, const int64_t* __restrict__ hist_indptr,
    const int32_t* __restrict__ hist_idx, unsigned long long* __restrict__ item_acc, double* __restrict__ block_sums) {
  __shared__ double acc[NOV_WARPS][5];
  __shared__ int32_t lists[NOV_WARPS][32 * 65];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = k_stride | 1;
  int32_t* L = lists[warp];
  double s_self = 0.0, s_iif = 0.0, s_uniq = 0.0, s_pnov = 0.0, s_n = 0.0;
  const int64_t w = (int64_t)blockIdx.x * NOV_WARPS + warp;
  const int64_t u0 = w * users_per_warp, u1 = min(n_users, u0 + users_per_warp);
  for (int64_t ub = u0; ub < u1; ub += 32) {
    const int nb = (int)min((int64_t)32, u1 - ub);
    const int32_t* src = topk + ub * k_stride;
    for (int i0 = lane; i0 < nb * k_stride; i0 += 32 * 8) {          // 8 independent loads per lane in flight
      int32_t v[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) { const int i = i0 + 32 * m; if (i < nb * k_stride) v[m] = src[i]; }
#pragma unroll
      for (int m = 0; m < 8; ++m) { const int i = i0 + 32 * m; if (i < nb * k_stride) L[(i / k_stride) * ld + (i % k_stride)] = v[m]; }
    }
    __syncwarp();
    if (lane < nb) {
      const int32_t* rec = L + lane * ld;
      int n = 0, n_si = 0, n_iif = 0, uniq = 0, novel = 0;
      double a_si = 0.0, a_iif = 0.0;
      int64_t h0 = 0, h1 = 0;
      if (hist_indptr) { h0 = hist_indptr[ub + lane]; h1 = hist_indptr[ub + lane + 1]; }
      for (int j = 0; j < k_stride; ++j) {
        const int32_t it = rec[j];
        if (it < 0) continue;
        ++n;
        const double a = self_info[it], b = iif[it];
        if (a == a) { a_si += a; ++n_si; }            // NaN = the item has no entry (not in the interaction table)
        if (b == b) { a_iif += b; ++n_iif; }
        bool first = true;
        for (int jj = 0; jj < j; ++jj) first = first && (rec[jj] != it);
        uniq += first;
        int64_t lo = h0, hi = h1;                     // history ascending per user
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (hist_idx[mid] < it) lo = mid + 1; else hi = mid; }
        novel += !(lo < h1 && hist_idx[lo] == it);
      }
      if (n > 0) {                                    // "if not recommendations: return {}" (novelty.py:104-105)
        s_self += n_si ? a_si / (double)n_si : 0.0;
        s_iif += n_iif ? a_iif / (double)n_iif : 0.0;
        s_uniq += (double)uniq;
        s_pnov += (double)novel / (double)n;
        s_n += 1.0;
        const unsigned long long wfix = (unsigned long long)(FIX / sqrt((double)uniq) + 0.5);
        for (int j = 0; j < k_stride; ++j) {
          const int32_t it = rec[j];
          if (it < 0) continue;
          bool first = true;
          for (int jj = 0; jj < j; ++jj) first = first && (rec[jj] != it);
          if (first) atomicAdd(&item_acc[it], wfix);
        }
      }
    }
    __syncwarp();
  }
  double v[5] = {s_self, s_iif, s_uniq, s_pnov, s_n};
#pragma unroll
  for (int c = 0; c < 5; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if (lane == 0) acc[warp][c] = v[c];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0.0;
    for (int ww = 0; ww < NOV_WARPS; ++ww) t += acc[ww][threadIdx.x];
    block_sums[(int64_t)blockIdx.x * 8 + threadIdx.x] = t;
  }
}

// out[0..4] = fixed-order sums of the block partials; out[5] = sum_i s_i^2 (fixed order per lane + butterfly)
__global__ void novelty_final_kernel(const double* __restrict__ block_sums, int64_t n_blocks,
                                     const unsigned long long* __restrict__ item_acc, int64_t n_items, double* out) {
  const int t = blockIdx.x, lane = threadIdx.x & 31;
  double s = 0.0;
  if (t < 5) {
    for (int64_t b = lane; b < n_blocks; b += 32) s += block_sums[b * 8 + t];
  } else {
    for (int64_t i = lane; i < n_items; i += 32) { const double si = (double)item_acc[i] / FIX; s += si * si; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[t] = s;
}

}  // namespace nov

extern "C" size_t pxr_novelty_bytes(int64_t n_users, int64_t n_items) {
  const int64_t blocks = std::max<int64_t>(1, (n_users + 127) / 128);
  return pxr_align_up((size_t)blocks * 8 * sizeof(double), 256) + pxr_align_up((size_t)std::max<int64_t>(n_items, 1) * sizeof(unsigned long long), 256);
}

extern "C" int pxr_novelty_metrics(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, int64_t n_items,
                                   const double* self_info, const double* iif, const int64_t* hist_indptr,
                                   const int32_t* hist_idx, double* out6, void* workspace, size_t workspace_bytes,
                                   pxr_stream stream) {
  if (k_stride <= 0 || k_stride > 64 || n_users < 0 || n_items <= 0 || !self_info || !iif || !out6) return PXR_ERR_INVALID;
  if (workspace_bytes < pxr_novelty_bytes(n_users, n_items) || !workspace) return PXR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = std::max<int64_t>(1, (n_users + 127) / 128);
  double* block_sums = (double*)workspace;
  unsigned long long* item_acc = (unsigned long long*)((char*)workspace + pxr_align_up((size_t)blocks * 8 * sizeof(double), 256));
  if (cudaMemsetAsync(item_acc, 0, (size_t)n_items * sizeof(unsigned long long), st) != cudaSuccess) return PXR_ERR_CUDA;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  blocks = std::min<int64_t>(blocks, (int64_t)n_sm * 16);
  const int64_t warps = blocks * NOV_WARPS;
  const int64_t upw = (n_users + warps - 1) / warps;
  nov::novelty_kernel<<<(unsigned)blocks, 32 * NOV_WARPS, 0, st>>>(topk_idx, k_stride, n_users, upw, self_info, iif, hist_indptr,
                                                                  hist_idx, item_acc, block_sums);
  nov::novelty_final_kernel<<<6, 32, 0, st>>>(block_sums, blocks, item_acc, n_items, out6);
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}
