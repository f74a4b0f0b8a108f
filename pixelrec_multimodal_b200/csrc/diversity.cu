// Two more list metrics of the reference's evaluation layer on the GPU:
//
//  * Gini coefficient of the recommendation counts per item (AdvancedMetrics.calculate_gini_coefficient,
//    reference src/evaluation/advanced_metrics.py:72-105): counts sorted ascending, index 1..n,
//        G = 2 sum_i i c_(i) / (n sum c) - (n + 1) / n.
//    No sort: items are counted with integer atomics, then counted again by count value (m_v items were recommended
//    v times).  Equal counts occupy a contiguous index range after the items with smaller counts, so
//        sum_i i c_(i) = sum_v v (m_v p_v + m_v (m_v + 1) / 2),   p_v = number of items with a smaller count,
//    which one block evaluates with a scan over v in a fixed order (deterministic, integer until the last step).
//  * Intra-list similarity (NoveltyMetrics.calculate_diversity, src/evaluation/novelty.py:295-340): the mean pairwise
//    cosine similarity of the item embeddings of a list.  With unit vectors e_i,
//        mean_{i<j} cos = (|sum_i e_i|^2 - n) / (n (n - 1)),
//    so a warp per user sums the normalised embeddings of its list (no K x K similarity matrix).  The embeddings are
//    the item records already resident for scoring (pxr_precompute_items: the projected item-side modality vectors),
//    or any caller-provided (n_items, dim) fp32 table.
#include <algorithm>

#include "pxr_common.cuh"

namespace divs {

__global__ void count_items_kernel(const int32_t* __restrict__ topk, int64_t n_entries, int64_t n_items, int32_t* __restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_entries; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t it = topk[i];
    if (it >= 0 && it < n_items) atomicAdd(&counts[it], 1);
  }
}

__global__ void count_values_kernel(const int32_t* __restrict__ counts, int64_t n_items, int64_t vmax, int32_t* __restrict__ m) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = counts[i];
    atomicAdd(&m[c < vmax ? c : vmax], 1);
  }
}

// one block: out[0] = gini, out[1] = n (items in the distribution), out[2] = sum of counts
__global__ void __launch_bounds__(1024) gini_final_kernel(const int32_t* __restrict__ m, int64_t vmax, int include_zero, double* __restrict__ out) {
  __shared__ unsigned long long cnt[1024], pre[1024];
  __shared__ double part[1024], tot[1024];
  const int t = threadIdx.x;
  const int64_t lo_all = include_zero ? 0 : 1;
  const int64_t span = vmax + 1 - lo_all;
  const int64_t per = (span + 1023) / 1024;
  const int64_t v0 = lo_all + t * per, v1 = min(vmax + 1, v0 + per);
  unsigned long long c = 0;
  for (int64_t v = v0; v < v1; ++v) c += (unsigned long long)m[v];
  cnt[t] = c;
  __syncthreads();
  if (t == 0) { unsigned long long run = 0; for (int i = 0; i < 1024; ++i) { pre[i] = run; run += cnt[i]; } }
  __syncthreads();
  unsigned long long p = pre[t];
  double s = 0.0, sc = 0.0;
  for (int64_t v = v0; v < v1; ++v) {
    const unsigned long long mv = (unsigned long long)m[v];
    // exact in integers up to 2^64 for any realistic size, rounded once per value
    const double idx_sum = (double)mv * (double)p + (double)(mv * (mv + 1ull) / 2ull);
    s += (double)v * idx_sum;
    sc += (double)v * (double)mv;
    p += mv;
  }
  part[t] = s; tot[t] = sc;
  __syncthreads();
  if (t == 0) {
    double S = 0.0, C = 0.0;
    for (int i = 0; i < 1024; ++i) { S += part[i]; C += tot[i]; }
    const double n = (double)(pre[1023] + cnt[1023]);
    out[1] = n; out[2] = C;
    out[0] = (n > 0.0 && C > 0.0) ? (2.0 * S) / (n * C) - (n + 1.0) / n : 0.0;
  }
}

// inverse L2 norms of the embedding rows (0 for a zero row: the item is then skipped, like an item without embedding)
__global__ void inv_norm_kernel(const float* __restrict__ emb, int64_t n_rows, int dim, float* __restrict__ inv) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float* e = emb + row * dim;
  double s = 0.0;
  for (int d = lane; d < dim; d += 32) s += (double)e[d] * (double)e[d];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) inv[row] = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.f;
}

// one warp per user; dim <= 32 * ILS_MAX_PER_LANE.  block_sums[block][0] = sum of the users' ILS, [1] = users counted
#define ILS_MAX_PER_LANE 16
#define ILS_WARPS 8
__global__ void __launch_bounds__(32 * ILS_WARPS) ils_kernel(const int32_t* __restrict__ topk, int k_stride, int64_t n_users,
                                                             const float* __restrict__ emb, const float* __restrict__ inv, int dim,
                                                             int64_t item_base, int64_t n_rows, double* __restrict__ block_sums) {
  __shared__ double acc[ILS_WARPS][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s_ils = 0.0, s_n = 0.0;
  const int64_t warps = (int64_t)gridDim.x * ILS_WARPS;
  for (int64_t u = (int64_t)blockIdx.x * ILS_WARPS + warp; u < n_users; u += warps) {
    float sum[ILS_MAX_PER_LANE];
#pragma unroll
    for (int j = 0; j < ILS_MAX_PER_LANE; ++j) sum[j] = 0.f;
    int n = 0;
    for (int k = 0; k < k_stride; ++k) {
      const int32_t it = __ldg(topk + u * k_stride + k);
      const int64_t r = (int64_t)it - item_base;
      if (it < 0 || r < 0 || r >= n_rows) continue;                 // warp-uniform
      const float w = __ldg(inv + r);
      if (w == 0.f) continue;
      ++n;
      const float* e = emb + r * dim;
#pragma unroll
      for (int j = 0; j < ILS_MAX_PER_LANE; ++j) { const int d = lane + 32 * j; if (d < dim) sum[j] = fmaf(w, __ldg(e + d), sum[j]); }
    }
    if (n >= 2) {                                                   // novelty.py:310-321: fewer than two embeddings -> 0.0
      double ss = 0.0;
#pragma unroll
      for (int j = 0; j < ILS_MAX_PER_LANE; ++j) ss += (double)sum[j] * (double)sum[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      s_ils += (ss - (double)n) / ((double)n * (double)(n - 1));
    }
    s_n += 1.0;
  }
  if (lane == 0) { acc[warp][0] = s_ils; acc[warp][1] = s_n; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < ILS_WARPS; ++w) t += acc[w][threadIdx.x];
    block_sums[(int64_t)blockIdx.x * 2 + threadIdx.x] = t;
  }
}

__global__ void ils_final_kernel(const double* __restrict__ block_sums, int64_t n_blocks, double* __restrict__ out) {
  const int lane = threadIdx.x & 31, c = blockIdx.x;
  double s = 0.0;
  for (int64_t b = lane; b < n_blocks; b += 32) s += block_sums[b * 2 + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = s;
}

static int64_t ils_blocks(int64_t n_users) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return std::max<int64_t>(1, std::min<int64_t>((n_users + ILS_WARPS - 1) / ILS_WARPS, (int64_t)n_sm * 8));
}

}  // namespace divs

extern "C" size_t pxr_gini_bytes(int64_t n_users, int64_t n_items) {
  if (n_users < 0 || n_items < 0) return 0;
  return pxr_align_up((size_t)n_items * 4, 256) + pxr_align_up((size_t)(n_users + 2) * 4, 256) + 256;
}

extern "C" int pxr_gini(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, int64_t n_items, int32_t include_zero,
                        double* out3, void* workspace, size_t workspace_bytes, pxr_stream stream) {
  if (!out3 || k_stride <= 0 || n_users < 0 || n_items <= 0 || (n_users && !topk_idx)) return PXR_ERR_INVALID;
  if (workspace_bytes < pxr_gini_bytes(n_users, n_items) || !workspace) return PXR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* counts = (int32_t*)workspace;
  int32_t* m = (int32_t*)((char*)workspace + pxr_align_up((size_t)n_items * 4, 256));
  const int64_t vmax = n_users;                                  // an item appears at most once per list
  if (cudaMemsetAsync(workspace, 0, pxr_gini_bytes(n_users, n_items) - 256, st) != cudaSuccess) return PXR_ERR_CUDA;
  const int64_t n_entries = n_users * k_stride;
  if (n_entries) divs::count_items_kernel<<<(unsigned)std::min<int64_t>((n_entries + 255) / 256, 148 * 32), 256, 0, st>>>(topk_idx, n_entries, n_items, counts);
  divs::count_values_kernel<<<(unsigned)std::min<int64_t>((n_items + 255) / 256, 148 * 32), 256, 0, st>>>(counts, n_items, vmax, m);
  divs::gini_final_kernel<<<1, 1024, 0, st>>>(m, vmax, include_zero, out3);
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}

extern "C" size_t pxr_ils_bytes(int64_t n_users, int64_t n_rows) {
  if (n_users < 0 || n_rows < 0) return 0;
  return pxr_align_up((size_t)n_rows * 4, 256) + pxr_align_up((size_t)divs::ils_blocks(n_users) * 2 * sizeof(double), 256) + 256;
}

extern "C" int pxr_intra_list_similarity(pxr_handle* h, const int32_t* topk_idx, int32_t k_stride, int64_t n_users, const float* emb,
                                         int64_t n_rows, int32_t dim, int64_t item_base, double* out2, void* workspace,
                                         size_t workspace_bytes, pxr_stream stream) {
  if (!out2 || k_stride <= 0 || n_users < 0 || (n_users && !topk_idx)) return PXR_ERR_INVALID;
  if (!emb) {                                                     // default: the item records resident for scoring
    if (!h || !h->items_ready || !h->item_feats) return PXR_ERR_STATE;
    emb = h->item_feats; n_rows = h->n_rows; dim = (h->M - 1) * h->cfg.embedding_dim; item_base = h->item_base;
  }
  if (dim <= 0 || dim > 32 * ILS_MAX_PER_LANE || n_rows < 0) return PXR_ERR_INVALID;
  if (workspace_bytes < pxr_ils_bytes(n_users, n_rows) || !workspace) return PXR_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* inv = (float*)workspace;
  double* bs = (double*)((char*)workspace + pxr_align_up((size_t)n_rows * 4, 256));
  if (n_rows) divs::inv_norm_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(emb, n_rows, dim, inv);
  const int64_t nb = divs::ils_blocks(n_users);
  divs::ils_kernel<<<(unsigned)nb, 32 * ILS_WARPS, 0, st>>>(topk_idx, k_stride, n_users, emb, inv, dim, item_base, n_rows, bs);
  divs::ils_final_kernel<<<2, 32, 0, st>>>(bs, nb, out2);
  if (h) h->launches += 3;
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}
