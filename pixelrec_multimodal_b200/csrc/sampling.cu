// Candidate construction of the sampled evaluation protocol on the GPU.
//
// Replaces TopKRetrievalEvaluator._process_user / _sample_negatives for sampling_strategy == 'random'
// (reference src/evaluation/tasks.py:181-224, 310-364): per user the candidate list is the user's positives
// plus `n_neg` negatives drawn uniformly WITHOUT replacement from the items that are not positives, in a
// shuffled order; the recommender then ranks exactly these candidates (filter_seen = False).
//
// The reference seeds Python's `random` with `hash(str(user_id))`, which is salted per process and therefore not
// reproducible.  Here every draw is a pure function of (seed, global user index):
//   * negatives: the user walks a keyed pseudo-random PERMUTATION of [0, n_items) (4-round balanced Feistel network
//     on 2h bits, cycle-walking back into range) from position 0 and keeps the first n_neg images that are not
//     positives -- uniform, without replacement, no rejection table;
//   * shuffle: the candidates are ordered by a 64-bit hash of (user key, item) (ties -> lower item index).
// The CPU test oracle restates this in numpy (sample_candidates); the parity test asks for bit-exact equality.
#include <algorithm>

#include "pxr_common.cuh"

namespace smp {

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {     // splitmix64 step
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Perm { unsigned int rk[4]; int h; unsigned int mask; };

__device__ __forceinline__ Perm make_perm(unsigned long long ku, long long n_items) {
  Perm p;
  int bits = 1;
  while ((1ll << bits) < n_items) ++bits;
  p.h = (bits + 1) >> 1;
  p.mask = (1u << p.h) - 1u;
#pragma unroll
  for (int r = 0; r < 4; ++r) p.rk[r] = (unsigned int)(mix64(ku + (unsigned long long)r) >> 32);
  return p;
}

__device__ __forceinline__ long long perm_at(const Perm& p, long long j, long long n_items) {
  unsigned long long x = (unsigned long long)j;
  do {
    unsigned int L = (unsigned int)(x >> p.h), R = (unsigned int)x & p.mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const unsigned int f = (unsigned int)mix64(((unsigned long long)p.rk[r] << 32) | (unsigned long long)R) & p.mask;
      const unsigned int nl = R;
      R = L ^ f; L = nl;
    }
    x = ((unsigned long long)L << p.h) | (unsigned long long)R;
  } while (x >= (unsigned long long)n_items);
  return (long long)x;
}

#define SMP_WARPS 4
#define SMP_MAX_STRIDE 1024

// one warp per user; positives ascending per user (CSR), out rows padded with -1
__global__ void __launch_bounds__(32 * SMP_WARPS) sample_candidates_kernel(
    const int64_t* __restrict__ user_idx, int64_t n_users, const int64_t* __restrict__ pos_indptr,
    const int32_t* __restrict__ pos_idx, int64_t n_items, int n_neg, unsigned long long seed, int stride,
    int32_t* __restrict__ out_cand, int32_t* __restrict__ out_len) {
  extern __shared__ unsigned long long smp_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* keys = smp_smem + (size_t)warp * stride;
  int32_t* items = reinterpret_cast<int32_t*>(smp_smem + (size_t)SMP_WARPS * stride) + (size_t)warp * stride;
  for (int64_t u = (int64_t)blockIdx.x * SMP_WARPS + warp; u < n_users; u += (int64_t)gridDim.x * SMP_WARPS) {
    const int64_t g = user_idx ? user_idx[u] : u;
    const unsigned long long ku = mix64(seed ^ mix64((unsigned long long)g));
    const int64_t p0 = pos_indptr[u], p1 = pos_indptr[u + 1];
    const int npos = (int)min((int64_t)stride, p1 - p0);
    for (int i = lane; i < npos; i += 32) items[i] = pos_idx[p0 + i];
    const int64_t avail = n_items - (p1 - p0);
    const int want = (int)min((int64_t)min(n_neg, stride - npos), avail > 0 ? avail : 0);
    const Perm pm = make_perm(ku, n_items);
    int got = 0;
    for (int64_t j0 = 0; got < want; j0 += 32) {
      const int64_t j = j0 + lane;
      bool ok = j < n_items;
      int32_t x = -1;
      if (ok) {
        x = (int32_t)perm_at(pm, j, n_items);
        int64_t lo = p0, hi = p1;                    // positives are ascending: binary search
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (pos_idx[mid] < x) lo = mid + 1; else hi = mid; }
        ok = !(lo < p1 && pos_idx[lo] == x);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      const int slot = got + __popc(m & ((1u << lane) - 1u));
      if (ok && slot < want) items[npos + slot] = x;
      got += __popc(m);
    }
    const int C = npos + want;
    __syncwarp();
    for (int i = lane; i < C; i += 32) keys[i] = mix64(ku ^ 0xD1B54A32D192ED03ull ^ ((unsigned long long)(uint32_t)items[i] << 1));
    __syncwarp();
    for (int i = lane; i < C; i += 32) {             // position = rank of (key, item) among the candidates
      const unsigned long long k = keys[i];
      const int32_t it = items[i];
      int rank = 0;
      for (int c = 0; c < C; ++c) rank += (keys[c] < k) || (keys[c] == k && items[c] < it);
      out_cand[u * stride + rank] = it;
    }
    for (int i = C + lane; i < stride; i += 32) out_cand[u * stride + i] = -1;
    if (lane == 0) out_len[u] = C;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Popularity-biased strategies (reference src/evaluation/tasks.py:225-308: `np.random.choice(pool, n, replace=False, p=...)`
// with p ~ the test-set item count or its reciprocal).  Reproducible form (Efraimidis-Spirakis): item i gets the key
// log(u_i) / w_i with u_i a hash-uniform of (seed, user, item); the `want` largest keys among the user's non-positive
// items are the sample (the same distribution as successive weighted draws without replacement).  One block per user
// streams the catalogue once: a key above the running `want`-th best is pushed into a 2*kp-slot buffer in shared memory,
// a full buffer is sorted (bitonic, key descending, ties -> lower item) and cut back to `want`.  No users x items array.
// ---------------------------------------------------------------------------
#define WC_THREADS 128

__device__ __forceinline__ bool wc_before(double ka, int32_t ia, double kb, int32_t ib) {   // a sorts before b
  return ka > kb || (ka == kb && ia < ib);
}

__device__ void wc_sort_desc(double* keys, int32_t* idx, int cap) {      // all threads; cap = power of two
  for (int k = 2; k <= cap; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < cap; i += WC_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const double ka = keys[i], kb = keys[l];
          const int32_t ia = idx[i], ib = idx[l];
          const bool up = (i & k) == 0;                                  // this pair belongs to a descending run
          if (up ? wc_before(kb, ib, ka, ia) : wc_before(ka, ia, kb, ib)) { keys[i] = kb; keys[l] = ka; idx[i] = ib; idx[l] = ia; }
        }
      }
    }
  __syncthreads();
}

__global__ void __launch_bounds__(WC_THREADS) weighted_candidates_kernel(
    const int64_t* __restrict__ user_idx, int64_t n_users, const int64_t* __restrict__ pos_indptr,
    const int32_t* __restrict__ pos_idx, const double* __restrict__ weights, int64_t n_items, int n_neg,
    unsigned long long seed, int stride, int kp, int32_t* __restrict__ out_cand, int32_t* __restrict__ out_len) {
  extern __shared__ unsigned long long smp_smem[];
  const int cap = 2 * kp, tid = threadIdx.x;
  double* keys = reinterpret_cast<double*>(smp_smem);                    // [cap]
  unsigned long long* hk = smp_smem + cap;                               // [stride] shuffle hashes
  int32_t* idx = reinterpret_cast<int32_t*>(hk + stride);                // [cap]
  int32_t* items = idx + cap;                                            // [stride] positives, then the sample
  __shared__ int s_cnt;
  __shared__ double s_thr;
  const double ninf = -__longlong_as_double(0x7ff0000000000000ll);
  for (int64_t u = blockIdx.x; u < n_users; u += gridDim.x) {
    const int64_t g = user_idx ? user_idx[u] : u;
    const unsigned long long ku = mix64(seed ^ mix64((unsigned long long)g));
    const int64_t p0 = pos_indptr[u], p1 = pos_indptr[u + 1];
    const int npos = (int)min((int64_t)stride, p1 - p0);
    const int64_t avail = n_items - (p1 - p0);
    const int want = (int)max((int64_t)0, min((int64_t)min(n_neg, stride - npos), avail));
    __syncthreads();                                                     // previous user's reads of the shared arrays
    for (int i = tid; i < npos; i += WC_THREADS) items[i] = pos_idx[p0 + i];
    if (tid == 0) { s_cnt = 0; s_thr = ninf; }
    __syncthreads();
    if (want > 0) {                                                      // => every positive of the user is in items[0, npos)
      for (int64_t base = 0; base < n_items; base += WC_THREADS) {
        const int64_t i = base + tid;
        if (i < n_items) {
          int lo = 0, hi = npos;                                         // ascending positives: binary search
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (items[mid] < (int32_t)i) lo = mid + 1; else hi = mid; }
          if (!(lo < npos && items[lo] == (int32_t)i)) {
            const unsigned long long h = mix64(ku ^ 0xA0761D6478BD642Full ^ ((unsigned long long)(uint32_t)i << 1));
            const double uni = ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
            const double key = log(uni) / weights[i];
            if (key > s_thr) { const int slot = atomicAdd(&s_cnt, 1); keys[slot] = key; idx[slot] = (int32_t)i; }
          }
        }
        __syncthreads();
        const int c = s_cnt;
        __syncthreads();                                                 // everyone holds the same count before the next pushes
        if (c > cap - WC_THREADS) {
          for (int j = c + tid; j < cap; j += WC_THREADS) { keys[j] = ninf; idx[j] = 0x7fffffff; }
          wc_sort_desc(keys, idx, cap);
          if (tid == 0) { s_cnt = min(c, want); if (c >= want) s_thr = keys[want - 1]; }
          __syncthreads();
        }
      }
      const int c = s_cnt;
      for (int j = c + tid; j < cap; j += WC_THREADS) { keys[j] = ninf; idx[j] = 0x7fffffff; }
      wc_sort_desc(keys, idx, cap);
      for (int j = tid; j < want; j += WC_THREADS) items[npos + j] = idx[j];
    }
    const int C = npos + want;
    __syncthreads();
    for (int i = tid; i < C; i += WC_THREADS) hk[i] = mix64(ku ^ 0xD1B54A32D192ED03ull ^ ((unsigned long long)(uint32_t)items[i] << 1));
    __syncthreads();
    for (int i = tid; i < C; i += WC_THREADS) {                          // position = rank of (hash, item) among the candidates
      const unsigned long long k = hk[i];
      const int32_t it = items[i];
      int rank = 0;
      for (int c = 0; c < C; ++c) rank += (hk[c] < k) || (hk[c] == k && items[c] < it);
      out_cand[u * stride + rank] = it;
    }
    for (int i = C + tid; i < stride; i += WC_THREADS) out_cand[u * stride + i] = -1;
    if (tid == 0) out_len[u] = C;
  }
}

}  // namespace smp

extern "C" int pxr_sample_candidates(const int64_t* user_idx, int64_t n_users, const int64_t* pos_indptr,
                                     const int32_t* pos_idx, int64_t n_items, int32_t n_neg, uint64_t seed, int32_t stride,
                                     int32_t* out_cand, int32_t* out_len, pxr_stream stream) {
  if (n_users < 0 || n_items <= 0 || n_neg < 0 || stride <= 0 || stride > SMP_MAX_STRIDE || !pos_indptr || !out_cand || !out_len)
    return PXR_ERR_INVALID;
  if (n_users == 0) return PXR_OK;
  const size_t smem = (size_t)SMP_WARPS * stride * (sizeof(unsigned long long) + sizeof(int32_t));
  const unsigned blocks = (unsigned)std::min<int64_t>((n_users + SMP_WARPS - 1) / SMP_WARPS, 148 * 16);
  smp::sample_candidates_kernel<<<blocks, 32 * SMP_WARPS, smem, (cudaStream_t)stream>>>(
      user_idx, n_users, pos_indptr, pos_idx, n_items, n_neg, (unsigned long long)seed, stride, out_cand, out_len);
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}

extern "C" int pxr_weighted_candidates(const int64_t* user_idx, int64_t n_users, const int64_t* pos_indptr,
                                       const int32_t* pos_idx, const double* weights, int64_t n_items, int32_t n_neg,
                                       uint64_t seed, int32_t stride, int32_t* out_cand, int32_t* out_len, pxr_stream stream) {
  if (n_users < 0 || n_items <= 0 || n_items > 0x7fffffffLL || n_neg < 0 || stride <= 0 || stride > SMP_MAX_STRIDE || !pos_indptr ||
      !weights || !out_cand || !out_len)
    return PXR_ERR_INVALID;
  if (n_users == 0) return PXR_OK;
  int kp = WC_THREADS;                                                   // buffer = 2 kp slots >= want + one sweep of the block
  while (kp < std::min<int64_t>(std::min<int64_t>(n_neg, stride), n_items)) kp <<= 1;
  const size_t smem = (size_t)2 * kp * (sizeof(double) + sizeof(int32_t)) + (size_t)stride * (sizeof(unsigned long long) + sizeof(int32_t));
  const unsigned blocks = (unsigned)std::min<int64_t>(n_users, 148 * 16);
  smp::weighted_candidates_kernel<<<blocks, WC_THREADS, smem, (cudaStream_t)stream>>>(
      user_idx, n_users, pos_indptr, pos_idx, weights, n_items, n_neg, (unsigned long long)seed, stride, kp, out_cand, out_len);
  return cudaGetLastError() == cudaSuccess ? PXR_OK : PXR_ERR_CUDA;
}

extern "C" int pxr_topk_rows(pxr_handle* h, const float* scores, int64_t n_rows, int64_t n_cols, int32_t k,
                             float* out_scores, int32_t* out_pos, pxr_stream stream) {
  if (!h) return PXR_ERR_INVALID;
  if (n_rows < 0 || n_cols < 0 || k <= 0 || (n_rows && (!scores || !out_scores || !out_pos))) PXR_FAIL(h, PXR_ERR_INVALID, "pxr_topk_rows: bad arguments");
  if (n_rows == 0) return PXR_OK;
  return pxr_launch_topk_rows(h, scores, n_rows, n_cols, 0, k, out_scores, out_pos, (cudaStream_t)stream);
}
