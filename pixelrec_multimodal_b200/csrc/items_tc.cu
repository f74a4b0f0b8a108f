// K1 + K2 on the tensor pipe: item-side gather and the dense item-only contractions of the scoring path
// (reference src/models/multimodal.py:553-570: embedding lookups and the modality projections
// `act(W x + b)`, :262-313; for concat fusion also the item partial of layer 1).
//
// The item records feed BOTH the fp32 generic kernels (forward / get_item_score, which are checked against
// the reference to 1e-5) and the fused tcgen05 kernel, so the contractions must keep fp32 accuracy: they run
// as 3xTF32 -- every fp32 operand is split into hi = tf32(x) and lo = tf32(x - hi) and
//     A.W  ~=  A_hi.W_hi + A_lo.W_hi + A_hi.W_lo      (the dropped lo.lo term is 2^-22 relative),
// three tcgen05.mma kind::tf32 per K step, fp32 accumulation in TMEM.  The stage is HBM-bound (3.6 KB of cached
// features read per item against 0.35 MFLOP), which a CUDA-core fp32 version cannot reach (measured 6 % of
// the HBM roofline).
//
//   gemm3x_kernel     C[M x N] = epilogue(A[M x K] . W[N x K]^T + b), 128-row tiles, persistent CTAs, 16 warps:
//     warps 0-7   stream A from global memory (coalesced 16-byte loads, three K chunks = 48 KB in flight per SM),
//                 split it into hi / lo in registers and write both operand tiles in the 128-byte-swizzled
//                 K-major layout the MMA descriptors read (the split needs a CUDA-core pass anyway, so a TMA
//                 copy of the raw tile would only add a shared-memory round trip);
//     warp  8     TMA engine: bulk copies of the pre-split, pre-swizzled weight chunk images (hi and lo);
//     warp  9     issues the MMAs (one thread), 4-stage (N <= 64) or 2-stage (N = 256) mbarrier ring,
//                 two TMEM accumulators so the epilogue of tile t overlaps the main loop of tile t + 1;
//     warps 12-15 epilogue: tcgen05.ld, bias, activation, fp32 record or 16-bit partial store.
//   gather_small_kernel   item / tag embedding rows copied by the TMA engine (bulk global -> shared -> global,
//                 no register staging), numerical projection (K = 7) on CUDA cores.
#include <algorithm>

#include <cuda_fp16.h>

#include "pxr_common.cuh"
#include "tc_ptx.cuh"

namespace itc {

constexpr int TM = 128;                 // rows per tile
constexpr int KC = 32;                  // K elements per chunk = one 128-byte swizzle row of fp32
constexpr int THREADS = 512;
constexpr int LOADER_WARPS = 8;
constexpr uint32_t A_TILE = TM * 128;   // bytes of one A operand tile (hi or lo)

enum { OUT_F32_ACT = 0, OUT_16 = 1 };

struct GemmParams {
  const float* A; int64_t lda;          // [M][lda] fp32, 16-byte aligned rows, K % 4 == 0
  int64_t M; int K, N, NT;              // NT = columns per N tile (N % NT == 0, NT <= 256, NT % 16 == 0)
  const uint8_t* wimg;                  // [N / NT][chunks][2][NT x 128 B] swizzled hi / lo chunk images
  const float* bias;                    // [N]
  int mode, act, fmt16;                 // OUT_F32_ACT: out_f[row * ldo + n] = act(acc + b) (act < 0: identity); OUT_16: 16-bit, no act
  int n_store;                          // columns of each N tile that are stored (0 = all): padded outputs
  int qt_nm;                            // OUT_16, > 0: row = (item, modality m < qt_nm), stored chunk-major for the wide gated front end
                                        // of score_tc.cu: [item / 16][column / 64][item % 16][qt_nm x 64 columns + 8 of padding]
  float* out_f; uint16_t* out_h; int64_t ldo;
  int n_stages;
};

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == 0) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void __launch_bounds__(THREADS, 1) gemm3x_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  const uint32_t b_tile = (uint32_t)p.NT * 128u;
  const uint32_t stage_bytes = 2 * A_TILE + 2 * b_tile;            // A_hi | A_lo | B_hi | B_lo
  const int NS = p.n_stages;
  __shared__ unsigned long long bars[2 * 4 + 4];                   // full[NS], empty[NS], acc_full[2], acc_empty[2]
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = ptx::smem_u32(&bars[0]);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (4 + s); };
  auto ACC_FULL = [&](int a) { return bar0 + 8u * (8 + a); };
  auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (10 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = (p.K + KC - 1) / KC;
  const int n_ntiles = p.N / p.NT;
  const int64_t m_tiles = (p.M + TM - 1) / TM;
  const int64_t n_work = m_tiles * n_ntiles;                       // work item w: m tile w / n_ntiles, n tile w % n_ntiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { ptx::mbar_init(FULL(s), LOADER_WARPS + 1); ptx::mbar_init(EMPTY(s), 1); }   // loader warps + weight producer
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(ACC_FULL(a), 1); ptx::mbar_init(ACC_EMPTY(a), 4); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (warp == 9) ptx::tmem_alloc_1cta(ptx::smem_u32(&tmem_slot), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < LOADER_WARPS) {
    // ============================================================ A loaders: global -> registers -> hi / lo tiles
    const int tid = threadIdx.x, r32 = tid >> 3, c = tid & 7;       // 8 lanes cover one 128-byte row segment
    int it = 0;
    for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int64_t row0 = (w / n_ntiles) * TM;
      float4 cur[4], nx1[4], nx2[4];
      auto load_chunk = [&](int kc, float4* v) {
        const int col = kc * KC + 4 * c;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t row = row0 + i * 32 + r32;
          v[i] = (row < p.M && col < p.K && kc < n_chunks) ? __ldcs(reinterpret_cast<const float4*>(p.A + row * p.lda + col))
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_chunk(0, cur);
      load_chunk(1, nx1);
      for (int kc = 0; kc < n_chunks; ++kc, ++it) {
        load_chunk(kc + 2, nx2);
        const int s = it % NS;
        if (it >= NS) ptx::mbar_wait(EMPTY(s), ((it / NS) - 1) & 1);
        uint8_t* a_hi = sm + (size_t)s * stage_bytes;
        uint8_t* a_lo = a_hi + A_TILE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = i * 32 + r32;
          const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
          float4 hi, lo;
          hi.x = ptx::to_tf32(cur[i].x); hi.y = ptx::to_tf32(cur[i].y); hi.z = ptx::to_tf32(cur[i].z); hi.w = ptx::to_tf32(cur[i].w);
          lo.x = ptx::to_tf32(cur[i].x - hi.x); lo.y = ptx::to_tf32(cur[i].y - hi.y);
          lo.z = ptx::to_tf32(cur[i].z - hi.z); lo.w = ptx::to_tf32(cur[i].w - hi.w);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_local(FULL(s));
#pragma unroll
        for (int i = 0; i < 4; ++i) { cur[i] = nx1[i]; nx1[i] = nx2[i]; }
      }
    }
  } else if (warp == LOADER_WARPS) {
    // ============================================================ weight chunk images through the TMA engine
    if (lane == 0) {
      int it = 0;
      for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int nt = (int)(w % n_ntiles);
        const uint8_t* src = p.wimg + (size_t)nt * n_chunks * 2 * b_tile;
        for (int kc = 0; kc < n_chunks; ++kc, ++it) {
          const int s = it % NS;
          if (it >= NS) ptx::mbar_wait(EMPTY(s), ((it / NS) - 1) & 1);
          const uint32_t dst = base + (uint32_t)s * stage_bytes + 2 * A_TILE;
          ptx::mbar_expect_tx(FULL(s), 2 * b_tile);
          for (uint32_t o = 0; o < 2 * b_tile; o += 8192)
            ptx::bulk_g2s(dst + o, src + (size_t)kc * 2 * b_tile + o, min(8192u, 2 * b_tile - o), FULL(s));
        }
      }
    }
  } else if (warp == LOADER_WARPS + 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::idesc_tf32(TM, p.NT);
      // N <= 128: the hi and lo weight tiles are adjacent in the stage, so ONE MMA with N = 2 NT multiplies A_hi by both
      // (columns [0, NT) collect A_hi.B_hi, columns [NT, 2 NT) collect A_hi.B_lo; the epilogue adds the halves).  An SS MMA
      // re-reads both operand tiles from shared memory (4 KB of A per K step), which -- not the math -- bounds N = 64:
      // two MMAs per K step read 14 KB instead of 18 KB in three.
      const bool fold = p.NT <= 128;
      const uint32_t idesc2 = ptx::idesc_tf32(TM, fold ? 2 * p.NT : p.NT);
      int it = 0, t_local = 0;
      for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++t_local) {
        const int acc = t_local & 1;
        if (t_local >= 2) { ptx::mbar_wait(ACC_EMPTY(acc), ((t_local >> 1) - 1) & 1); ptx::tc_fence_after(); }
        const uint32_t d = tmem + (uint32_t)acc * 256u;
        for (int kc = 0; kc < n_chunks; ++kc, ++it) {
          const int s = it % NS;
          ptx::mbar_wait(FULL(s), (it / NS) & 1);
          ptx::tc_fence_after();
          const uint32_t sb = base + (uint32_t)s * stage_bytes;
          const uint64_t dAh = ptx::smem_desc_sw128(sb), dAl = ptx::smem_desc_sw128(sb + A_TILE);
          const uint64_t dBh = ptx::smem_desc_sw128(sb + 2 * A_TILE), dBl = ptx::smem_desc_sw128(sb + 2 * A_TILE + b_tile);
#pragma unroll
          for (int kk = 0; kk < KC / 8; ++kk) {          // K = 8 per tf32 MMA = 32 bytes along the swizzled row
            if (fold) {
              ptx::mma1_tf32_ss(d, dAh + 2 * kk, dBh + 2 * kk, idesc2, (kc > 0 || kk > 0));
              ptx::mma1_tf32_ss(d, dAl + 2 * kk, dBh + 2 * kk, idesc, 1);
            } else {
              ptx::mma1_tf32_ss(d, dAl + 2 * kk, dBh + 2 * kk, idesc, (kc > 0 || kk > 0));
              ptx::mma1_tf32_ss(d, dAh + 2 * kk, dBl + 2 * kk, idesc, 1);
              ptx::mma1_tf32_ss(d, dAh + 2 * kk, dBh + 2 * kk, idesc, 1);
            }
          }
          ptx::commit1(EMPTY(s));
        }
        ptx::commit1(ACC_FULL(acc));
      }
    }
  } else if (warp >= 12) {
    // ============================================================ epilogue
    const int q = warp & 3;
    int t_local = 0;
    for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++t_local) {
      const int acc = t_local & 1;
      const int nt = (int)(w % n_ntiles);
      const int64_t row = (w / n_ntiles) * TM + q * 32 + lane;
      ptx::mbar_wait(ACC_FULL(acc), (t_local >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      for (int n0 = 0; n0 < p.NT; n0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(t0 + n0, v);
        if (p.NT <= 128) {                 // folded form: second half of the accumulator holds A_hi.B_lo
          uint32_t v2[32];
          ptx::tmem_ld32(t0 + p.NT + n0, v2);
          ptx::tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        }
        ptx::tc_wait_ld();
        if (row < p.M) {
          const float* b = p.bias + nt * p.NT + n0;
          const int ncol = min(32, (p.n_store ? p.n_store : p.NT) - n0);   // NT is a multiple of 16: the last chunk may be half full
          if (p.mode == OUT_F32_ACT) {
            float* o = p.out_f + row * p.ldo + nt * p.NT + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (i >= ncol) break;
              float4 r;
              r.x = __uint_as_float(v[i]) + b[i]; r.y = __uint_as_float(v[i + 1]) + b[i + 1];
              r.z = __uint_as_float(v[i + 2]) + b[i + 2]; r.w = __uint_as_float(v[i + 3]) + b[i + 3];
              if (p.act >= 0) { r.x = pxr_apply_act(r.x, p.act); r.y = pxr_apply_act(r.y, p.act); r.z = pxr_apply_act(r.z, p.act); r.w = pxr_apply_act(r.w, p.act); }
              *reinterpret_cast<float4*>(o + i) = r;
            }
          } else {
            uint16_t* o = p.out_h + row * p.ldo + nt * p.NT + n0;
            if (p.qt_nm > 0) {
              const int64_t item = row / p.qt_nm; const int m = (int)(row % p.qt_nm);
              const int n = nt * p.NT + n0;                            // 32 columns inside one 64-column chunk
              const size_t item_b = (size_t)p.qt_nm * 128 + 16, stage_b = 16 * item_b;
              o = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(p.out_h) + ((size_t)(item >> 4) * 8 + (n >> 6)) * stage_b +
                                              (size_t)(item & 15) * item_b + m * 128 + (n & 63) * 2);
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                pk[j] = (uint32_t)to16(__uint_as_float(v[i + 2 * j]) + b[i + 2 * j], p.fmt16) |
                        ((uint32_t)to16(__uint_as_float(v[i + 2 * j + 1]) + b[i + 2 * j + 1], p.fmt16) << 16);
              *reinterpret_cast<uint4*>(o + i) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_local(ACC_EMPTY(acc));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == LOADER_WARPS + 1) ptx::tmem_dealloc_1cta(tmem, 512);
}

// Weight W[N x K] (row stride ldw, first column k_off) -> per N tile, per K chunk: hi image | lo image, each NT rows of
// 128 bytes in the swizzled K-major layout (16-byte chunk index XOR row-in-group); columns past K are zero.
__global__ void split_weights_kernel(const float* __restrict__ W, int64_t ldw, int k_off, int N, int K, int NT,
                                     uint8_t* __restrict__ img, int n_valid) {
  const int n_chunks = (K + KC - 1) / KC;
  const int64_t total = (int64_t)N * n_chunks * KC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kin = (int)(e % KC);
    const int kc = (int)((e / KC) % n_chunks);
    const int n = (int)(e / ((int64_t)KC * n_chunks));
    const int nt = n / NT, nl = n % NT;
    const int k = kc * KC + kin;
    const float v = (k < K && n < n_valid) ? W[(int64_t)n * ldw + k_off + k] : 0.f;   // rows past n_valid: zero padding
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float hi = __uint_as_float(hb);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v - hi));
    const size_t b_tile = (size_t)NT * 128;
    uint8_t* chunk = img + ((size_t)nt * n_chunks + kc) * 2 * b_tile;
    const uint32_t off = (uint32_t)(nl >> 3) * 1024u + (uint32_t)(nl & 7) * 128u + (uint32_t)(((kin >> 2) ^ (nl & 7)) << 4) + (kin & 3) * 4;
    *reinterpret_cast<uint32_t*>(chunk + off) = hb;
    *reinterpret_cast<uint32_t*>(chunk + b_tile + off) = lb;
  }
}

// Slots 0 / 1 of the item record: item and tag embedding rows, moved by the TMA engine (bulk copy global -> shared,
// bulk copy shared -> global; the payload never enters registers); numerical projection act(W x + b) with K <= 32
// on CUDA cores.  32 items per block of 128 threads.
struct GatherParams {
  const float* item_embedding; const int64_t* item_idx; int64_t item_base;
  const float* tag_emb; const int64_t* tag_idx;
  const float* num; int num_dim; const float* num_wt; const float* num_b; int num_slot; int act;
  float* feats; int FD, D; int64_t n_rows;
  int rows;                          // items per block (<= 32; fewer for wide embeddings so the staging fits 48 KB)
};

__global__ void __launch_bounds__(128) gather_small_kernel(const GatherParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  __shared__ unsigned long long bar;
  const int rows = p.rows;
  const int64_t row0 = (int64_t)blockIdx.x * rows;
  const int n = (int)min((int64_t)rows, p.n_rows - row0);
  const uint32_t rb = (uint32_t)p.D * 4;                            // bytes of one embedding row
  const uint32_t sbase = ptx::smem_u32(gsm), b = ptx::smem_u32(&bar);
  if (threadIdx.x == 0) { ptx::mbar_init(b, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    if (r == 0) ptx::mbar_expect_tx(b, 2u * rb * n);
    __syncwarp();
    if (r < n) {
      const int64_t it = p.item_idx ? p.item_idx[row0 + r] : (p.item_base + row0 + r);
      ptx::bulk_g2s(sbase + (2 * r) * rb, p.item_embedding + it * p.D, rb, b);
      ptx::bulk_g2s(sbase + (2 * r + 1) * rb, p.tag_emb + p.tag_idx[row0 + r] * p.D, rb, b);
    }
    ptx::mbar_wait(b, 0);
    if (r < n) {                                                    // slots 0 and 1 are adjacent in the record
      ptx::bulk_s2g(p.feats + (row0 + r) * p.FD, sbase + (2 * r) * rb, 2 * rb);
      ptx::bulk_commit();
      ptx::bulk_wait_read0();
    }
  } else if (p.num) {
    // numerical projection: thread = (row, 4 outputs); 96 threads sweep 32 rows x D / 4 column groups
    for (int idx = threadIdx.x - 32; idx < n * (p.D / 4); idx += 96) {
      const int r = idx / (p.D / 4), c4 = (idx % (p.D / 4)) * 4;
      const float* x = p.num + (row0 + r) * p.num_dim;
      float4 acc = *reinterpret_cast<const float4*>(p.num_b + c4);
      for (int k = 0; k < p.num_dim; ++k) {
        const float xv = x[k];
        const float4 wv = *reinterpret_cast<const float4*>(p.num_wt + (size_t)k * p.D + c4);
        acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
      }
      acc.x = pxr_apply_act(acc.x, p.act); acc.y = pxr_apply_act(acc.y, p.act);
      acc.z = pxr_apply_act(acc.z, p.act); acc.w = pxr_apply_act(acc.w, p.act);
      *reinterpret_cast<float4*>(p.feats + (row0 + r) * p.FD + p.num_slot * p.D + c4) = acc;
    }
  }
}

struct ItemImages {          // device pointers into h->tc_items_w
  uint8_t* proj[2];          // vision, language projection images
  uint8_t* pi;               // concat: item columns of layer 1
};

static size_t img_bytes(int N, int K) { return (size_t)2 * N * ((K + KC - 1) / KC) * KC * 4; }

static int n_tile_for(int N) {          // largest divisor of N that is a multiple of 16 and at most 256
  for (int nt = std::min(N, 256) / 16 * 16; nt >= 16; nt -= 16) if (N % nt == 0) return nt;
  return 0;
}

static int launch_gemm(pxr_handle* h, const GemmParams& gp, cudaStream_t st) {
  GemmParams p = gp;
  p.n_stages = p.NT <= 64 ? 4 : 2;
  const size_t smem = (size_t)p.n_stages * (2 * A_TILE + 2 * (size_t)p.NT * 128) + 1024;
  if (!(h->tc_attr_set & (1ull << 62))) {
    PXR_CUDA(h, cudaFuncSetAttribute(gemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    h->tc_attr_set |= (1ull << 62);
  }
  const int64_t n_work = ((p.M + TM - 1) / TM) * (p.N / p.NT);
  const unsigned grid = (unsigned)std::min<int64_t>(n_work, h->n_sm);
  gemm3x_kernel<<<grid, THREADS, smem, st>>>(p);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

}  // namespace itc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// The tensor-pipe item path covers single-layer projections with 16-byte aligned feature rows.
bool pxr_items_tc_supported(const pxr_handle* h) {
  const pxr_config& c = h->cfg;
  if (!h->fast_ok || c.projection_hidden != 0 || c.embedding_dim % 16 != 0 || c.embedding_dim > 512) return false;
  if ((c.vision_dim && c.vision_dim % 4) || (c.language_dim && c.language_dim % 4) || c.num_numerical > 32) return false;
  return true;
}

int pxr_items_tc_prepare_weights(pxr_handle* h, cudaStream_t st) {
  if (!pxr_items_tc_supported(h)) return PXR_OK;
  const pxr_config& c = h->cfg;
  const int D = c.embedding_dim;
  const int kdim[2] = {c.vision_dim, c.language_dim};
  size_t total = 0;
  for (int m = 0; m < 2; ++m) if (h->has_mod[m]) total += pxr_align_up(itc::img_bytes(D, kdim[m]), 1024);
  const bool concat = c.fusion == PXR_FUSION_CONCAT;
  const int FD = (h->M - 1) * D;
  if (concat) total += pxr_align_up(itc::img_bytes(PXR_TC_H1, FD), 1024);          // layer-1 partials: 512 columns, rows past hidden[0] zero
  const bool gated = c.fusion == PXR_FUSION_GATED;
  const bool wide = pxr_tc_gated_wide(h);     // gated at embedding_dim != 64: per-modality layer-1 partials, W1 (512 x D) as one image
  if (gated) total += pxr_align_up(itc::img_bytes(16, FD), 1024) + 1024;      // gate logits: 16 padded outputs + padded bias
  if (wide) total += pxr_align_up(itc::img_bytes(PXR_TC_H1, D), 1024);
  if (h->tc_items_w) { cudaFree(h->tc_items_w); h->tc_items_w = nullptr; }
  PXR_CUDA(h, cudaMalloc(&h->tc_items_w, total + 1024));
  uint8_t* cur = reinterpret_cast<uint8_t*>(h->tc_items_w);
  for (int m = 0; m < 2; ++m) {
    h->tc_items_img[m] = nullptr;
    if (!h->has_mod[m]) continue;
    h->tc_items_img[m] = cur;
    itc::split_weights_kernel<<<256, 256, 0, st>>>(h->proj[m][0].w, kdim[m], 0, D, kdim[m], itc::n_tile_for(D), cur, D);
    h->launches++;
    cur += pxr_align_up(itc::img_bytes(D, kdim[m]), 1024);
  }
  h->tc_items_img[2] = nullptr;
  if (concat) {
    h->tc_items_img[2] = cur;
    itc::split_weights_kernel<<<512, 256, 0, st>>>(h->mlp[0].w, h->mlp[0].k, D, PXR_TC_H1, FD, 256, cur, c.hidden[0]);
    h->launches++;
  }
  if (wide) {
    h->tc_items_img[2] = cur;
    itc::split_weights_kernel<<<512, 256, 0, st>>>(h->mlp[0].w, h->mlp[0].k, 0, PXR_TC_H1, D, 256, cur, c.hidden[0]);
    h->launches++;
    cur += pxr_align_up(itc::img_bytes(PXR_TC_H1, D), 1024);
  }
  h->tc_items_img[3] = nullptr; h->tc_gate_bias = nullptr;
  if (gated) {
    if (concat) cur += pxr_align_up(itc::img_bytes(PXR_TC_H1, FD), 1024);
    h->tc_items_img[3] = cur;
    itc::split_weights_kernel<<<64, 256, 0, st>>>(h->gate.w, (int64_t)h->M * D, D, 16, FD, 16, cur, h->M);
    h->launches++;
    cur += pxr_align_up(itc::img_bytes(16, FD), 1024);
    h->tc_gate_bias = reinterpret_cast<float*>(cur);
    PXR_CUDA(h, cudaMemsetAsync(cur, 0, 64, st));
    PXR_CUDA(h, cudaMemcpyAsync(cur, h->gate.b, sizeof(float) * h->M, cudaMemcpyDeviceToDevice, st));
  }
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

// gated: item part of the gate logits, Wg[:, D:] . record + bg  (layers.py:207 split) -> [rows][8] fp32
int pxr_launch_item_logit_tc(pxr_handle* h, int64_t n_rows, float* out, cudaStream_t st) {
  const pxr_config& c = h->cfg;
  const int D = c.embedding_dim, FD = (h->M - 1) * D;
  itc::GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.A = h->item_feats; gp.lda = FD; gp.M = n_rows; gp.K = FD; gp.N = 16; gp.NT = 16;
  gp.wimg = h->tc_items_img[3]; gp.bias = h->tc_gate_bias;
  gp.mode = itc::OUT_F32_ACT; gp.act = -1; gp.out_f = out; gp.ldo = 8; gp.n_store = 8;
  return itc::launch_gemm(h, gp, st);
}

// wide gated: the M - 1 layer-1 partials of every item, Q[row][m] = W1 f_m + b1 -> 16 bit: the record viewed as
// n_rows * (M - 1) vectors of embedding_dim (it is stored [rows][M-1][D]) times W1^T
int pxr_launch_item_q_tc(pxr_handle* h, int64_t n_rows, uint16_t* out, int fmt16, cudaStream_t st) {
  const pxr_config& c = h->cfg;
  const int D = c.embedding_dim;
  itc::GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.A = h->item_feats; gp.lda = D; gp.M = n_rows * (h->M - 1); gp.K = D; gp.N = PXR_TC_H1; gp.NT = 256;
  gp.wimg = h->tc_items_img[2]; gp.bias = pxr_tc_b1_padded(h);
  gp.mode = itc::OUT_16; gp.fmt16 = fmt16; gp.out_h = out; gp.ldo = PXR_TC_H1; gp.qt_nm = h->M - 1;
  return itc::launch_gemm(h, gp, st);
}

// K1 + K2 for n_rows items -> feats_out [rows][M-1][D] fp32 (same record as pxr_launch_items_simt)
int pxr_launch_items_tc(pxr_handle* h, const float* item_embedding, const int64_t* item_idx, const int64_t* tag_idx,
                        const float* vis, const float* txt, const float* num, int64_t n_rows, int64_t item_base,
                        float* feats_out, cudaStream_t st) {
  const pxr_config& c = h->cfg;
  const int D = c.embedding_dim, FD = (h->M - 1) * D;
  const float* ins[2] = {vis, txt};
  const int kdim[2] = {c.vision_dim, c.language_dim};
  if ((h->has_mod[0] && (((uintptr_t)vis) & 15)) || (h->has_mod[1] && (((uintptr_t)txt) & 15)))
    PXR_FAIL(h, PXR_ERR_INVALID, "feature matrices must be 16-byte aligned");
  int slot = 2;
  for (int m = 0; m < 2; ++m) {
    if (!h->has_mod[m]) continue;
    if (!ins[m]) PXR_FAIL(h, PXR_ERR_INVALID, "modality %d is configured but its feature pointer is NULL", m);
    itc::GemmParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.A = ins[m]; gp.lda = kdim[m]; gp.M = n_rows; gp.K = kdim[m]; gp.N = D; gp.NT = itc::n_tile_for(D);
    gp.wimg = h->tc_items_img[m]; gp.bias = h->proj[m][0].b;
    gp.mode = itc::OUT_F32_ACT; gp.act = c.activation; gp.out_f = feats_out + slot * D; gp.ldo = FD;
    int rc = itc::launch_gemm(h, gp, st);
    if (rc) return rc;
    slot++;
  }
  itc::GatherParams g;
  memset(&g, 0, sizeof(g));
  g.item_embedding = item_embedding; g.item_idx = item_idx; g.item_base = item_base;
  g.tag_emb = h->tag_emb; g.tag_idx = tag_idx;
  if (h->has_mod[2]) {
    if (!num) PXR_FAIL(h, PXR_ERR_INVALID, "numerical features are configured but the pointer is NULL");
    g.num = num; g.num_dim = c.num_numerical; g.num_wt = h->proj[2][0].wt; g.num_b = h->proj[2][0].b; g.num_slot = slot;
  }
  g.act = c.activation; g.feats = feats_out; g.FD = FD; g.D = D; g.n_rows = n_rows;
  g.rows = std::max(1, std::min(32, (40 * 1024) / (2 * D * 4)));
  itc::gather_small_kernel<<<(unsigned)((n_rows + g.rows - 1) / g.rows), 128, (size_t)g.rows * 2 * D * 4, st>>>(g);
  h->launches++;
  PXR_CUDA(h, cudaGetLastError());
  return PXR_OK;
}

// concat: item partial of layer 1, Pi[row] = W1[:, D:] . record + b1 -> 16 bit  
int pxr_launch_item_pi_tc(pxr_handle* h, int64_t n_rows, uint16_t* out, int fmt16, cudaStream_t st) {
  const pxr_config& c = h->cfg;
  const int D = c.embedding_dim, FD = (h->M - 1) * D;
  itc::GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.A = h->item_feats; gp.lda = FD; gp.M = n_rows; gp.K = FD; gp.N = PXR_TC_H1; gp.NT = 256;
  gp.wimg = h->tc_items_img[2]; gp.bias = pxr_tc_b1_padded(h);
  gp.mode = itc::OUT_16; gp.fmt16 = fmt16; gp.out_h = out; gp.ldo = PXR_TC_H1;
  return itc::launch_gemm(h, gp, st);
}
