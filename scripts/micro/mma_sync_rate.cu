// Micro-benchmark: issue rate of legacy warp-level mma.sync.m16n8k16 (bf16, fp32 accumulate) on sm_100a, per SM,
// for 1..16 warps per SM with 1/2/4/8 independent accumulator chains per warp.  Decides whether the attention front
// end can run its small per-item GEMMs on register-level MMAs next to the tcgen05 MLP chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma1688_tf32(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void ktf32(float* out, int iters, long long* cycles) {
  float c[CHAINS][4];
  uint32_t a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3c000000u + threadIdx.x, 0x3c000000u};
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma1688_tf32(c[j], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
__global__ void k(float* out, int iters, long long* cycles) {
  float c[CHAINS][4];
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3c003c00u + threadIdx.x, 0x3c003c00u};
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma16816(c[j], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void run(int warps, float* out, long long* cyc) {
  const int iters = 4096;
  k<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double mmas_per_sm = (double)iters * CHAINS * warps;
  printf("warps/SM %2d chains %d: %.1f cycles per MMA per SM (%.3f MMA/clk/SM), chip %.1f TFLOP/s dense bf16, dependent-chain latency %.1f clk\n",
         warps, CHAINS, (double)h[0] / mmas_per_sm, mmas_per_sm / (double)h[0], 148.0 * mmas_per_sm * 4096.0 / (ms * 1e-3) / 1e12,
         (double)h[0] / iters);
}

template <int CHAINS>
void run_tf32(int warps, float* out, long long* cyc) {
  const int iters = 4096;
  ktf32<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ktf32<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double mmas_per_sm = (double)iters * CHAINS * warps;
  printf("tf32 m16n8k8  warps/SM %2d chains %d: %.1f cycles per MMA per SM, chip %.1f TFLOP/s dense tf32 (x1/3 for 3xTF32), chain latency %.1f clk\n",
         warps, CHAINS, (double)h[0] / mmas_per_sm, 148.0 * mmas_per_sm * 2048.0 / (ms * 1e-3) / 1e12, (double)h[0] / iters);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int w : {1, 4, 8, 16}) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<4>(w, out, cyc); run<8>(w, out, cyc); }
  for (int w : {4, 8, 16}) { run_tf32<4>(w, out, cyc); run_tf32<8>(w, out, cyc); }
  return 0;
}
