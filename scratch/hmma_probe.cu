// Probe: latency and throughput of the register-operand tensor-core instruction (mma.sync m16n8k8 TF32 -> HMMA.1688.F32.TF32)
// on sm_100a, per warp and per SM (1..16 warps, 1..8 independent accumulator chains per warp).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CH>
__global__ void probe(float* out, long long* cyc, int iters) {
  float c[CH][4];
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
#pragma unroll
  for (int j = 0; j < CH; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CH; ++j) mma_tf32(c[j], a, (uint32_t)i, (uint32_t)j);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  probe<CH><<<148, warps * 32>>>(out, cyc, iters);
  probe<CH><<<148, warps * 32>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)iters * CH);
  printf("warps/SM %2d chains %d: %.1f cycles per HMMA per warp, %.2f cycles per HMMA per SM\n", warps, CH, per, per / warps);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 12, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
