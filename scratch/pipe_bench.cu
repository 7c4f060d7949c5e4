// Throughput probes: scalar FFMA vs packed FFMA2 vs mma.sync tf32 (planning data, not product code)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void ffma2_kernel(float* out, int iters, float a, float b) {
  float2 acc[ILP];
  float2 aa = make_float2(a, a + 1e-3f), bb = make_float2(b, b * 0.5f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x + i, i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = __ffma2_rn(acc[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void mma_tf32_kernel(float* out, int iters) {
  float c[ILP][4];
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3f800000u, 0x3e800000u};
  unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = 0; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void mma_bf16_kernel(float* out, int iters) {
  float c[ILP][4];
  unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f003f00u, 0x3f803f80u, 0x3e803e80u};
  unsigned b[2] = {0x3f803f80u, 0x3f003f00u + threadIdx.x};
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = 0; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float run(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 4));
  int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int thr = warps * 32, grid = 148;
    float ms = run([&] { ffma_kernel<16><<<grid, thr>>>(out, iters, 1.0001f, 0.5f); });
    double fl = 2.0 * grid * thr * 16.0 * iters;
    printf("FFMA   warps/SM=%2d: %.1f TFLOP/s  (%.2f FMA-warp-instr/clk/SM @1.9GHz)\n", warps, fl / ms / 1e9, fl / 2 / 32 / (ms * 1e-3) / 148 / 1.9e9);
    ms = run([&] { ffma2_kernel<16><<<grid, thr>>>(out, iters, 1.0001f, 0.5f); });
    fl = 4.0 * grid * thr * 16.0 * iters;
    printf("FFMA2  warps/SM=%2d: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
    ms = run([&] { mma_tf32_kernel<8><<<grid, thr>>>(out, iters / 4); });
    fl = 2.0 * 16 * 8 * 8 * (double)grid * warps * 8.0 * (iters / 4);
    printf("MMA.tf32 m16n8k8 warps/SM=%2d: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
    ms = run([&] { mma_bf16_kernel<8><<<grid, thr>>>(out, iters / 4); });
    fl = 2.0 * 16 * 8 * 16 * (double)grid * warps * 8.0 * (iters / 4);
    printf("MMA.bf16 m16n8k16 warps/SM=%2d: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
  }
  return 0;
}
