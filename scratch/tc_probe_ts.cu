// Probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (TS form), written by tcgen05.st (scratch).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__global__ void probe(float* out, int Npad, int a_col) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  float* Bm = (float*)smem;                // K-major: plane[k/4][o][k%4], K = 16 (two k-steps)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16 * Npad; i += blockDim.x) { int k = i / Npad, o = i % Npad; Bm[(k / 4) * (Npad * 4) + o * 4 + (k % 4)] = 16.f * k + o + 1.f; }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  // A[r][k] = 1 if k == r % 16 else 0, 16 columns at a_col, written row-wise by 4 warps
  {
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < 16; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(((c0 + j) == (r % 16)) ? 1.f : 0.f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tb + ((uint32_t)(warp * 32) << 16) + a_col + c0),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Npad >> 3) << 17) | (8u << 24);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t db = mkdesc(s32(Bm) + ks * 2 * Npad * 16, Npad * 16, 128);
      const uint32_t ta = tb + a_col + ks * 8;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tb), "r"(ta), "l"(db), "r"(idesc), "r"((uint32_t)ks) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < Npad; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * Npad + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}
int main() {
  const int Npad = 32;
  float* d; cudaMalloc(&d, 128 * Npad * 4);
  float* h = (float*)malloc(128 * Npad * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int a_col = 64; a_col <= 160; a_col += 96) {
    cudaMemset(d, 0xff, 128 * Npad * 4);
    probe<<<1, 128, 65536>>>(d, Npad, a_col);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("a_col %d: %s\n", a_col, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 128 * Npad * 4, cudaMemcpyDeviceToHost);
    int ok = 0;
    for (int r = 0; r < 128; ++r) for (int o = 0; o < Npad; ++o) ok += (h[r * Npad + o] == 16.f * (r % 16) + o + 1.f);
    printf("A in TMEM at column %d: %d / %d correct; row0: %g %g %g | row9: %g %g | row17: %g %g\n", a_col, ok, 128 * Npad, h[0], h[1], h[2],
           h[9 * Npad], h[9 * Npad + 1], h[17 * Npad], h[17 * Npad + 1]);
  }
  return 0;
}
