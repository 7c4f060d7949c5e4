// Standalone probe of tcgen05.mma kind::tf32 descriptor conventions (scratch; not part of the library).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int version) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         ((uint64_t)version << 46);
}
// variant bits: 1 = swap LBO/SBO of A, 2 = swap LBO/SBO of B, 4 = B K-major (B stored [n][k] planes), 8 = version 0
__global__ void probe(float* out, int variant, int Npad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  float* A = (float*)smem;                 // 2 planes [128][4]
  float* Bm = A + 2 * 512;                 // MN-major planes over o: plane[o/4][k(8)][o%4]   (Npad/4 planes of 32 floats)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < 128 * 8; i += blockDim.x) { int r = i / 8, k = i % 8; A[(k / 4) * 512 + r * 4 + (k % 4)] = (k == (r % 8)) ? 1.f : 0.f; }
  for (int i = tid; i < 8 * Npad; i += blockDim.x) {
    int k = i / Npad, o = i % Npad; float v = 16.f * k + o + 1.f;
    if (variant & 4) Bm[(k / 4) * (Npad * 4) + o * 4 + (k % 4)] = v;      // K-major: plane[k/4][o][k%4]
    else Bm[(o / 4) * 32 + k * 4 + (o % 4)] = v;
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  if (tid == ((variant & 32) ? 32 : 0)) {
    uint32_t a_lbo = 2048, a_sbo = 128;
    if (variant & 1) { a_lbo = 128; a_sbo = 2048; }
    uint32_t b_lbo, b_sbo;
    if (variant & 4) { b_lbo = Npad * 16; b_sbo = 128; } else { b_lbo = 128; b_sbo = 128; /* plane stride = 8 rows*16B = 128 */ }
    if (variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const int ver = (variant & 8) ? 0 : 1;
    const uint64_t da = mkdesc(s32(A), a_lbo, a_sbo, ver), db = mkdesc(s32(Bm), b_lbo, b_sbo, ver);
    const uint32_t bmaj = (variant & 4) ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (bmaj << 16) | ((uint32_t)(Npad >> 3) << 17) | (8u << 24);
    if (variant & 16) {
      uint32_t z = 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(z), "r"(z), "r"(z), "r"(z) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    if (variant == 0) printf("tbase=%u idesc=%08x da=%016llx db=%016llx\n", tb, idesc, (unsigned long long)da, (unsigned long long)db);
  }
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int c0 = 0; c0 < Npad; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * Npad + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128u) : "memory");
}
int main(int argc, char** argv) {
  const int Npad = 16;
  float* d; cudaMalloc(&d, 128 * Npad * 4);
  float* h = (float*)malloc(128 * Npad * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int variant = 0; variant < 64; ++variant) {
    if ((variant & 8)) continue;
    cudaMemset(d, 0xff, 128 * Npad * 4);
    probe<<<1, 128, 65536>>>(d, variant, Npad);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 128 * Npad * 4, cudaMemcpyDeviceToHost);
    int ok = 0, nz = 0;
    for (int r = 0; r < 128; ++r) for (int o = 0; o < Npad; ++o) { float want = 16.f * (r % 8) + o + 1.f; ok += (h[r * Npad + o] == want); nz += (h[r * Npad + o] != 0.f); }
    printf("variant %2d (A swap %d, B swap %d, B K-major %d, ver0 %d): %4d / %d correct, %d non-zero; row0: ", variant, variant & 1, (variant >> 1) & 1,
           (variant >> 2) & 1, (variant >> 3) & 1, ok, 128 * Npad, nz);
    for (int o = 0; o < 8; ++o) printf("%g ", h[o]);
    printf("| row1: "); for (int o = 0; o < 8; ++o) printf("%g ", h[Npad + o]);
    printf("| row9: "); for (int o = 0; o < 4; ++o) printf("%g ", h[9 * Npad + o]);
    printf("\n");
  }
  return 0;
}
