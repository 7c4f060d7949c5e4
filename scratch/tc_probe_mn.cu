// Probe of tcgen05.mma smem-descriptor conventions for MN-major operands without swizzle (scratch; not part of the library).
// Question: can ONE plane layout  plane[f / T][row][f % T]  (T = 16 B / sizeof(elem)) serve BOTH as the K-major operand of
// the forward / data-gradient contractions (MN = row, K = f) and as the MN-major operand of the weight-gradient
// contraction (MN = f, K = row)?  CUTLASS (cute/atom/mma_traits_sm100.hpp) says, in 16-byte units, no swizzle:
//   K-major : ((8,n),2):((1,SBO),LBO)     MN-major: ((1,n),(8,k)):((X,SBO),(1,LBO))
// Test: D[M=128][N] = A^T-stored [M x K] * B-stored [N x K] with K = rows (64), both operands MN-major, bf16 and tf32.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         (1ull << 46);
}
// variant bit 0: swap LBO/SBO of A; bit 1: swap LBO/SBO of B; bit 2: A K-major control (A stored as row planes over k)
// kind: 0 = bf16 (K = 16 per MMA), 1 = tf32 (K = 8 per MMA)
template <int KIND>
__global__ void probe(float* out, const float* Ain, const float* Bin, int variant, int Nn, int Kr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tbase;
  constexpr int T = KIND == 0 ? 8 : 4;          // elements per 16 bytes
  constexpr int ES = KIND == 0 ? 2 : 4;
  constexpr int KSTEP = KIND == 0 ? 16 : 8;
  const int M = 128;
  // operand X^T for the wgrad form: logical A[m = feature][k = row]; stored as plane[f / T][row][f % T], Kr rows per plane
  uint8_t* As = smem;
  const int planeA = Kr * 16;                   // bytes per plane
  uint8_t* Bs = smem + (M / T) * planeA;
  const int planeB = Kr * 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * Kr; i += blockDim.x) {
    const int f = i / Kr, r = i % Kr;
    const float v = Ain[f * Kr + r];
    uint8_t* p = As + (f / T) * planeA + r * 16 + (f % T) * ES;
    if (KIND == 0) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v); else *reinterpret_cast<float*>(p) = v;
  }
  for (int i = tid; i < Nn * Kr; i += blockDim.x) {
    const int f = i / Kr, r = i % Kr;
    const float v = Bin[f * Kr + r];
    uint8_t* p = Bs + (f / T) * planeB + r * 16 + (f % T) * ES;
    if (KIND == 0) *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v); else *reinterpret_cast<float*>(p) = v;
  }
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
  if (tid == 0) {
    // MN-major, no swizzle: SBO = stride between 16-byte groups along MN (= one plane), LBO = stride between groups of
    // 8 k (= 8 rows x 16 B = 128 B)
    uint32_t a_lbo = 128, a_sbo = planeA, b_lbo = 128, b_sbo = planeB;
    if (variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
    if (variant & 2) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t fmt = KIND == 0 ? 1u : 2u;   // bf16 / tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int ksteps = Kr / KSTEP;
    for (int ks = 0; ks < ksteps; ++ks) {
      // advancing along K = rows: KSTEP rows x 16 B per plane
      const uint64_t da = mkdesc(s32(As) + ks * KSTEP * 16, a_lbo, a_sbo), db = mkdesc(s32(Bs) + ks * KSTEP * 16, b_lbo, b_sbo);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    for (int c0 = 0; c0 < Nn; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tb + ((uint32_t)(warp * 32) << 16) + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * Nn + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float tf32t(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int M = 128, Nn = 48, Kr = 64;
  float *hA = (float*)malloc(M * Kr * 4), *hB = (float*)malloc(Nn * Kr * 4), *h = (float*)malloc(M * Nn * 4);
  srand(1);
  for (int i = 0; i < M * Kr; ++i) hA[i] = (float)((rand() % 17) - 8) / 4.f;     // exactly representable in bf16
  for (int i = 0; i < Nn * Kr; ++i) hB[i] = (float)((rand() % 13) - 6) / 2.f;
  float *dA, *dB, *d;
  cudaMalloc(&dA, M * Kr * 4); cudaMalloc(&dB, Nn * Kr * 4); cudaMalloc(&d, M * Nn * 4);
  cudaMemcpy(dA, hA, M * Kr * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, Nn * Kr * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int kind = 0; kind < 2; ++kind)
    for (int variant = 0; variant < 4; ++variant) {
      cudaMemset(d, 0xff, M * Nn * 4);
      if (kind == 0) probe<0><<<1, 128, 65536>>>(d, dA, dB, variant, Nn, Kr); else probe<1><<<1, 128, 65536>>>(d, dA, dB, variant, Nn, Kr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kind %d variant %d: %s\n", kind, variant, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, M * Nn * 4, cudaMemcpyDeviceToHost);
      int ok = 0, nz = 0;
      double maxerr = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < Nn; ++n) {
          double want = 0;
          for (int k = 0; k < Kr; ++k) want += (double)hA[m * Kr + k] * hB[n * Kr + k];
          const double err = fabs(h[m * Nn + n] - want);
          ok += err < 1e-3; nz += h[m * Nn + n] != 0.f;
          if (err > maxerr) maxerr = err;
        }
      printf("kind %s variant %d (A lbo/sbo swapped %d, B swapped %d): %5d / %d correct, %d non-zero, max err %g; D[0][0..3] = %g %g %g %g, D[1][0] = %g\n",
             kind == 0 ? "bf16" : "tf32", variant, variant & 1, (variant >> 1) & 1, ok, M * Nn, nz, maxerr, h[0], h[1], h[2], h[3], h[Nn]);
    }
  return 0;
}
