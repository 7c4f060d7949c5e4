// Probe: how fast can ONE CTA per SM (384 threads, operands in shared memory, true fp32 on the FP32 pipe) run the GEMM
// shapes of the fused training kernel's tile (140 rows, feature-major arena)?  Variants of the inner loop:
//   fwd   out[o][r] = sum_k W[k][o] * X[k][r]            (gemm_phase)
//   dgrad dX[k][r]  = sum_o W[k][o] * dZ[o][r]           (bwd_phase, data gradient)
//   wgrad dW[k][o]  = sum_r X[k][r] * dZ[o][r]           (wgrad_block)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scratch/gemm_probe scratch/gemm_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#ifndef T_THREADS
#define T_THREADS 384
#endif
constexpr int T = T_THREADS, RP = 140, RG = RP / 4;

__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long d, a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// ---- forward, baseline structure of fused.cu (k in steps of 4, optional row table)
template <int TC, bool TABLE, bool F2>
__device__ __forceinline__ void fwd_base(const float* X, const float* W, float* OUT, const int* tab, int K, int O, int tid) {
  const int items = RG * (O / TC);
  const int K4 = (K + 3) >> 2;
  for (int item = tid; item < items; item += T) {
    const int cg = item / RG, rg = item - cg * RG;
    const int r0 = rg * 4, o0 = cg * TC;
    float acc[TC][4];
#pragma unroll
    for (int j = 0; j < TC; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    const float* wp = W + o0;
    const float* xb = X + r0;
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      float4 x[4];
      if (TABLE) {
        const int4 rows = reinterpret_cast<const int4*>(tab)[k4];
        x[0] = *reinterpret_cast<const float4*>(xb + rows.x * RP); x[1] = *reinterpret_cast<const float4*>(xb + rows.y * RP);
        x[2] = *reinterpret_cast<const float4*>(xb + rows.z * RP); x[3] = *reinterpret_cast<const float4*>(xb + rows.w * RP);
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) x[kk] = *reinterpret_cast<const float4*>(xb + (k4 * 4 + kk) * RP);
      }
      float w[4][TC];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j4 = 0; j4 < TC; j4 += 4) {
          const float4 t = *reinterpret_cast<const float4*>(wp + kk * O + j4);
          w[kk][j4] = t.x; w[kk][j4 + 1] = t.y; w[kk][j4 + 2] = t.z; w[kk][j4 + 3] = t.w;
        }
      wp += 4 * O;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < TC; ++j) {
          if (F2) {
            fma2(acc[j][0], acc[j][1], x[kk].x, x[kk].y, w[kk][j], w[kk][j]);
            fma2(acc[j][2], acc[j][3], x[kk].z, x[kk].w, w[kk][j], w[kk][j]);
          } else {
            acc[j][0] = fmaf(x[kk].x, w[kk][j], acc[j][0]); acc[j][1] = fmaf(x[kk].y, w[kk][j], acc[j][1]);
            acc[j][2] = fmaf(x[kk].z, w[kk][j], acc[j][2]); acc[j][3] = fmaf(x[kk].w, w[kk][j], acc[j][3]);
          }
        }
    }
#pragma unroll
    for (int j = 0; j < TC; ++j)
      *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0) = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
  }
}

// ---- forward, explicit register double buffering: operands of step k+1 are loaded before the FFMAs of step k (k in steps of KS)
template <int TC, int KS, bool F2>
__device__ __forceinline__ void fwd_pipe(const float* X, const float* W, float* OUT, int K, int O, int tid) {
  const int items = RG * (O / TC);
  const int KN = K / KS;                       // K must be a multiple of KS here
  for (int item = tid; item < items; item += T) {
    const int cg = item / RG, rg = item - cg * RG;
    const int r0 = rg * 4, o0 = cg * TC;
    float acc[TC][4];
#pragma unroll
    for (int j = 0; j < TC; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    const float* wp = W + o0;
    const float* xp = X + r0;
    float4 x[2][KS];
    float4 w[2][KS][TC / 4];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      x[0][kk] = *reinterpret_cast<const float4*>(xp + kk * RP);
#pragma unroll
      for (int j4 = 0; j4 < TC / 4; ++j4) w[0][kk][j4] = *reinterpret_cast<const float4*>(wp + kk * O + j4 * 4);
    }
#pragma unroll 2
    for (int ks = 0; ks < KN; ++ks) {
      const int cur = ks & 1, nxt = cur ^ 1;
      if (ks + 1 < KN) {
        xp += KS * RP; wp += KS * O;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          x[nxt][kk] = *reinterpret_cast<const float4*>(xp + kk * RP);
#pragma unroll
          for (int j4 = 0; j4 < TC / 4; ++j4) w[nxt][kk][j4] = *reinterpret_cast<const float4*>(wp + kk * O + j4 * 4);
        }
      }
#pragma unroll
      for (int kk = 0; kk < KS; ++kk)
#pragma unroll
        for (int j4 = 0; j4 < TC / 4; ++j4) {
          const float4 wv = w[cur][kk][j4];
          const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j4 * 4 + jj;
            if (F2) {
              fma2(acc[j][0], acc[j][1], x[cur][kk].x, x[cur][kk].y, ws[jj], ws[jj]);
              fma2(acc[j][2], acc[j][3], x[cur][kk].z, x[cur][kk].w, ws[jj], ws[jj]);
            } else {
              acc[j][0] = fmaf(x[cur][kk].x, ws[jj], acc[j][0]); acc[j][1] = fmaf(x[cur][kk].y, ws[jj], acc[j][1]);
              acc[j][2] = fmaf(x[cur][kk].z, ws[jj], acc[j][2]); acc[j][3] = fmaf(x[cur][kk].w, ws[jj], acc[j][3]);
            }
          }
        }
    }
#pragma unroll
    for (int j = 0; j < TC; ++j)
      *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0) = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
  }
}

// ---- forward, 8 rows x TC columns per thread (two row vectors): half the weight loads and index math per FFMA
template <int TC, bool F2>
__device__ __forceinline__ void fwd_r8(const float* X, const float* W, float* OUT, int K, int O, int tid) {
  constexpr int RG8 = (RP + 7) / 8;            // 18 groups; the last one is half empty (rows 136..139 + 4 pad rows of the arena)
  const int items = RG8 * (O / TC);
  for (int item = tid; item < items; item += T) {
    const int cg = item / RG8, rg = item - cg * RG8;
    const int r0 = rg * 8, o0 = cg * TC;
    const bool two = r0 + 4 < RP;
    float acc[TC][8];
#pragma unroll
    for (int j = 0; j < TC; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    const float* wp = W + o0;
    const float* xp = X + r0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 xa = *reinterpret_cast<const float4*>(xp + k * RP);
      const float4 xb = two ? *reinterpret_cast<const float4*>(xp + k * RP + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float w[TC];
#pragma unroll
      for (int j4 = 0; j4 < TC; j4 += 4) {
        const float4 t = *reinterpret_cast<const float4*>(wp + k * O + j4);
        w[j4] = t.x; w[j4 + 1] = t.y; w[j4 + 2] = t.z; w[j4 + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TC; ++j) {
        if (F2) {
          fma2(acc[j][0], acc[j][1], xa.x, xa.y, w[j], w[j]); fma2(acc[j][2], acc[j][3], xa.z, xa.w, w[j], w[j]);
          fma2(acc[j][4], acc[j][5], xb.x, xb.y, w[j], w[j]); fma2(acc[j][6], acc[j][7], xb.z, xb.w, w[j], w[j]);
        } else {
          acc[j][0] = fmaf(xa.x, w[j], acc[j][0]); acc[j][1] = fmaf(xa.y, w[j], acc[j][1]);
          acc[j][2] = fmaf(xa.z, w[j], acc[j][2]); acc[j][3] = fmaf(xa.w, w[j], acc[j][3]);
          acc[j][4] = fmaf(xb.x, w[j], acc[j][4]); acc[j][5] = fmaf(xb.y, w[j], acc[j][5]);
          acc[j][6] = fmaf(xb.z, w[j], acc[j][6]); acc[j][7] = fmaf(xb.w, w[j], acc[j][7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0) = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
      if (two) *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0 + 4) = make_float4(fmaxf(acc[j][4], 0.f), fmaxf(acc[j][5], 0.f), fmaxf(acc[j][6], 0.f), fmaxf(acc[j][7], 0.f));
    }
  }
}

// ---- weight gradient: KB x OB blocks of 4x4, all rows (baseline), static assignment block = tid
template <bool F2>
__device__ __forceinline__ void wgrad_base(const float* X, const float* DZ, float* DW, int K, int O, int tid) {
  const int OB = O >> 2, KB = (K + 3) >> 2;
  for (int blk = tid; blk < KB * OB; blk += T) {
    const int kb = blk / OB, ob = blk - kb * OB;
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    const float* xr[4]; const float* dr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { xr[i] = X + (kb * 4 + i) * RP; dr[i] = DZ + (ob + i * OB) * RP; }
#pragma unroll 2
    for (int r = 0; r < RP; r += 4) {
      float4 xv[4], dv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { xv[i] = *reinterpret_cast<const float4*>(xr[i] + r); dv[i] = *reinterpret_cast<const float4*>(dr[i] + r); }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (F2) {
            float t0 = 0.f, t1 = 0.f;
            fma2(a[i * 4 + j], t0, xv[i].x, xv[i].y, dv[j].x, dv[j].y);   // not bit-compatible with the scalar order: probe only
            fma2(a[i * 4 + j], t1, xv[i].z, xv[i].w, dv[j].z, dv[j].w);
            a[i * 4 + j] += t0 + t1;
          } else {
            a[i * 4 + j] = fmaf(xv[i].x, dv[j].x, a[i * 4 + j]); a[i * 4 + j] = fmaf(xv[i].y, dv[j].y, a[i * 4 + j]);
            a[i * 4 + j] = fmaf(xv[i].z, dv[j].z, a[i * 4 + j]); a[i * 4 + j] = fmaf(xv[i].w, dv[j].w, a[i * 4 + j]);
          }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) DW[(kb * 4 + i) * O + ob + j * OB] = a[i * 4 + j];
  }
}

// ---- weight gradient with the rows split over RS lanes (every thread busy), shuffle reduction; block BK x BO (4x4 or 8x4 ...)
template <int BK, int BO, int RS>
__device__ __forceinline__ void wgrad_split(const float* X, const float* DZ, float* DW, int K, int O, int tid) {
  const int OB = O / BO, KB = (K + BK - 1) / BK;
  const int tasks = KB * OB * RS;
  for (int task = tid; task < ((tasks + 31) & ~31); task += T) {
    const int blk = task / RS, split = task - blk * RS;
    const int kb = blk / OB, ob = blk - kb * OB;
    float a[BK][BO];
#pragma unroll
    for (int i = 0; i < BK; ++i)
#pragma unroll
      for (int j = 0; j < BO; ++j) a[i][j] = 0.f;
    if (task < tasks) {
#pragma unroll 1
      for (int r = split * 4; r < RP; r += RS * 4) {
        float4 xv[BK], dv[BO];
#pragma unroll
        for (int i = 0; i < BK; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (kb * BK + i) * RP + r);
#pragma unroll
        for (int j = 0; j < BO; ++j) dv[j] = *reinterpret_cast<const float4*>(DZ + (ob + j * OB) * RP + r);
#pragma unroll
        for (int i = 0; i < BK; ++i)
#pragma unroll
          for (int j = 0; j < BO; ++j) {
            a[i][j] = fmaf(xv[i].x, dv[j].x, a[i][j]); a[i][j] = fmaf(xv[i].y, dv[j].y, a[i][j]);
            a[i][j] = fmaf(xv[i].z, dv[j].z, a[i][j]); a[i][j] = fmaf(xv[i].w, dv[j].w, a[i][j]);
          }
      }
    }
    for (int off = RS >> 1; off > 0; off >>= 1)
#pragma unroll
      for (int i = 0; i < BK; ++i)
#pragma unroll
        for (int j = 0; j < BO; ++j) a[i][j] += __shfl_xor_sync(0xffffffffu, a[i][j], off);
    if (task < tasks && split == 0)
#pragma unroll
      for (int i = 0; i < BK; ++i)
#pragma unroll
        for (int j = 0; j < BO; ++j) DW[(kb * BK + i) * O + ob + j * OB] = a[i][j];
  }
}

// ---- data gradient: 4 input columns x 4 rows per item, o in steps of 4 (baseline, static item = tid)
template <bool F2>
__device__ __forceinline__ void dgrad_base(const float* DZ, const float* W, float* DX, int K, int O, int tid) {
  const int items = RG * (K / 4);
  for (int item = tid; item < items; item += T) {
    const int kg = item / RG, rg = item - kg * RG;
    const int r0 = rg * 4;
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wr = W + kg * 4 * O;
    const float* ab = DZ + r0;
#pragma unroll 2
    for (int o = 0; o < O; o += 4) {
      float4 dz[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { dz[i] = *reinterpret_cast<const float4*>(ab + (o + i) * RP); w[i] = *reinterpret_cast<const float4*>(wr + i * O + o); }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ws[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
        for (int oo = 0; oo < 4; ++oo) {
          if (F2) {
            fma2(acc[i].x, acc[i].y, dz[oo].x, dz[oo].y, ws[oo], ws[oo]); fma2(acc[i].z, acc[i].w, dz[oo].z, dz[oo].w, ws[oo], ws[oo]);
          } else {
            acc[i].x = fmaf(dz[oo].x, ws[oo], acc[i].x); acc[i].y = fmaf(dz[oo].y, ws[oo], acc[i].y);
            acc[i].z = fmaf(dz[oo].z, ws[oo], acc[i].z); acc[i].w = fmaf(dz[oo].w, ws[oo], acc[i].w);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(DX + (kg * 4 + i) * RP + r0) = acc[i];
  }
}
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void fma2p(unsigned long long& d, unsigned long long a, unsigned long long b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ float sum2(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// ---- weight gradient with packed accumulators: lane pair = (even row, odd row) of the sum over rows, no duplicated operand;
//      BK x BO block per thread, rows split over RS adjacent lanes when RS > 1
template <int BK, int BO, int RS>
__device__ __forceinline__ void wgrad_f2(const float* X, const float* DZ, float* DW, int K, int O, int tid) {
  const int OB = O / BO, KB = (K + BK - 1) / BK;
  const int tasks = KB * OB * RS;
  for (int task = tid; task < ((tasks + 31) & ~31); task += T) {
    const int blk = task / RS, split = task - blk * RS;
    const int kb = blk / OB, ob = blk - kb * OB;
    unsigned long long a[BK][BO];
#pragma unroll
    for (int i = 0; i < BK; ++i)
#pragma unroll
      for (int j = 0; j < BO; ++j) a[i][j] = 0ull;
    if (task < tasks) {
#pragma unroll 2
      for (int r = split * 4; r < RP; r += RS * 4) {
        float4 xv[BK], dv[BO];
#pragma unroll
        for (int i = 0; i < BK; ++i) xv[i] = *reinterpret_cast<const float4*>(X + (kb * BK + i) * RP + r);
#pragma unroll
        for (int j = 0; j < BO; ++j) dv[j] = *reinterpret_cast<const float4*>(DZ + (ob + j * OB) * RP + r);
#pragma unroll
        for (int i = 0; i < BK; ++i)
#pragma unroll
          for (int j = 0; j < BO; ++j) {
            fma2p(a[i][j], pk(xv[i].x, xv[i].y), pk(dv[j].x, dv[j].y));
            fma2p(a[i][j], pk(xv[i].z, xv[i].w), pk(dv[j].z, dv[j].w));
          }
      }
    }
    float s[BK][BO];
#pragma unroll
    for (int i = 0; i < BK; ++i)
#pragma unroll
      for (int j = 0; j < BO; ++j) s[i][j] = sum2(a[i][j]);
    for (int off = RS >> 1; off > 0; off >>= 1)
#pragma unroll
      for (int i = 0; i < BK; ++i)
#pragma unroll
        for (int j = 0; j < BO; ++j) s[i][j] += __shfl_xor_sync(0xffffffffu, s[i][j], off);
    if (task < tasks && split == 0)
#pragma unroll
      for (int i = 0; i < BK; ++i)
#pragma unroll
        for (int j = 0; j < BO; ++j) DW[(kb * BK + i) * O + ob + j * OB] = s[i][j];
  }
}

// ---- data gradient: 8 input columns x 4 rows per item, packed along rows (weights duplicated by a MOV)

// ---- data gradient, 8 input columns x 4 rows per item
template <bool F2>
__device__ __forceinline__ void dgrad_k8(const float* DZ, const float* W, float* DX, int K, int O, int tid) {
  const int items = RG * (K / 8);
  for (int item = tid; item < items; item += T) {
    const int kg = item / RG, rg = item - kg * RG;
    const int r0 = rg * 4;
    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wr = W + kg * 8 * O;
    const float* ab = DZ + r0;
#pragma unroll 2
    for (int o = 0; o < O; o += 4) {
      float4 dz[4], w[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) dz[i] = *reinterpret_cast<const float4*>(ab + (o + i) * RP);
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = *reinterpret_cast<const float4*>(wr + i * O + o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ws[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
        for (int oo = 0; oo < 4; ++oo) {
          if (F2) {
            fma2(acc[i].x, acc[i].y, dz[oo].x, dz[oo].y, ws[oo], ws[oo]); fma2(acc[i].z, acc[i].w, dz[oo].z, dz[oo].w, ws[oo], ws[oo]);
          } else {
            acc[i].x = fmaf(dz[oo].x, ws[oo], acc[i].x); acc[i].y = fmaf(dz[oo].y, ws[oo], acc[i].y);
            acc[i].z = fmaf(dz[oo].z, ws[oo], acc[i].z); acc[i].w = fmaf(dz[oo].w, ws[oo], acc[i].w);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(DX + (kg * 8 + i) * RP + r0) = acc[i];
  }
}

// ---- forward, 8 rows x 8 columns per thread, K split over SK adjacent lanes (shuffle reduction): smem words per FFMA halve
template <int SK, bool F2>
__device__ __forceinline__ void fwd_8x8_sk(const float* X, const float* W, float* OUT, int K, int O, int tid) {
  constexpr int RG8 = (RP + 7) / 8;
  const int items = RG8 * (O / 8) * SK;
  const int KS = K / SK;
  for (int item = tid; item < ((items + 31) & ~31); item += T) {
    const int sk = item % SK, it = item / SK;
    const int cg = it / RG8, rg = it - cg * RG8;
    const int r0 = rg * 8, o0 = cg * 8;
    const bool live = item < items;
    const bool two = r0 + 4 < RP;
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    if (live) {
      const float* wp = W + o0 + sk * KS * O;
      const float* xp = X + r0 + sk * KS * RP;
#pragma unroll 2
      for (int k = 0; k < KS; ++k) {
        const float4 xa = *reinterpret_cast<const float4*>(xp + k * RP);
        const float4 xb = two ? *reinterpret_cast<const float4*>(xp + k * RP + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 wa = *reinterpret_cast<const float4*>(wp + k * O), wb = *reinterpret_cast<const float4*>(wp + k * O + 4);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (F2) {
            fma2(acc[j][0], acc[j][1], xa.x, xa.y, w[j], w[j]); fma2(acc[j][2], acc[j][3], xa.z, xa.w, w[j], w[j]);
            fma2(acc[j][4], acc[j][5], xb.x, xb.y, w[j], w[j]); fma2(acc[j][6], acc[j][7], xb.z, xb.w, w[j], w[j]);
          } else {
            acc[j][0] = fmaf(xa.x, w[j], acc[j][0]); acc[j][1] = fmaf(xa.y, w[j], acc[j][1]);
            acc[j][2] = fmaf(xa.z, w[j], acc[j][2]); acc[j][3] = fmaf(xa.w, w[j], acc[j][3]);
            acc[j][4] = fmaf(xb.x, w[j], acc[j][4]); acc[j][5] = fmaf(xb.y, w[j], acc[j][5]);
            acc[j][6] = fmaf(xb.z, w[j], acc[j][6]); acc[j][7] = fmaf(xb.w, w[j], acc[j][7]);
          }
        }
      }
    }
    // reduce over the SK lanes; lane sk ends up owning columns j with j % SK == sk (every lane stores its share)
#pragma unroll
    for (int off = 1; off < SK; off <<= 1)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] += __shfl_xor_sync(0xffffffffu, acc[j][i], off);
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j % SK == sk) {
          *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0) = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
          if (two) *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0 + 4) = make_float4(fmaxf(acc[j][4], 0.f), fmaxf(acc[j][5], 0.f), fmaxf(acc[j][6], 0.f), fmaxf(acc[j][7], 0.f));
        }
    }
  }
}

// ---- forward, 4 rows x 16 columns per thread, K split over SK adjacent lanes
template <int SK, bool F2>
__device__ __forceinline__ void fwd_4x16_sk(const float* X, const float* W, float* OUT, int K, int O, int tid) {
  const int items = RG * (O / 16) * SK;
  const int KS = K / SK;
  for (int item = tid; item < ((items + 31) & ~31); item += T) {
    const int sk = item % SK, it = item / SK;
    const int cg = it / RG, rg = it - cg * RG;
    const int r0 = rg * 4, o0 = cg * 16;
    const bool live = item < items;
    float acc[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    if (live) {
      const float* wp = W + o0 + sk * KS * O;
      const float* xp = X + r0 + sk * KS * RP;
#pragma unroll 2
      for (int k = 0; k < KS; ++k) {
        const float4 xa = *reinterpret_cast<const float4*>(xp + k * RP);
        float w[16];
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const float4 t = *reinterpret_cast<const float4*>(wp + k * O + j4);
          w[j4] = t.x; w[j4 + 1] = t.y; w[j4 + 2] = t.z; w[j4 + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (F2) {
            fma2(acc[j][0], acc[j][1], xa.x, xa.y, w[j], w[j]); fma2(acc[j][2], acc[j][3], xa.z, xa.w, w[j], w[j]);
          } else {
            acc[j][0] = fmaf(xa.x, w[j], acc[j][0]); acc[j][1] = fmaf(xa.y, w[j], acc[j][1]);
            acc[j][2] = fmaf(xa.z, w[j], acc[j][2]); acc[j][3] = fmaf(xa.w, w[j], acc[j][3]);
          }
        }
      }
    }
#pragma unroll
    for (int off = 1; off < SK; off <<= 1)
#pragma unroll
      for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] += __shfl_xor_sync(0xffffffffu, acc[j][i], off);
    if (live) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j % SK == sk)
          *reinterpret_cast<float4*>(OUT + (o0 + j) * RP + r0) = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
    }
  }
}

// ---- data gradient, 8 input columns x 8 rows per item, the sum over o split over SO adjacent lanes
template <int SO, bool F2>
__device__ __forceinline__ void dgrad_8x8_so(const float* DZ, const float* W, float* DX, int K, int O, int tid) {
  constexpr int RG8 = (RP + 7) / 8;
  const int items = RG8 * (K / 8) * SO;
  const int OS = O / SO;                       // multiple of 4
  for (int item = tid; item < ((items + 31) & ~31); item += T) {
    const int so = item % SO, it = item / SO;
    const int kg = it / RG8, rg = it - kg * RG8;
    const int r0 = rg * 8;
    const bool live = item < items;
    const bool two = r0 + 4 < RP;
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    if (live) {
      const float* wr = W + kg * 8 * O + so * OS;
      const float* ab = DZ + r0 + so * OS * RP;
#pragma unroll 1
      for (int o = 0; o < OS; o += 4) {
        float4 w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = *reinterpret_cast<const float4*>(wr + i * O + o);
#pragma unroll
        for (int oo = 0; oo < 4; ++oo) {
          const float4 za = *reinterpret_cast<const float4*>(ab + (o + oo) * RP);
          const float4 zb = two ? *reinterpret_cast<const float4*>(ab + (o + oo) * RP + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ws = oo == 0 ? w[i].x : oo == 1 ? w[i].y : oo == 2 ? w[i].z : w[i].w;
            if (F2) {
              fma2(acc[i][0], acc[i][1], za.x, za.y, ws, ws); fma2(acc[i][2], acc[i][3], za.z, za.w, ws, ws);
              fma2(acc[i][4], acc[i][5], zb.x, zb.y, ws, ws); fma2(acc[i][6], acc[i][7], zb.z, zb.w, ws, ws);
            } else {
              acc[i][0] = fmaf(za.x, ws, acc[i][0]); acc[i][1] = fmaf(za.y, ws, acc[i][1]);
              acc[i][2] = fmaf(za.z, ws, acc[i][2]); acc[i][3] = fmaf(za.w, ws, acc[i][3]);
              acc[i][4] = fmaf(zb.x, ws, acc[i][4]); acc[i][5] = fmaf(zb.y, ws, acc[i][5]);
              acc[i][6] = fmaf(zb.z, ws, acc[i][6]); acc[i][7] = fmaf(zb.w, ws, acc[i][7]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int off = 1; off < SO; off <<= 1)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] += __shfl_xor_sync(0xffffffffu, acc[j][i], off);
    if (live) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i % SO == so) {
          *reinterpret_cast<float4*>(DX + (kg * 8 + i) * RP + r0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          if (two) *reinterpret_cast<float4*>(DX + (kg * 8 + i) * RP + r0 + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
    }
  }
}

template <int variant>
__global__ void __launch_bounds__(T, 1) probe(int K, int O, int reps, long long* out, float* sink) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;                       // [96][RP] (+ 4 pad rows for the 8-row variant's tail)
  float* Z = X + 100 * RP;             // [96][RP]
  float* W = Z + 100 * RP;             // [96][96]
  int* tab = reinterpret_cast<int*>(W + 96 * 96);
  const int tid = threadIdx.x;
  for (int i = tid; i < 100 * RP; i += T) { X[i] = (float)((i * 7) % 13) * 0.01f - 0.05f; Z[i] = (float)((i * 5) % 11) * 0.01f - 0.04f; }
  for (int i = tid; i < 96 * 96; i += T) W[i] = (float)((i * 3) % 17) * 0.01f - 0.08f;
  for (int i = tid; i < 96; i += T) tab[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if constexpr (variant == 0) fwd_base<4, true, false>(X, W, Z, tab, K, O, tid);
    if constexpr (variant == 1) fwd_base<4, false, false>(X, W, Z, tab, K, O, tid);
    if constexpr (variant == 2) fwd_base<8, true, false>(X, W, Z, tab, K, O, tid);
    if constexpr (variant == 3) fwd_base<4, false, true>(X, W, Z, tab, K, O, tid);
    if constexpr (variant == 4) fwd_base<8, false, true>(X, W, Z, tab, K, O, tid);
    if constexpr (variant == 5) fwd_pipe<4, 4, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 6) fwd_pipe<4, 2, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 7) fwd_pipe<8, 2, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 8) fwd_pipe<4, 4, true>(X, W, Z, K, O, tid);
    if constexpr (variant == 9) fwd_r8<4, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 10) fwd_r8<4, true>(X, W, Z, K, O, tid);
    if constexpr (variant == 11) fwd_r8<8, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 12) fwd_r8<8, true>(X, W, Z, K, O, tid);
    if constexpr (variant == 13) fwd_8x8_sk<2, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 14) fwd_8x8_sk<2, true>(X, W, Z, K, O, tid);
    if constexpr (variant == 15) fwd_8x8_sk<4, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 16) fwd_8x8_sk<4, true>(X, W, Z, K, O, tid);
    if constexpr (variant == 17) fwd_4x16_sk<2, false>(X, W, Z, K, O, tid);
    if constexpr (variant == 18) fwd_4x16_sk<2, true>(X, W, Z, K, O, tid);
    if constexpr (variant == 34) dgrad_8x8_so<2, false>(Z, W, X, K, O, tid);
    if constexpr (variant == 35) dgrad_8x8_so<2, true>(Z, W, X, K, O, tid);
    if constexpr (variant == 36) dgrad_8x8_so<4, false>(Z, W, X, K, O, tid);
    if constexpr (variant == 37) dgrad_8x8_so<4, true>(Z, W, X, K, O, tid);
    if constexpr (variant == 20) wgrad_base<false>(X, Z, W, K, O, tid);
    if constexpr (variant == 21) wgrad_base<true>(X, Z, W, K, O, tid);
    if constexpr (variant == 22) wgrad_split<4, 4, 2>(X, Z, W, K, O, tid);
    if constexpr (variant == 23) wgrad_split<8, 4, 4>(X, Z, W, K, O, tid);
    if constexpr (variant == 24) wgrad_split<8, 8, 8>(X, Z, W, K, O, tid);
    if constexpr (variant == 25) wgrad_split<8, 4, 2>(X, Z, W, K, O, tid);
    if constexpr (variant == 26) wgrad_f2<4, 4, 1>(X, Z, W, K, O, tid);
    if constexpr (variant == 27) wgrad_f2<4, 4, 2>(X, Z, W, K, O, tid);
    if constexpr (variant == 28) wgrad_f2<8, 4, 2>(X, Z, W, K, O, tid);
    if constexpr (variant == 29) wgrad_f2<4, 8, 2>(X, Z, W, K, O, tid);
    if constexpr (variant == 30) dgrad_base<false>(Z, W, X, K, O, tid);
    if constexpr (variant == 31) dgrad_base<true>(Z, W, X, K, O, tid);
    if constexpr (variant == 32) dgrad_k8<false>(Z, W, X, K, O, tid);
    if constexpr (variant == 33) dgrad_k8<true>(Z, W, X, K, O, tid);
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / reps;
  if (tid == 0) sink[blockIdx.x] = X[5] + Z[7] + W[9];
}

static void launch(int id, size_t smem, int K, int O, int reps, long long* out, float* sink) {
  switch (id) {
    case 0: cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<0><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 1: cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<1><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 2: cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 3: cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<3><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 4: cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<4><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 5: cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<5><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 6: cudaFuncSetAttribute(probe<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<6><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 7: cudaFuncSetAttribute(probe<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<7><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 8: cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<8><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 9: cudaFuncSetAttribute(probe<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<9><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 10: cudaFuncSetAttribute(probe<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<10><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 11: cudaFuncSetAttribute(probe<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<11><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 12: cudaFuncSetAttribute(probe<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<12><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 13: cudaFuncSetAttribute(probe<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<13><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 14: cudaFuncSetAttribute(probe<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<14><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 15: cudaFuncSetAttribute(probe<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<15><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 16: cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<16><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 17: cudaFuncSetAttribute(probe<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<17><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 18: cudaFuncSetAttribute(probe<18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<18><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 34: cudaFuncSetAttribute(probe<34>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<34><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 35: cudaFuncSetAttribute(probe<35>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<35><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 36: cudaFuncSetAttribute(probe<36>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<36><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 37: cudaFuncSetAttribute(probe<37>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<37><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 20: cudaFuncSetAttribute(probe<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<20><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 21: cudaFuncSetAttribute(probe<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<21><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 22: cudaFuncSetAttribute(probe<22>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<22><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 23: cudaFuncSetAttribute(probe<23>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<23><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 24: cudaFuncSetAttribute(probe<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<24><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 25: cudaFuncSetAttribute(probe<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<25><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 26: cudaFuncSetAttribute(probe<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<26><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 27: cudaFuncSetAttribute(probe<27>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<27><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 28: cudaFuncSetAttribute(probe<28>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<28><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 29: cudaFuncSetAttribute(probe<29>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<29><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 30: cudaFuncSetAttribute(probe<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<30><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 31: cudaFuncSetAttribute(probe<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<31><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 32: cudaFuncSetAttribute(probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<32><<<148, T, smem>>>(K, O, reps, out, sink); break;
    case 33: cudaFuncSetAttribute(probe<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<33><<<148, T, smem>>>(K, O, reps, out, sink); break;
    default: break;
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4 * 148);
  const size_t smem = (size_t)(200 * RP + 96 * 96 + 96) * 4;
  struct Shape { const char* name; int K, O; };
  const Shape shapes[] = {{"mlp1 K=40 O=80", 40, 80}, {"mlp2 K=80 O=40", 80, 40}, {"mlp3 K=40 O=20", 40, 20}, {"stage1 K=44 O=16", 44, 16}};
  struct Var { int id; const char* name; int kind; };   // kind 0 fwd, 1 wgrad, 2 dgrad
  const Var vars[] = {{0, "fwd 4x4 table (fused.cu today)", 0}, {1, "fwd 4x4 direct rows", 0}, {2, "fwd 4x8 table (fused.cu, O%8==0)", 0},
                      {3, "fwd 4x4 direct FFMA2", 0}, {4, "fwd 4x8 direct FFMA2", 0}, {5, "fwd 4x4 pipelined k4", 0}, {6, "fwd 4x4 pipelined k2", 0},
                      {7, "fwd 4x8 pipelined k2", 0}, {8, "fwd 4x4 pipelined k4 FFMA2", 0}, {9, "fwd 8x4 (8 rows)", 0}, {10, "fwd 8x4 FFMA2", 0},
                      {11, "fwd 8x8", 0}, {12, "fwd 8x8 FFMA2", 0},
                      {13, "fwd 8x8 split-K 2", 0}, {14, "fwd 8x8 split-K 2 FFMA2", 0}, {15, "fwd 8x8 split-K 4", 0}, {16, "fwd 8x8 split-K 4 FFMA2", 0},
                      {17, "fwd 4x16 split-K 2", 0}, {18, "fwd 4x16 split-K 2 FFMA2", 0},
                      {34, "dgrad 8x8 split-O 2", 2}, {35, "dgrad 8x8 split-O 2 FFMA2", 2}, {36, "dgrad 8x8 split-O 4", 2}, {37, "dgrad 8x8 split-O 4 FFMA2", 2},
                      {20, "wgrad 4x4 all rows (today)", 1}, {21, "wgrad 4x4 FFMA2 (pair sums)", 1}, {22, "wgrad 4x4 rows/2 + shfl", 1},
                      {23, "wgrad 8x4 rows/4 + shfl", 1}, {24, "wgrad 8x8 rows/8 + shfl", 1}, {25, "wgrad 8x4 rows/2 + shfl", 1},
                      {26, "wgrad 4x4 FFMA2 packed acc", 1}, {27, "wgrad 4x4 FFMA2 packed, rows/2", 1}, {28, "wgrad 8x4 FFMA2 packed, rows/2", 1}, {29, "wgrad 4x8 FFMA2 packed, rows/2", 1},
                      {30, "dgrad 4x4 (today, static)", 2}, {31, "dgrad 4x4 FFMA2", 2}, {32, "dgrad 8x4", 2}, {33, "dgrad 8x4 FFMA2", 2}};
  for (const Shape& s : shapes) {
    printf("%s: %d rows, ideal FMA cycles at 128/clk = %d\n", s.name, RP, RP * s.K * s.O / 128);
    for (const Var& v : vars) {
      if ((v.id == 2 || v.id == 4 || v.id == 7 || v.id == 11 || v.id == 12 || (v.id >= 13 && v.id <= 16)) && s.O % 8) continue;
      if ((v.id == 17 || v.id == 18) && s.O % 16) continue;
      if ((v.id == 13 || v.id == 14 || v.id == 17 || v.id == 18) && s.K % 2) continue;
      if ((v.id == 15 || v.id == 16) && s.K % 4) continue;
      if (v.id >= 34 && v.id <= 37 && (s.K % 8 || s.O % ((v.id >= 36 ? 4 : 2) * 4))) continue;
      if ((v.id == 23 || v.id == 24 || v.id == 25 || v.id == 32 || v.id == 33) && s.K % 8) continue;
      if (v.id == 24 && s.O % 8) continue;
      if (v.id == 28 && s.K % 8) continue;
      if (v.id == 29 && s.O % 8) continue;
      launch(v.id, smem, s.K, s.O, 20, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      printf("  %-36s %7lld cycles  %5.1f%% of FMA peak%s\n", v.name, c, 100.0 * (RP * s.K * s.O / 128.0) / (double)c, e ? "  CUDA ERROR" : "");
    }
  }
  return 0;
}
