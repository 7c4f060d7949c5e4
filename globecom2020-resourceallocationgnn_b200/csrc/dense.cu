// Combination layers as row-GEMMs  (GNNLayer.call BS_brain.py:44-51, Dense :176-200)
//
//   out[r][:] = act([seg0[r] | seg1[r] | ...] . W[g(r)] + bias[g(r)])
//
// a.W1 + b.W2 + c.W3 of the reference is ONE contraction against the row-stacked
// [W1;W2;W3]; the concatenations the reference materialises (:154-175) become
// "segments" that are gathered straight into the shared-memory operand tile, so no
// concat tensor ever exists in HBM.  Weight sets are either per node slot (G == N,
// what the reference instantiates) or shared (G == 1).
//
// This file is the true-fp32 (FFMA) implementation: required for the 1e-4 parity of
// the fp32 configuration (TF32 tensor-core products alone do not meet it).
#include "v2v_common.cuh"

namespace v2v {

struct RowGemmArgs {
  const float* seg[V2V_MAX_SEG];
  int seg_w[V2V_MAX_SEG];
  int n_seg;
  const float* gate_in;   // optional, same layout as seg[0] (n_seg == 1): x = gate > 0 ? x : 0
  const float* W;         // [G][ldw][w_cols]
  int ldw, w_cols, G;
  int transposed;         // 0: B[k][o] = W[g][k][o]; 1: B[k][o] = W[g][rowsel(o)][k]
  int k0_a, w_a, k0_b, w_b;
  const float* bias;      // [G][NOUT] or null
  int act;
  float* out_a; int out_a_w;
  float* out_b; int out_b_w;
  const float* gate_out;  // optional [rows][out_a_w]
  int K;                  // contraction length
  int L;                  // logical rows per group: B*N (G==1) or B (G==N)
  int N;
};

__device__ __forceinline__ long phys_row(int i, int g, int G, int N) {
  return G == 1 ? (long)i : (long)i * N + g;
}

// NOUT: output columns (multiple of 4). TR: rows per thread. RT: row-threads.
template <int NOUT, int TR, int RT>
__global__ void __launch_bounds__((NOUT / 4) * RT)
rowgemm_kernel(const RowGemmArgs a) {
  constexpr int CG = NOUT / 4;
  constexpr int TM = TR * RT;
  constexpr int NT = CG * RT;
  extern __shared__ __align__(16) float smem[];
  const int K = a.K;
  int KP = (K + 3) & ~3;
  if (((KP >> 2) & 1) == 0) KP += 4;          // KP/4 odd: conflict-free LDS.128 across rows
  float* Xs = smem;                           // [TM][KP]
  float* Ws = smem + TM * KP;                 // [KP][NOUT]

  const int g = blockIdx.y;
  const int row0 = blockIdx.x * TM;
  const int tid = threadIdx.x;
  const int G = a.G, N = a.N;

  // ---- stage weights (B operand) ----
  const float* Wg = a.W + (size_t)g * a.ldw * a.w_cols;
  for (int idx = tid; idx < KP * NOUT; idx += NT) {
    const int k = idx / NOUT, o = idx - k * NOUT;
    float v = 0.f;
    if (k < K) {
      if (!a.transposed) {
        v = Wg[(size_t)k * a.w_cols + o];
      } else {
        const int r = (o < a.w_a) ? (a.k0_a + o) : (a.k0_b + (o - a.w_a));
        v = Wg[(size_t)r * a.w_cols + k];
      }
    }
    Ws[idx] = v;
  }
  // ---- stage the operand tile, gathering the segments ----
  {
    int off = 0;
    for (int s = 0; s < a.n_seg; ++s) {
      const int w = a.seg_w[s];
      const float* src = a.seg[s];
      for (int idx = tid; idx < TM * w; idx += NT) {
        const int i = idx / w, cidx = idx - i * w;
        const int li = row0 + i;
        float v = 0.f;
        if (li < a.L) {
          const long pr = phys_row(li, g, G, N);
          v = src[pr * w + cidx];
          if (a.gate_in) v = (a.gate_in[pr * w + cidx] > 0.f) ? v : 0.f;
        }
        Xs[i * KP + off + cidx] = v;
      }
      off += w;
    }
    const int padw = KP - K;
    for (int idx = tid; idx < TM * padw; idx += NT) {
      const int i = idx / padw, cidx = idx - i * padw;
      Xs[i * KP + K + cidx] = 0.f;
    }
  }
  __syncthreads();

  const int cg = tid % CG, rt = tid / CG;
  float4 acc[TR];
  {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) b4 = *reinterpret_cast<const float4*>(a.bias + (size_t)g * NOUT + cg * 4);
#pragma unroll
    for (int i = 0; i < TR; ++i) acc[i] = b4;
  }
  const float4* Ws4 = reinterpret_cast<const float4*>(Ws) + cg;
  for (int k4 = 0; k4 < KP; k4 += 4) {
    float4 x[TR];
#pragma unroll
    for (int i = 0; i < TR; ++i) x[i] = *reinterpret_cast<const float4*>(Xs + (rt + i * RT) * KP + k4);
    const float4 w0 = Ws4[(k4 + 0) * CG], w1 = Ws4[(k4 + 1) * CG], w2 = Ws4[(k4 + 2) * CG], w3 = Ws4[(k4 + 3) * CG];
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      acc[i].x = fmaf(x[i].x, w0.x, acc[i].x); acc[i].y = fmaf(x[i].x, w0.y, acc[i].y);
      acc[i].z = fmaf(x[i].x, w0.z, acc[i].z); acc[i].w = fmaf(x[i].x, w0.w, acc[i].w);
      acc[i].x = fmaf(x[i].y, w1.x, acc[i].x); acc[i].y = fmaf(x[i].y, w1.y, acc[i].y);
      acc[i].z = fmaf(x[i].y, w1.z, acc[i].z); acc[i].w = fmaf(x[i].y, w1.w, acc[i].w);
      acc[i].x = fmaf(x[i].z, w2.x, acc[i].x); acc[i].y = fmaf(x[i].z, w2.y, acc[i].y);
      acc[i].z = fmaf(x[i].z, w2.z, acc[i].z); acc[i].w = fmaf(x[i].z, w2.w, acc[i].w);
      acc[i].x = fmaf(x[i].w, w3.x, acc[i].x); acc[i].y = fmaf(x[i].w, w3.y, acc[i].y);
      acc[i].z = fmaf(x[i].w, w3.z, acc[i].z); acc[i].w = fmaf(x[i].w, w3.w, acc[i].w);
    }
  }
  // ---- epilogue ----
  const int col = cg * 4;
#pragma unroll
  for (int i = 0; i < TR; ++i) {
    const int li = row0 + rt + i * RT;
    if (li >= a.L) continue;
    const long pr = phys_row(li, g, G, N);
    float4 v = acc[i];
    if (a.act) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (col < a.out_a_w) {
      if (a.gate_out) {
        const float4 gt = *reinterpret_cast<const float4*>(a.gate_out + pr * a.out_a_w + col);
        v.x = gt.x > 0.f ? v.x : 0.f; v.y = gt.y > 0.f ? v.y : 0.f;
        v.z = gt.z > 0.f ? v.z : 0.f; v.w = gt.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(a.out_a + pr * a.out_a_w + col) = v;
    } else {
      *reinterpret_cast<float4*>(a.out_b + pr * a.out_b_w + (col - a.out_a_w)) = v;
    }
  }
}

// Fallback for output widths without a tuned instantiation: one thread per element.
__global__ void rowgemm_generic_kernel(const RowGemmArgs a, int NOUT) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (idx >= (long)a.L * NOUT) return;
  const int li = (int)(idx / NOUT), o = (int)(idx - (long)li * NOUT);
  const long pr = phys_row(li, g, a.G, a.N);
  const float* Wg = a.W + (size_t)g * a.ldw * a.w_cols;
  float acc = a.bias ? a.bias[(size_t)g * NOUT + o] : 0.f;
  int k = 0;
  for (int s = 0; s < a.n_seg; ++s) {
    const int w = a.seg_w[s];
    for (int cidx = 0; cidx < w; ++cidx, ++k) {
      float x = a.seg[s][pr * w + cidx];
      if (a.gate_in) x = (a.gate_in[pr * w + cidx] > 0.f) ? x : 0.f;
      float wv;
      if (!a.transposed) {
        wv = Wg[(size_t)k * a.w_cols + o];
      } else {
        const int r = (o < a.w_a) ? (a.k0_a + o) : (a.k0_b + (o - a.w_a));
        wv = Wg[(size_t)r * a.w_cols + k];
      }
      acc = fmaf(x, wv, acc);
    }
  }
  if (a.act) acc = fmaxf(acc, 0.f);
  if (o < a.out_a_w) {
    if (a.gate_out) acc = a.gate_out[pr * a.out_a_w + o] > 0.f ? acc : 0.f;
    a.out_a[pr * a.out_a_w + o] = acc;
  } else {
    a.out_b[pr * a.out_b_w + (o - a.out_a_w)] = acc;
  }
}

template <int NOUT, int TR, int RT>
static int launch_rowgemm_t(const RowGemmArgs& a, cudaStream_t st) {
  constexpr int TM = TR * RT;
  int KP = (a.K + 3) & ~3;
  if (((KP >> 2) & 1) == 0) KP += 4;
  size_t smem = (size_t)(TM * KP + KP * NOUT) * sizeof(float);
  auto k = rowgemm_kernel<NOUT, TR, RT>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid(ceil_div(a.L, TM), a.G);
  k<<<grid, (NOUT / 4) * RT, smem, st>>>(a);
  return launch_status("rowgemm_kernel");
}

static int launch_rowgemm(const RowGemmArgs& a, int NOUT, cudaStream_t st) {
  if (a.L <= 0) return 0;
  const bool vec_ok = (a.out_a_w % 4 == 0) && (a.out_b_w % 4 == 0) &&
                      ((((uintptr_t)a.out_a) | ((uintptr_t)a.out_b) | ((uintptr_t)a.gate_out) | ((uintptr_t)a.bias)) & 15u) == 0 &&
                      a.K <= 256;
  if (vec_ok) {
    switch (NOUT) {
      case 4: return launch_rowgemm_t<4, 1, 128>(a, st);
      case 8: return launch_rowgemm_t<8, 2, 64>(a, st);
      case 16: return launch_rowgemm_t<16, 4, 32>(a, st);
      case 20: return launch_rowgemm_t<20, 4, 32>(a, st);
      case 32: return launch_rowgemm_t<32, 8, 16>(a, st);
      case 40: return launch_rowgemm_t<40, 8, 16>(a, st);
      case 64: return launch_rowgemm_t<64, 8, 8>(a, st);
      case 80: return launch_rowgemm_t<80, 8, 8>(a, st);
      default: break;
    }
  }
  int threads = 256;
  dim3 grid((unsigned)(((long)a.L * NOUT + threads - 1) / threads), a.G);
  rowgemm_generic_kernel<<<grid, threads, 0, st>>>(a, NOUT);
  return launch_status("rowgemm_generic_kernel");
}

// ---------------------------------------------------------------------------
// weight / bias gradient
// ---------------------------------------------------------------------------
struct WGradArgs {
  const float* seg[V2V_MAX_SEG];
  int seg_w[V2V_MAX_SEG];
  int n_seg;
  const float* dY;        // [rows][NOUT]
  const float* gate_in;   // optional [rows][NOUT]: dZ = gate > 0 ? dY : 0
  float* dW; int ldw;     // [G][ldw][NOUT], accumulated
  float* db;              // [G][NOUT], accumulated
  int K, NOUT, L, N, G;
};

constexpr int kWgTM = 128;
constexpr int kWgThreads = 256;

__global__ void __launch_bounds__(kWgThreads)
wgrad_kernel(const WGradArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int K = a.K, NOUT = a.NOUT;
  const int KP = (K + 3) & ~3;
  float* Xs = smem;                       // [TM][KP]
  float* Zs = Xs + kWgTM * KP;            // [TM][NOUT]
  float* Rs = Zs + kWgTM * NOUT;          // [KP+1][NOUT] block-level reduction (last row: bias)
  const int g = blockIdx.y, tid = threadIdx.x;
  const int row0 = blockIdx.x * kWgTM;
  const int G = a.G, N = a.N;

  for (int idx = tid; idx < (KP + 1) * NOUT; idx += kWgThreads) Rs[idx] = 0.f;
  {
    int off = 0;
    for (int s = 0; s < a.n_seg; ++s) {
      const int w = a.seg_w[s];
      const float* src = a.seg[s];
      for (int idx = tid; idx < kWgTM * w; idx += kWgThreads) {
        const int i = idx / w, cidx = idx - i * w;
        const int li = row0 + i;
        Xs[i * KP + off + cidx] = (li < a.L) ? src[phys_row(li, g, G, N) * w + cidx] : 0.f;
      }
      off += w;
    }
    const int padw = KP - K;
    for (int idx = tid; idx < kWgTM * padw; idx += kWgThreads) {
      const int i = idx / padw, cidx = idx - i * padw;
      Xs[i * KP + K + cidx] = 0.f;
    }
    for (int idx = tid; idx < kWgTM * NOUT; idx += kWgThreads) {
      const int i = idx / NOUT, o = idx - i * NOUT;
      const int li = row0 + i;
      float v = 0.f;
      if (li < a.L) {
        const long pr = phys_row(li, g, G, N);
        v = a.dY[pr * NOUT + o];
        if (a.gate_in) v = a.gate_in[pr * NOUT + o] > 0.f ? v : 0.f;
      }
      Zs[idx] = v;
    }
  }
  __syncthreads();

  const int KG = KP >> 2, OG = NOUT >> 2;
  const int tiles = KG * OG;
  const int RG = max(1, kWgThreads / tiles);        // row groups working on the same 4x4 tile
  const int rg = tid / tiles;
  if (rg < RG || tiles > kWgThreads) {
    for (int t = (tiles > kWgThreads) ? tid : (tid - rg * tiles); t < tiles; t += kWgThreads) {
      const int kg = t / OG, og = t - kg * OG;
      float acc[4][4];
      float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      const int r_begin = (tiles > kWgThreads) ? 0 : rg;
      const int r_step = (tiles > kWgThreads) ? 1 : RG;
#pragma unroll 4
      for (int r = r_begin; r < kWgTM; r += r_step) {
        const float4 x = *reinterpret_cast<const float4*>(Xs + r * KP + kg * 4);
        const float4 z = *reinterpret_cast<const float4*>(Zs + r * NOUT + og * 4);
        acc[0][0] = fmaf(x.x, z.x, acc[0][0]); acc[0][1] = fmaf(x.x, z.y, acc[0][1]);
        acc[0][2] = fmaf(x.x, z.z, acc[0][2]); acc[0][3] = fmaf(x.x, z.w, acc[0][3]);
        acc[1][0] = fmaf(x.y, z.x, acc[1][0]); acc[1][1] = fmaf(x.y, z.y, acc[1][1]);
        acc[1][2] = fmaf(x.y, z.z, acc[1][2]); acc[1][3] = fmaf(x.y, z.w, acc[1][3]);
        acc[2][0] = fmaf(x.z, z.x, acc[2][0]); acc[2][1] = fmaf(x.z, z.y, acc[2][1]);
        acc[2][2] = fmaf(x.z, z.z, acc[2][2]); acc[2][3] = fmaf(x.z, z.w, acc[2][3]);
        acc[3][0] = fmaf(x.w, z.x, acc[3][0]); acc[3][1] = fmaf(x.w, z.y, acc[3][1]);
        acc[3][2] = fmaf(x.w, z.z, acc[3][2]); acc[3][3] = fmaf(x.w, z.w, acc[3][3]);
        if (kg == 0) { bs.x += z.x; bs.y += z.y; bs.z += z.z; bs.w += z.w; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&Rs[(kg * 4 + i) * NOUT + og * 4 + j], acc[i][j]);
      if (kg == 0) {
        atomicAdd(&Rs[KP * NOUT + og * 4 + 0], bs.x); atomicAdd(&Rs[KP * NOUT + og * 4 + 1], bs.y);
        atomicAdd(&Rs[KP * NOUT + og * 4 + 2], bs.z); atomicAdd(&Rs[KP * NOUT + og * 4 + 3], bs.w);
      }
    }
  }
  __syncthreads();
  float* dWg = a.dW + (size_t)g * a.ldw * NOUT;
  for (int idx = tid; idx < K * NOUT; idx += kWgThreads) atomicAdd(&dWg[idx], Rs[idx]);
  if (a.db)
    for (int o = tid; o < NOUT; o += kWgThreads) atomicAdd(&a.db[(size_t)g * NOUT + o], Rs[KP * NOUT + o]);
}

// scalar fallback (NOUT not a multiple of 4, or tile too large for shared memory)
__global__ void wgrad_generic_kernel(const WGradArgs a) {
  const int g = blockIdx.y;
  const int ko = blockIdx.x * blockDim.x + threadIdx.x;      // k*NOUT + o, k == K -> bias
  if (ko >= (a.K + 1) * a.NOUT) return;
  const int k = ko / a.NOUT, o = ko - k * a.NOUT;
  int s = 0, kk = k;
  if (k < a.K)
    while (kk >= a.seg_w[s]) { kk -= a.seg_w[s]; ++s; }
  float acc = 0.f;
  for (int li = 0; li < a.L; ++li) {
    const long pr = phys_row(li, g, a.G, a.N);
    float z = a.dY[pr * a.NOUT + o];
    if (a.gate_in) z = a.gate_in[pr * a.NOUT + o] > 0.f ? z : 0.f;
    const float x = (k < a.K) ? a.seg[s][pr * a.seg_w[s] + kk] : 1.f;
    acc = fmaf(x, z, acc);
  }
  if (k < a.K) atomicAdd(&a.dW[((size_t)g * a.ldw + k) * a.NOUT + o], acc);
  else if (a.db) atomicAdd(&a.db[(size_t)g * a.NOUT + o], acc);
}

static int launch_wgrad(const WGradArgs& a, cudaStream_t st) {
  if (a.L <= 0) return 0;
  const int KP = (a.K + 3) & ~3;
  size_t smem = (size_t)(kWgTM * KP + kWgTM * a.NOUT + (KP + 1) * a.NOUT) * sizeof(float);
  if (a.NOUT % 4 == 0 && smem <= 200 * 1024) {
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      V2V_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    dim3 grid(ceil_div(a.L, kWgTM), a.G);
    wgrad_kernel<<<grid, kWgThreads, smem, st>>>(a);
    return launch_status("wgrad_kernel");
  }
  dim3 grid(ceil_div((a.K + 1) * a.NOUT, 128), a.G);
  wgrad_generic_kernel<<<grid, 128, 0, st>>>(a);
  return launch_status("wgrad_generic_kernel");
}

}  // namespace v2v

using namespace v2v;

static int check_group(int B, int N, int G, const char* who) {
  V2V_REQUIRE(B >= 0 && N > 0, "%s: bad shape B=%d N=%d", who, B, N);
  V2V_REQUIRE(G == 1 || G == N, "%s: G must be 1 (shared) or N (per slot), got G=%d N=%d", who, G, N);
  return 0;
}

extern "C" int v2v_dense_fwd(int n_seg, const float* const* seg_dev, const int* seg_width,
                             const float* W_dev, int ldw, const float* bias_dev, float* out_dev,
                             int B, int N, int G, int n_out, int act, void* stream) {
  if (int rc = check_group(B, N, G, "v2v_dense_fwd")) return rc;
  V2V_REQUIRE(n_seg >= 1 && n_seg <= V2V_MAX_SEG, "v2v_dense_fwd: n_seg=%d out of range", n_seg);
  V2V_REQUIRE(n_out > 0 && W_dev && out_dev, "v2v_dense_fwd: bad arguments");
  RowGemmArgs a{};
  a.n_seg = n_seg;
  int K = 0;
  for (int s = 0; s < n_seg; ++s) {
    V2V_REQUIRE(seg_dev[s] && seg_width[s] > 0, "v2v_dense_fwd: segment %d invalid", s);
    a.seg[s] = seg_dev[s]; a.seg_w[s] = seg_width[s]; K += seg_width[s];
  }
  V2V_REQUIRE(K <= ldw, "v2v_dense_fwd: sum of segment widths %d exceeds weight rows %d", K, ldw);
  a.W = W_dev; a.ldw = ldw; a.w_cols = n_out; a.G = G; a.transposed = 0;
  a.bias = bias_dev; a.act = act;
  a.out_a = out_dev; a.out_a_w = n_out; a.out_b = nullptr; a.out_b_w = 0;
  a.K = K; a.L = (G == 1) ? B * N : B; a.N = N;
  return launch_rowgemm(a, n_out, (cudaStream_t)stream);
}

extern "C" int v2v_dense_bwd_data(const float* dY_dev, const float* gate_in_dev, const float* W_dev,
                                  int ldw, int k0_a, int w_a, float* dxa_dev, int k0_b, int w_b,
                                  float* dxb_dev, const float* gate_out_dev, int B, int N, int G,
                                  int n_out, void* stream) {
  if (int rc = check_group(B, N, G, "v2v_dense_bwd_data")) return rc;
  V2V_REQUIRE(dY_dev && W_dev && dxa_dev && w_a > 0, "v2v_dense_bwd_data: bad arguments");
  V2V_REQUIRE(w_b == 0 || dxb_dev, "v2v_dense_bwd_data: second range needs an output");
  V2V_REQUIRE(!(gate_out_dev && w_b > 0), "v2v_dense_bwd_data: gate_out only with one range");
  V2V_REQUIRE(k0_a >= 0 && k0_a + w_a <= ldw && k0_b >= 0 && k0_b + w_b <= ldw, "v2v_dense_bwd_data: column range outside the weight");
  RowGemmArgs a{};
  a.n_seg = 1; a.seg[0] = dY_dev; a.seg_w[0] = n_out; a.gate_in = gate_in_dev;
  a.W = W_dev; a.ldw = ldw; a.w_cols = n_out; a.G = G; a.transposed = 1;
  a.k0_a = k0_a; a.w_a = w_a; a.k0_b = k0_b; a.w_b = w_b;
  a.bias = nullptr; a.act = 0;
  a.out_a = dxa_dev; a.out_a_w = w_a; a.out_b = dxb_dev; a.out_b_w = w_b;
  a.gate_out = gate_out_dev;
  a.K = n_out; a.L = (G == 1) ? B * N : B; a.N = N;
  return launch_rowgemm(a, w_a + w_b, (cudaStream_t)stream);
}

extern "C" int v2v_dense_bwd_weight(int n_seg, const float* const* seg_dev, const int* seg_width,
                                    const float* dY_dev, const float* gate_in_dev, float* dW_dev,
                                    int ldw, float* db_dev, int B, int N, int G, int n_out,
                                    void* stream) {
  if (int rc = check_group(B, N, G, "v2v_dense_bwd_weight")) return rc;
  V2V_REQUIRE(n_seg >= 1 && n_seg <= V2V_MAX_SEG, "v2v_dense_bwd_weight: n_seg=%d out of range", n_seg);
  V2V_REQUIRE(dY_dev && dW_dev && n_out > 0, "v2v_dense_bwd_weight: bad arguments");
  WGradArgs a{};
  a.n_seg = n_seg;
  int K = 0;
  for (int s = 0; s < n_seg; ++s) {
    V2V_REQUIRE(seg_dev[s] && seg_width[s] > 0, "v2v_dense_bwd_weight: segment %d invalid", s);
    a.seg[s] = seg_dev[s]; a.seg_w[s] = seg_width[s]; K += seg_width[s];
  }
  V2V_REQUIRE(K <= ldw, "v2v_dense_bwd_weight: sum of segment widths %d exceeds weight rows %d", K, ldw);
  a.dY = dY_dev; a.gate_in = gate_in_dev; a.dW = dW_dev; a.ldw = ldw; a.db = db_dev;
  a.K = K; a.NOUT = n_out; a.L = (G == 1) ? B * N : B; a.N = N; a.G = G;
  return launch_wgrad(a, (cudaStream_t)stream);
}
