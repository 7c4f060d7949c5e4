// Tensor-core forward of the shared-weight brain (BS.predict, BS_brain.py:147-200, :225-235) on tcgen05 / TMEM.
//
// The combine layers (GNNLayer.call, :44-51) and the decision MLP (:176-200) are genuine dense contractions
// [rows x K] x [K x O]; here they run on the 5th-generation tensor cores:
//   * a tile = TG = floor(128 / N) whole graphs = up to 128 node rows = the 128 TMEM lanes of one accumulator (UMMA M = 128);
//   * every operand lives in shared memory as 4-feature PLANES  plane[k / 4][row][k % 4]  (2 KB per plane): this is the
//     no-swizzle K-major core-matrix layout of the A operand (8 rows x 16 bytes contiguous, SBO = 128 B, LBO = one plane),
//     concatenations ([h | agg | node,edge], :154-175) are plane lists, and the weights W[k][o] are staged once per CTA
//     in the same form with o as the row (plane[k / 4][o][k % 4] = K-major B operand, LBO = one plane).  (A probe
//     of the descriptor conventions, scratch/tc_probe.cu, showed MN-major TF32 operands reading as zeros without
//     swizzle, so both operands are K-major);
//   * fp32 parity (1e-4, north star) rules out single-pass TF32 (10-bit mantissa), so every operand is split
//     x = hi + lo into two TF32 values and every k-step issues THREE tcgen05.mma.kind::tf32 into the same fp32 TMEM
//     accumulator: hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-22 relative);
//   * one elected thread issues the MMAs and a tcgen05.commit on an mbarrier; all 8 warps then read the accumulator
//     with tcgen05.ld (warp w: lanes 32*(w%4).., column half w/4), add the bias, apply ReLU and write the NEXT
//     layer's operand planes (already split into hi/lo) -- activations of the MLP never exist in fp32 anywhere;
//   * the neighbour aggregation between the combine stages is the same register gather-reduce over bit masks as in the
//     FP32 kernel, on the plane layout (zero HBM traffic).
// Used for predict at batch sizes that give every SM at least two tiles; smaller batches stay on the FP32-pipe fused
// kernel, which has finer tiles (see brain.cu).
#include <algorithm>

#include "tc_forward.cuh"

namespace v2v {

namespace {

constexpr int kPlaneBytes = kTcRows * 16;      // one plane: 128 rows x 4 floats
constexpr int kPlaneFloats = kTcRows * 4;
constexpr uint32_t kTmemCols = 128;            // accumulator columns (power of two >= the widest layer, 80)

// ---- tcgen05 primitives (inline PTX, sm_100a) ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {          // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {            // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate, M = 128
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor, no swizzle (layout type 0), sm_100 version field = 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor: fp32 accumulator, TF32 A and B (both K-major), M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
  hi.x = tf32_rn(x.x); hi.y = tf32_rn(x.y); hi.z = tf32_rn(x.z); hi.w = tf32_rn(x.w);
  lo.x = tf32_rn(x.x - hi.x); lo.y = tf32_rn(x.y - hi.y); lo.z = tf32_rn(x.z - hi.z); lo.w = tf32_rn(x.w - hi.w);
}

__global__ void __launch_bounds__(kTcThreads, 1)
tc_forward_kernel(const TcPlan* __restrict__ P, const float* __restrict__ params, const float* __restrict__ node,
                  const float* __restrict__ edge, const uint32_t* __restrict__ in_mask, float* __restrict__ q_out, int B,
                  float* __restrict__ dbg, int dbg_layer) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P->N, TG = P->TG, Dn = P->Dn, De = P->De, CH = P->CH, XP = P->x_planes, dn_pad = P->dn_pad;
  const int FP = P->F >> 2;                                   // planes of h / agg
  const int n_layers = P->n_layers;

  float* Wsm = reinterpret_cast<float*>(smem);
  float* bias_s = Wsm + P->w_floats;
  float* xs = bias_s + P->bias_floats;                        // x0 planes (fp32)
  float* hs = xs + XP * kPlaneFloats;                         // h planes (fp32)
  float* as = hs + FP * kPlaneFloats;                         // aggregated planes (fp32)
  float* stage_hi = as + FP * kPlaneFloats;                   // current MMA operand, hi / lo planes
  float* stage_lo = stage_hi + P->stage_planes * kPlaneFloats;
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(stage_lo + P->stage_planes * kPlaneFloats);

  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, kTmemCols);
  asm volatile("griddepcontrol.wait;" ::: "memory");          // parameters may come from the preceding optimiser kernel
  // ---- weights: parameters -> hi/lo planes over k (K-major B operand), biases zero-padded
  for (int l = 0; l < n_layers; ++l) {
    const TcLayer& L = P->layers[l];
    const int Kpad = L.Kpad, Npad = L.Npad, No = L.N;
    float* Wh = Wsm + L.w_off;
    float* Wl = Wh + Kpad * Npad;
    for (int idx = tid; idx < Kpad * Npad; idx += kTcThreads) {
      const int o = idx % Npad, k = idx / Npad;               // consecutive threads read consecutive o of one W row
      const int row = L.kmap[k];
      const float w = (row >= 0 && o < No) ? params[L.pw_off + row * No + o] : 0.f;
      const float h = tf32_rn(w);
      const int at = (k >> 2) * (Npad * 4) + o * 4 + (k & 3);
      Wh[at] = h;
      Wl[at] = tf32_rn(w - h);
    }
    for (int o = tid; o < Npad; o += kTcThreads) bias_s[L.bias_off + o] = (o < No) ? params[L.pb_off + o] : 0.f;
  }
  for (int i = tid; i < FP * kPlaneFloats; i += kTcThreads) as[i] = 0.f;   // rows past the tile's last node stay zero
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int q4 = warp & 3, half = warp >> 2;
  const int row = q4 * 32 + lane;                             // this thread's accumulator lane = tile row
  const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
  const int num_tiles = (B + TG - 1) / TG;
  const int MS = (N + 1) >> 1;                                // aggregation: two targets per item
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int g0 = tile * TG;
    const int ng = min(TG, B - g0);
    const int R = ng * N;
    // ---- inputs -> x0 planes [node (padded) | edge (padded)], masks
    for (int idx = tid; idx < XP * kTcRows; idx += kTcThreads) {
      const int r = idx & (kTcRows - 1), pl = idx >> 7;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < R) {
        const size_t gr = (size_t)g0 * N + r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = pl * 4 + j;
          if (f < Dn) v[j] = node[gr * Dn + f];
          else if (f >= dn_pad && f - dn_pad < De) v[j] = edge[gr * De + (f - dn_pad)];
        }
      }
      *reinterpret_cast<float4*>(xs + pl * kPlaneFloats + r * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int i = tid; i < TG * N; i += kTcThreads) mask_s[i] = (i < R) ? in_mask[(size_t)g0 * N + i] : 0u;
    __syncthreads();

    for (int l = 0; l < n_layers; ++l) {
      const TcLayer& L = P->layers[l];
      const int Kpad = L.Kpad, Npad = L.Npad;
      // ---- operand planes for the combine layers: split the fp32 planes into hi / lo
      if (L.a_src != 2) {
        const int np = Kpad >> 2;
        for (int idx = tid; idx < np * kTcRows; idx += kTcThreads) {
          const int r = idx & (kTcRows - 1), pl = idx >> 7;
          const float* src;
          if (L.a_src == 0) src = xs + pl * kPlaneFloats;
          else src = pl < FP ? hs + pl * kPlaneFloats : (pl < 2 * FP ? as + (pl - FP) * kPlaneFloats : xs + (pl - 2 * FP) * kPlaneFloats);
          const float4 x = *reinterpret_cast<const float4*>(src + r * 4);
          float4 hi, lo;
          split4(x, hi, lo);
          *reinterpret_cast<float4*>(stage_hi + pl * kPlaneFloats + r * 4) = hi;
          *reinterpret_cast<float4*>(stage_lo + pl * kPlaneFloats + r * 4) = lo;
        }
      }
      fence_async_smem();                 // generic-proxy writes of the operand planes -> visible to the tensor core
      tc_fence_before();                  // the previous epilogue's TMEM reads are ordered before the barrier
      __syncthreads();
      // ---- one thread issues the 3xTF32 MMAs of the layer and commits them to the mbarrier
      if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc_tf32(Npad);
        const uint32_t a_hi = smem_u32(stage_hi), a_lo = smem_u32(stage_lo);
        const uint32_t b_hi = smem_u32(Wsm + L.w_off), b_lo = b_hi + (uint32_t)(Kpad * Npad * 4);
        const uint32_t b_plane = (uint32_t)(Npad * 16);                   // one k-plane of W: Npad output rows x 16 bytes
        const int ksteps = Kpad >> 3;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t ao = (uint32_t)(ks * 2 * kPlaneBytes), bo = (uint32_t)(ks * 2) * b_plane;
          const uint64_t dah = umma_desc(a_hi + ao, kPlaneBytes, 128), dal = umma_desc(a_lo + ao, kPlaneBytes, 128);
          const uint64_t dbh = umma_desc(b_hi + bo, b_plane, 128), dbl = umma_desc(b_lo + bo, b_plane, 128);
          tc_mma_tf32(tmem_base, dal, dbh, idesc, ks > 0 ? 1u : 0u);      // small terms first
          tc_mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          tc_mma_tf32(tmem_base, dah, dbh, idesc, 1u);
        }
        tc_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1u;
      tc_fence_after();
      // ---- epilogue: TMEM -> registers, bias, activation, then the consumer's layout
      const int c_begin = half * (Npad >> 1), c_end = c_begin + (Npad >> 1);
      const float* bs = bias_s + L.bias_off;
      for (int c0 = c_begin; c0 < c_end; c0 += 8) {
        float v[8];
        tc_ld8(t_lane + (uint32_t)c0, v);
        if (dbg && l == dbg_layer && tile == 0) {          // debugging aid: raw accumulator of one layer, [128][Npad]
#pragma unroll
          for (int j = 0; j < 8; ++j) dbg[row * Npad + c0 + j] = v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] += bs[c0 + j];
          if (L.relu) v[j] = fmaxf(v[j], 0.f);
        }
        const float4 x0 = make_float4(v[0], v[1], v[2], v[3]), x1 = make_float4(v[4], v[5], v[6], v[7]);
        const int pl = c0 >> 2;
        if (L.out_kind == 0) {
          *reinterpret_cast<float4*>(hs + pl * kPlaneFloats + row * 4) = x0;
          *reinterpret_cast<float4*>(hs + (pl + 1) * kPlaneFloats + row * 4) = x1;
        } else if (L.out_kind == 1) {
          float4 hi, lo;
          split4(x0, hi, lo);
          *reinterpret_cast<float4*>(stage_hi + pl * kPlaneFloats + row * 4) = hi;
          *reinterpret_cast<float4*>(stage_lo + pl * kPlaneFloats + row * 4) = lo;
          split4(x1, hi, lo);
          *reinterpret_cast<float4*>(stage_hi + (pl + 1) * kPlaneFloats + row * 4) = hi;
          *reinterpret_cast<float4*>(stage_lo + (pl + 1) * kPlaneFloats + row * 4) = lo;
        } else if (row < R) {
          float* qd = q_out + ((size_t)g0 * N + row) * CH;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + j < CH) qd[c0 + j] = v[j];
        }
      }
      // ---- neighbour aggregation after a combine stage: as[pc][g, m] = sum_n Adj[g][n][m] * hs[pc][g, n]
      if (L.out_kind == 0) {
        __syncthreads();
        const int items = FP * TG * MS;
        for (int item = tid; item < items; item += kTcThreads) {
          const int pc = item % FP;
          const int rest = item / FP;
          const int g = rest % TG, ms = rest / TG;
          const int m0 = ms, m1 = ms + MS;
          const uint32_t k0 = mask_s[g * N + m0];
          const uint32_t k1 = (m1 < N) ? mask_s[g * N + m1] : 0u;
          const float* src = hs + pc * kPlaneFloats + (g * N) * 4;
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
          for (int n = 0; n < N; ++n) {
            const float4 x = *reinterpret_cast<const float4*>(src + n * 4);
            if (k0 & (1u << n)) { a0.x += x.x; a0.y += x.y; a0.z += x.z; a0.w += x.w; }
            if (k1 & (1u << n)) { a1.x += x.x; a1.y += x.y; a1.z += x.z; a1.w += x.w; }
          }
          float* dst = as + pc * kPlaneFloats + (g * N) * 4;
          *reinterpret_cast<float4*>(dst + m0 * 4) = a0;
          if (m1 < N) *reinterpret_cast<float4*>(dst + m1 * 4) = a1;
        }
        __syncthreads();                  // the next layer's split pass reads the aggregated planes
      }
    }
    tc_fence_before();
    __syncthreads();                      // the tile's last TMEM reads and shared-memory reads are done
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int tc_build_plan(const TcShape& s, TcPlan* out) {
  TcPlan& P = *out;
  P = TcPlan{};
  V2V_REQUIRE(s.N >= 1 && s.N <= 32, "tensor-core forward: N=%d outside [1,32]", s.N);
  V2V_REQUIRE(s.F % 8 == 0 && s.F >= 8, "tensor-core forward: feedback width %d must be a multiple of 8", s.F);
  V2V_REQUIRE(s.S + 4 <= kTcMaxLayers, "tensor-core forward: too many stages");
  P.N = s.N; P.TG = kTcRows / s.N; P.Dn = s.Dn; P.De = s.De; P.F = s.F; P.CH = s.CH; P.S = s.S;
  P.dn_pad = (s.Dn + 3) & ~3;
  int xp = (P.dn_pad + ((s.De + 3) & ~3)) / 4;
  if (xp & 1) ++xp;                                      // k-steps consume plane pairs
  P.x_planes = xp;
  const int F = s.F, Dn = s.Dn, De = s.De;
  auto r16 = [](int v) { return (v + 15) & ~15; };
  int w_floats = 0, bias_floats = 0, stage_planes = 0;
  int n = 0;
  auto finish = [&](TcLayer& L) {
    V2V_REQUIRE(L.Kpad % 8 == 0 && L.Kpad <= kTcMaxK, "tensor-core forward: contraction length %d unsupported", L.Kpad);
    V2V_REQUIRE(L.Npad >= 16 && L.Npad <= 128, "tensor-core forward: layer width %d unsupported", L.Npad);
    L.w_off = w_floats; w_floats += 2 * L.Kpad * L.Npad;
    L.bias_off = bias_floats; bias_floats += L.Npad;
    stage_planes = std::max(stage_planes, L.Kpad / 4);
    return 0;
  };
  // x0 plane features: node f (f < Dn), zero padding, edge e at dn_pad + e
  auto x0_row = [&](int f, int node_base, int edge_base) -> int {
    if (f < Dn) return node_base >= 0 ? node_base + f : -1;
    if (f >= P.dn_pad && f - P.dn_pad < De) return edge_base >= 0 ? edge_base + (f - P.dn_pad) : -1;
    return -1;
  };
  for (int st = 0; st < s.S; ++st) {
    TcLayer& L = P.layers[n];
    L = TcLayer{};
    L.N = F; L.Npad = r16(F); L.relu = (st < s.S - 1) ? 1 : 0;       // the last combine stage is linear (:161-164)
    L.pw_off = (int)s.w_off[n]; L.pb_off = (int)s.b_off[n];
    L.out_kind = 0;
    for (int k = 0; k < kTcMaxK; ++k) L.kmap[k] = -1;
    if (st == 0) {          // rows of W: node (Dn), edge (De), neighbour (F, all-zero input: skipped)
      L.a_src = 0; L.Kpad = xp * 4;
      for (int f = 0; f < xp * 4; ++f) L.kmap[f] = (short)x0_row(f, 0, Dn);
    } else {                // rows of W: [h | node] (F + Dn), edge (De), aggregated (F)
      L.a_src = 1; L.Kpad = 2 * F + xp * 4;
      for (int f = 0; f < F; ++f) { L.kmap[f] = (short)f; L.kmap[F + f] = (short)(F + Dn + De + f); }
      for (int f = 0; f < xp * 4; ++f) L.kmap[2 * F + f] = (short)x0_row(f, F, F + Dn);
    }
    if (int rc = finish(L)) return rc;
    ++n;
  }
  const int hid[4] = {s.H1, s.H2, s.H3, s.CH};
  int prev_npad = 0;
  for (int j = 0; j < 4; ++j) {
    TcLayer& L = P.layers[n];
    L = TcLayer{};
    L.N = hid[j]; L.Npad = r16(hid[j]); L.relu = j < 3 ? 1 : 0;
    L.pw_off = (int)s.w_off[n]; L.pb_off = (int)s.b_off[n];
    L.out_kind = j < 3 ? 1 : 2;
    for (int k = 0; k < kTcMaxK; ++k) L.kmap[k] = -1;
    if (j == 0) {           // rows of W: node (Dn), h (F), aggregated (F)   ([node | h | agg], :175)
      L.a_src = 1; L.Kpad = 2 * F + xp * 4;
      for (int f = 0; f < F; ++f) { L.kmap[f] = (short)(Dn + f); L.kmap[F + f] = (short)(Dn + F + f); }
      for (int f = 0; f < xp * 4; ++f) L.kmap[2 * F + f] = (short)x0_row(f, 0, -1);
    } else {
      L.a_src = 2; L.Kpad = prev_npad;
      for (int k = 0; k < hid[j - 1]; ++k) L.kmap[k] = (short)k;
    }
    if (int rc = finish(L)) return rc;
    prev_npad = L.Npad;
    ++n;
  }
  P.n_layers = n;
  P.w_floats = w_floats;
  P.bias_floats = (bias_floats + 63) & ~63;
  P.stage_planes = stage_planes;
  const size_t bytes = (size_t)(P.w_floats + P.bias_floats) * 4 + (size_t)(P.x_planes + 2 * (F / 4) + 2 * stage_planes) * kPlaneBytes +
                       (size_t)kTcRows * 4 + 64;
  V2V_REQUIRE(bytes <= 227 * 1024, "tensor-core forward: %zu bytes of shared memory do not fit", bytes);
  P.smem_bytes = (int)bytes;
  return 0;
}

int tc_grid(const TcPlan& p, int B) { return std::max(1, std::min(ceil_div(B, p.TG), sm_count())); }

int tc_forward_launch(const TcPlan& ph, const TcPlan* plan_dev, const float* params, const float* node, const float* edge,
                      const uint32_t* in_mask, float* q_out, int B, cudaStream_t st, float* dbg, int dbg_layer) {
  static int smem_set = 0;
  if (ph.smem_bytes > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ph.smem_bytes));
    smem_set = ph.smem_bytes;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(tc_grid(ph, B));
  lc.blockDim = dim3(kTcThreads);
  lc.dynamicSmemBytes = ph.smem_bytes;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, tc_forward_kernel, plan_dev, params, node, edge, in_mask, q_out, B, dbg, dbg_layer));
  return launch_status("tc_forward_kernel");
}

}  // namespace v2v
