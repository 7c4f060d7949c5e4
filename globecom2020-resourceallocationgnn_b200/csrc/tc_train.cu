// bf16 tensor-core training step of the shared-weight brain (BASELINE configs[2]: "3-layer GNN bf16"):
//   forward (BS_brain.py:147-200) + Huber head (:86-87) + full backward, every contraction on tcgen05 / TMEM.
//
// One plane layout serves all three contraction families.  A tile = TG = floor(128 / N) whole graphs = up to 128 node
// rows = the 128 TMEM lanes of an accumulator (UMMA M = 128).  Every activation and every back-propagated gradient of
// the tile lives in shared memory as bf16 PLANES of 8 features:  plane[f / 8][row][f % 8]  (2 KB per plane, 16 bytes per
// row), the weights W[k][o] of a layer as  wplane[k / 8][o][k % 8].  The same bytes are read
//   * K-major  (SBO = 128 B, LBO = one plane)  as the A operand of the forward / data-gradient contractions
//     (M = rows, K = features) and, for the weights, as the B operand of the forward (N = o, K = k);
//   * MN-major (LBO = 128 B, SBO = one plane)  as BOTH operands of the weight gradient  dW = X^T dZ  (M = input feature,
//     N = output feature, K = the 128 rows) and, for the weights, as the B operand of the data gradient
//     dX = dZ W^T  (N = k, K = o) -- no transposed copy of anything exists (scratch/tc_probe_mn.cu: no-swizzle
//     MN-major operands work for kind::f16; they read as zeros for kind::tf32, which is why the fp32 configuration
//     keeps the FP32-pipe kernel of fused.cu and this kernel is the bf16 configuration's).
// The reference's concatenations ([h | node | edge | agg], :154-164; [node | h | agg], :175) are plane lists.
//
// Weight gradients never leave tensor memory during the kernel: three accumulator groups
//   chain 1: [x0 | h0 | a0 | h1 | a1 | h2 | a2 | ..]^T  x  [dz0 | dz1 | dz2 | dm1 | dm2 | dm3 | dq]   (128 x 208 at S = 3)
//   chain 2: [m1 | m2 | ..]^T x [dm2 | dm3],    chain 3: [m3 | ..]^T x [dq | 0]
// are accumulated over ALL tiles of the CTA (24 MMAs per tile) and read out once, at the end, into the CTA's partial
// row; the wanted blocks (layer l: inputs of l x dz of l) are picked by a lane / column table, the rest of the cross
// product is ignored.  x0 carries a ones feature, so the bias gradients are row 15 of chain 1.  The per-CTA partials
// (+ per-head Huber sums in the row tail) go through the same reduce + Keras-Adam kernel as the FP32 path (fused.cu).
//
// Rounding points (restated in oracle/bf16_emul.py): contraction operands are bf16 (RNE), accumulation is fp32 in
// TMEM; bias, ReLU, the neighbour aggregation, the output layer and the Huber head are fp32; master weights, gradients
// and Adam are fp32.  Warp roles: warp 16 issues MMAs from a per-tile table, 16 epilogue warps (4 TMEM lane quarters x
// 4 column quarters) run the epilogues; one tile in flight, inputs of the next tile wait in registers.
#include <algorithm>
#include <vector>

#include "tc_train.cuh"

namespace v2v {

namespace {

constexpr uint32_t kTtTmemCols = 512;

// ---- tcgen05 primitives (inline PTX, sm_100a) ------------------------------------------------------------------
__device__ __forceinline__ void tt_tmem_alloc(uint32_t* smem_dst, uint32_t cols) {          // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tt_tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, M = 128
__device__ __forceinline__ void tt_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tt_epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTtEpiThreads) : "memory"); }
__device__ __forceinline__ void tt_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tt_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor, no swizzle (layout type 0), sm_100 version field = 1
__device__ __forceinline__ uint64_t tt_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
// relu gate from a saved bf16 activation: positive and non-zero
__device__ __forceinline__ bool bf16_pos(uint32_t word, int half) {
  const int16_t h = (int16_t)(half ? (word >> 16) : (word & 0xffffu));
  return h > 0;
}

// fp32 master weights -> bf16 weight image (k-planes of 8, see above) + fp32 bias image, once per call
__global__ void __launch_bounds__(256)
tt_stage_weights_kernel(const TtPlan* __restrict__ P, const float* __restrict__ params, __nv_bfloat16* __restrict__ wimg) {
  asm volatile("griddepcontrol.wait;" ::: "memory");          // parameters may come from the preceding optimiser kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int n_layers = P->n_layers;
  float* bias_img = reinterpret_cast<float*>(wimg + P->w_elems);
  for (int l = 0; l < n_layers; ++l) {
    const TtLayerImg& L = P->layers[l];
    const int Kpad = L.Kpad, Npad = L.Npad, No = L.No;
    __nv_bfloat16* Wc = wimg + L.w_off;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < Kpad * Npad; idx += gridDim.x * blockDim.x) {
      const int o = idx % Npad, k = idx / Npad;               // consecutive threads read consecutive o of one W row
      const int wrow = L.kmap[k];
      const float w = (wrow >= 0 && o < No) ? params[L.pw_off + wrow * No + o] : 0.f;
      Wc[(k >> 3) * (Npad * 8) + o * 8 + (k & 7)] = __float2bfloat16_rn(w);
    }
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < Npad; o += gridDim.x * blockDim.x)
      bias_img[L.bias_off + o] = (o < No) ? params[L.pb_off + o] : 0.f;
  }
}

__global__ void __launch_bounds__(kTtThreads, 1)
tt_kernel(const TtPlan* __restrict__ P, const uint8_t* __restrict__ wimg, const float* __restrict__ node,
          const float* __restrict__ edge, const uint32_t* __restrict__ in_mask, const uint32_t* __restrict__ out_mask,
          const float* __restrict__ y, float* __restrict__ q_out, float* __restrict__ partial, int B, int train, float inv_cnt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t ops_bar;          // operands of the next step are in place (16 warp arrivals)
  __shared__ __align__(8) uint64_t acc_bar;          // the step's accumulator is complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t wg_bar;           // the tile's weight-gradient MMAs have consumed the planes
  __shared__ __align__(8) uint64_t w_bar;            // the weight image has landed
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P->N, TG = P->TG, Dn = P->Dn, De = P->De, CH = P->CH, F = P->F, XP = P->XP;
  const int n_steps = train ? P->n_steps_train : P->n_steps_fwd;
  const int ones_f = P->ones_feature;
  float* bias_s = reinterpret_cast<float*>(smem + P->off_bias);
  uint8_t* planes = smem + P->off_planes;
  float* scr = reinterpret_cast<float*>(smem + P->off_scr);          // fp32 scratch planes [c / 4][row][4]
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(smem + P->off_mask); // in_mask[128] | out_mask[128]
  float* hl_s = reinterpret_cast<float*>(smem + P->off_misc);        // per-head Huber sums [32]
  float* rowloss = hl_s + 32;                                        // [128]
  TtMma* mma_s = reinterpret_cast<TtMma*>(smem + P->off_tab);
  TtStep* step_s = reinterpret_cast<TtStep*>(mma_s + kTtMaxMma);

  if (tid == 0) {
    mbar_init(&ops_bar, kTtEpiThreads / 32);
    mbar_init(&acc_bar, 1);
    mbar_init(&wg_bar, 1);
    mbar_init(&w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tt_tmem_alloc(&tmem_base_s, kTtTmemCols);
  {                                                           // planes, scratch, masks, loss sums: finite everywhere
    uint32_t* z = reinterpret_cast<uint32_t*>(planes);
    const int nz = (P->off_tab - P->off_planes) / 4;
    for (int i = tid; i < nz; i += kTtThreads) z[i] = 0u;
    const int n_mma_w = P->n_mma * (int)(sizeof(TtMma) / 4), n_step_w = kTtMaxSteps * (int)(sizeof(TtStep) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P->mma);
    uint32_t* dst = reinterpret_cast<uint32_t*>(mma_s);
    for (int i = tid; i < n_mma_w; i += kTtThreads) dst[i] = src[i];
    src = reinterpret_cast<const uint32_t*>(P->steps);
    dst = reinterpret_cast<uint32_t*>(step_s);
    for (int i = tid; i < n_step_w; i += kTtThreads) dst[i] = src[i];
  }
  __syncthreads();                                            // mbarrier inits and tables visible
  asm volatile("griddepcontrol.wait;" ::: "memory");          // the weight image comes from tt_stage_weights_kernel
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(P->w_elems * 2 + P->bias_floats * 4);
    mbar_arrive_expect_tx(&w_bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u)
      bulk_g2s(smem + P->off_w + off, wimg + off, min(32768u, bytes - off), &w_bar);
  }
  mbar_wait(&w_bar, 0u);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  fence_async_smem();
  tt_fence_before();
  __syncthreads();
  tt_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int num_tiles = (B + TG - 1) / TG;

  if (warp == kTtEpiThreads / 32) {
    // =============================== MMA warp: one thread issues everything ===============================
    if (lane == 0) {
      uint32_t po = 0u;
      bool first = true;
      const uint32_t sbase = smem_u32(smem);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int s = 0; s < n_steps; ++s) {
          mbar_wait(&ops_bar, po);
          po ^= 1u;
          tt_fence_after();
          const int m0 = step_s[s].mma0, m1 = m0 + step_s[s].n_mma;
          for (int i = m0; i < m1; ++i) {
            const TtMma m = mma_s[i];
            const uint64_t da = tt_desc(sbase + m.a_off, m.a_lbo, m.a_sbo), db = tt_desc(sbase + m.b_off, m.b_lbo, m.b_sbo);
            const uint32_t acc = (m.acc == 0u) ? 0u : ((m.acc == 1u) ? 1u : (first ? 0u : 1u));
            tt_mma_bf16(tmem_base + m.dcol, da, db, m.idesc, acc);
          }
          tt_commit(step_s[s].kind == TT_WGRAD ? &wg_bar : &acc_bar);
        }
        first = false;
      }
    }
  } else {
    // =============================== 16 epilogue warps ===============================
    const int q4 = warp & 3, cq = warp >> 2;                  // TMEM lane quarter, column quarter
    const int row = q4 * 32 + lane;                           // this thread's accumulator lane = tile row
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    uint32_t pa = 0u, pw = 0u;
    // ---- input prefetch: the next tile's features / masks / targets wait in registers
    float xin[8], ynext[8];
    uint32_t mk = 0u;
    auto fetch = [&](int tile) {
      const int g0 = tile * TG;
      const int R = (tile < num_tiles) ? min(TG, B - g0) * N : 0;
      const int r = tid & (kTtRows - 1), pl = tid >> 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) { xin[j] = 0.f; ynext[j] = 0.f; }
      if (pl < XP && r < R) {
        const size_t gr = (size_t)g0 * N + r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int f = pl * 8 + j;
          if (f < Dn) xin[j] = __ldg(node + gr * Dn + f);
          else if (f < Dn + De) xin[j] = __ldg(edge + gr * De + (f - Dn));
          else if (f == ones_f) xin[j] = 1.f;
        }
      }
      mk = 0u;
      if (tid < kTtRows) { if (tid < R) mk = __ldg(in_mask + (size_t)g0 * N + tid); }
      else if (tid < 2 * kTtRows && train) { if (tid - kTtRows < R) mk = __ldg(out_mask + (size_t)g0 * N + (tid - kTtRows)); }
      if (train && cq == 0 && row < R) {
        const float* yp = y + ((size_t)g0 * N + row) * CH;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < CH) ynext[j] = __ldg(yp + j);
      }
    };
    fetch(blockIdx.x);
    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int g0 = tile * TG;
      const int R = min(TG, B - g0) * N;
      if (train && !first) {                                  // the previous tile's weight-gradient MMAs still read the planes
        mbar_wait(&wg_bar, pw);
        pw ^= 1u;
        tt_fence_after();
      }
      first = false;
      {
        const int r = tid & (kTtRows - 1), pl = tid >> 7;
        if (pl < XP) *reinterpret_cast<uint4*>(planes + (size_t)pl * kTtPlaneBytes + r * 16) = pack8(xin);
        if (tid < 2 * kTtRows) mask_s[tid] = mk;
      }
      float ycur[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ycur[j] = ynext[j];
      fence_async_smem();
      tt_fence_before();
      __syncwarp();
      if (lane == 0) tt_mbar_arrive(&ops_bar);
      fetch(tile + gridDim.x);

      for (int s = 0; s < n_steps; ++s) {
        const TtStep st = step_s[s];
        if (st.kind == TT_WGRAD) break;                       // no epilogue: its completion is awaited before the planes are reused
        mbar_wait(&acc_bar, pa);
        pa ^= 1u;
        tt_fence_after();
        const float* bs = bias_s + st.bias_off;
        // ---- accumulator chunks of this thread: 8 columns each, chunk c8 = cq + 4 j
        uint32_t acc[4][8];
        const int nch_ld = (st.kind == TT_DGRAD_AGG) ? (st.npad >> 3) : st.n_chunks;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c8 = cq + 4 * j;
          if (c8 < nch_ld) tt_ld8_issue(t_lane + (uint32_t)(c8 * 8), acc[j]);
        }
        tt_wait_ld();
        if (st.kind == TT_COMBINE || st.kind == TT_MLP) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c8 = cq + 4 * j;
            if (c8 < nch_ld) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[i] = __uint_as_float(acc[j][i]) + bs[c8 * 8 + i];
                if (st.relu) v[i] = fmaxf(v[i], 0.f);
              }
              if (st.kind == TT_COMBINE) {                    // fp32 copy for the aggregation
                *reinterpret_cast<float4*>(scr + ((size_t)(2 * c8) * kTtRows + row) * 4) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(scr + ((size_t)(2 * c8 + 1) * kTtRows + row) * 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
              *reinterpret_cast<uint4*>(planes + (size_t)(st.out_plane + c8) * kTtPlaneBytes + row * 16) = pack8(v);
            }
          }
          if (st.kind == TT_COMBINE) {
            // neighbour aggregation (AggLayer.call, :69-76): agg[g, m] = sum_n Adj[g][n][m] h[g, n], fp32, then bf16 planes
            tt_epi_barrier();
            const int items = (F >> 2) * TG * N;
            for (int item = tid; item < items; item += kTtEpiThreads) {
              const int m = item % N;                         // lanes of a warp share (feature quad, graph): broadcast reads
              const int rest = item / N;
              const int g = rest % TG, c4 = rest / TG;
              const uint32_t k0 = mask_s[g * N + m];
              const float* src = scr + ((size_t)c4 * kTtRows + g * N) * 4;
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int n = 0; n < N; ++n) {
                if ((k0 >> n) & 1u) {
                  const float4 t = *reinterpret_cast<const float4*>(src + n * 4);
                  a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                }
              }
              *reinterpret_cast<uint2*>(planes + (size_t)(st.out2_plane + (c4 >> 1)) * kTtPlaneBytes + (g * N + m) * 16 + (c4 & 1) * 8) =
                  make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
            }
          }
        } else if (st.kind == TT_Q) {
          if (train) {
            if (cq == 0) {
              float dq[8];
              float hub = 0.f;
              const bool valid = row < R;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                dq[i] = 0.f;
                if (i < CH && valid) {
                  const float e = (__uint_as_float(acc[0][i]) + bs[i]) - ycur[i];
                  const float ae = fabsf(e);
                  const float quad = fminf(ae, 1.f);
                  hub += 0.5f * quad * quad + (ae - quad);
                  dq[i] = fminf(fmaxf(e, -1.f), 1.f) * inv_cnt;
                }
              }
              rowloss[row] = hub;
              *reinterpret_cast<uint4*>(planes + (size_t)st.out_plane * kTtPlaneBytes + row * 16) = pack8(dq);
            }
            tt_epi_barrier();
            if (tid < N) {                                    // fixed order: bit-stable per-head sums
              float sacc = 0.f;
              for (int g = 0; g < TG; ++g) sacc += rowloss[g * N + tid];
              hl_s[tid] += sacc;
            }
          } else if (cq == 0 && row < R) {
            float* qd = q_out + ((size_t)g0 * N + row) * CH;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < CH) qd[i] = __uint_as_float(acc[0][i]) + bs[i];
          }
        } else if (st.kind == TT_DGRAD_MLP) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c8 = cq + 4 * j;
            if (c8 < nch_ld) {
              const uint4 gt = *reinterpret_cast<const uint4*>(planes + (size_t)(st.gate_plane + c8) * kTtPlaneBytes + row * 16);
              const uint32_t gw[4] = {gt.x, gt.y, gt.z, gt.w};
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = bf16_pos(gw[i >> 1], i & 1) ? __uint_as_float(acc[j][i]) : 0.f;
              *reinterpret_cast<uint4*>(planes + (size_t)(st.out_plane + c8) * kTtPlaneBytes + row * 16) = pack8(v);
            }
          }
        } else {   // TT_DGRAD_AGG: accumulator = [dh (F) | dagg (F)] -> dz = gate * (dh + Agg^T dagg)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c8 = cq + 4 * j;
            if (c8 < nch_ld) {
              *reinterpret_cast<float4*>(scr + ((size_t)(2 * c8) * kTtRows + row) * 4) =
                  make_float4(__uint_as_float(acc[j][0]), __uint_as_float(acc[j][1]), __uint_as_float(acc[j][2]), __uint_as_float(acc[j][3]));
              *reinterpret_cast<float4*>(scr + ((size_t)(2 * c8 + 1) * kTtRows + row) * 4) =
                  make_float4(__uint_as_float(acc[j][4]), __uint_as_float(acc[j][5]), __uint_as_float(acc[j][6]), __uint_as_float(acc[j][7]));
            }
          }
          tt_epi_barrier();
          const int FQ = F >> 2;
          const int items = FQ * TG * N;
          const uint32_t* om = mask_s + kTtRows;
          for (int item = tid; item < items; item += kTtEpiThreads) {
            const int n = item % N;
            const int rest = item / N;
            const int g = rest % TG, c4 = rest / TG;
            const int r = g * N + n;
            const uint32_t k0 = om[r];                        // bit m = Adj[n][m]: whom n contributes to
            float4 a = *reinterpret_cast<const float4*>(scr + ((size_t)c4 * kTtRows + r) * 4);
            const float* src = scr + ((size_t)(FQ + c4) * kTtRows + g * N) * 4;
            for (int m = 0; m < N; ++m) {
              if ((k0 >> m) & 1u) {
                const float4 t = *reinterpret_cast<const float4*>(src + m * 4);
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
              }
            }
            if (st.gate_plane >= 0) {
              const uint2 gt = *reinterpret_cast<const uint2*>(planes + (size_t)(st.gate_plane + (c4 >> 1)) * kTtPlaneBytes + r * 16 + (c4 & 1) * 8);
              a.x = bf16_pos(gt.x, 0) ? a.x : 0.f; a.y = bf16_pos(gt.x, 1) ? a.y : 0.f;
              a.z = bf16_pos(gt.y, 0) ? a.z : 0.f; a.w = bf16_pos(gt.y, 1) ? a.w : 0.f;
            }
            *reinterpret_cast<uint2*>(planes + (size_t)(st.out_plane + (c4 >> 1)) * kTtPlaneBytes + r * 16 + (c4 & 1) * 8) =
                make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
          }
        }
        if (s + 1 < n_steps) {                                // hand the next step's operands to the MMA warp
          fence_async_smem();
          tt_fence_before();
          __syncwarp();
          if (lane == 0) tt_mbar_arrive(&ops_bar);
        }
      }
    }
    if (train) {
      // ---- weight gradients: TMEM -> this CTA's partial row (parameter layout), once per kernel
      mbar_wait(&wg_bar, pw);
      tt_fence_after();
      const long n_params = P->n_params;
      float* dst = partial + (size_t)blockIdx.x * (size_t)(n_params + kTtPartialTail);
      for (long i = tid; i < n_params; i += kTtEpiThreads) dst[i] = 0.f;      // dead parameters (stage-0 neighbour rows)
      tt_epi_barrier();
      const int n_blocks = P->n_blocks;
      for (int b = 0; b < n_blocks; ++b) {
        const TtBlock& blk = P->blocks[b];
        const int n = blk.n, nch = (n + 7) >> 3;
        const int d = blk.dst[row];
        for (int c = cq; c < nch; c += 4) {
          uint32_t r8[8];
          tt_ld8_issue(t_lane + (uint32_t)(blk.tcol + c * 8), r8);
          tt_wait_ld();
          if (d >= 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c * 8 + i < n) dst[d + c * 8 + i] = __uint_as_float(r8[i]);
          }
        }
      }
      if (tid < kTtPartialTail) dst[n_params + tid] = (tid < N) ? hl_s[tid] * inv_cnt : 0.f;   // per-head Huber sums
    }
  }
  tt_fence_before();
  __syncthreads();
  if (warp == 0) tt_tmem_dealloc(tmem_base, kTtTmemCols);
}

}  // namespace

// ---------------------------------------------------------------------------
// host side: plan builder
// ---------------------------------------------------------------------------
int tt_build_plan(const TtShape& s, TtPlan* out) {
  TtPlan& P = *out;
  P = TtPlan{};
  V2V_REQUIRE(s.N >= 1 && s.N <= 32, "bf16 tensor-core brain: N=%d outside [1,32]", s.N);
  V2V_REQUIRE(s.F == 16, "bf16 tensor-core brain: feedback width %d (needs 16)", s.F);
  V2V_REQUIRE(s.S >= 1 && s.S <= 3, "bf16 tensor-core brain: %d stages (supports 1..3)", s.S);
  V2V_REQUIRE(s.CH >= 1 && s.CH <= 8, "bf16 tensor-core brain: %d channels (supports <= 8)", s.CH);
  V2V_REQUIRE(s.H1 >= 8 && s.H2 >= 8 && s.H3 >= 8 && s.H1 <= 128 && s.H2 <= 128 && s.H3 <= 128,
              "bf16 tensor-core brain: hidden widths outside [8,128]");
  const int F = s.F, Dn = s.Dn, De = s.De, S = s.S, CH = s.CH;
  auto r16 = [](int v) { return (v + 15) & ~15; };
  auto planes_of = [](int v) { return (v + 7) / 8; };
  const int FP = F / 8;
  int XP = planes_of(Dn + De + 1);
  if (XP & 1) ++XP;
  V2V_REQUIRE(XP * kTtRows <= kTtEpiThreads, "bf16 tensor-core brain: %d input planes unsupported", XP);
  const int P1 = planes_of(s.H1), P2 = planes_of(s.H2), P3 = planes_of(s.H3), PQ = 1;
  P.N = s.N; P.TG = kTtRows / s.N; P.Dn = Dn; P.De = De; P.F = F; P.CH = CH; P.S = S;
  P.XP = XP; P.FP = FP; P.n_params = (long)s.n_params;
  P.ones_feature = XP * 8 - 1;
  V2V_REQUIRE(P.ones_feature >= Dn + De, "bf16 tensor-core brain: no room for the ones feature");
  // ---- planes (unified index: activations, then gradients)
  const int px0 = 0;
  std::vector<int> ph(S), pa(S), pdz(S);
  for (int st = 0; st < S; ++st) { ph[st] = XP + 2 * FP * st; pa[st] = ph[st] + FP; }
  const int pm1 = XP + 2 * FP * S, pm2 = pm1 + P1, pm3 = pm2 + P2, XT = pm3 + P3;
  for (int st = 0; st < S; ++st) pdz[st] = XT + FP * st;
  const int pdm1 = XT + FP * S, pdm2 = pdm1 + P1, pdm3 = pdm2 + P2, pdq = pdm3 + P3;
  const int DT = FP * S + P1 + P2 + P3 + PQ;
  const int DTp = ((DT + 1) & 1) ? DT + 2 : DT + 1;           // at least one all-zero plane behind dq, even total
  P.x_planes = XT; P.dz_planes = DTp; P.plane_dq = pdq;
  V2V_REQUIRE(XP + 2 * FP * S <= 16 && P1 + P2 <= 16 && P3 <= 16, "bf16 tensor-core brain: weight-gradient blocks exceed 128 lanes");
  V2V_REQUIRE(pm3 + 16 <= XT + DTp && pm1 + 16 <= XT + DTp, "bf16 tensor-core brain: weight-gradient operand block leaves the planes");
  V2V_REQUIRE(DTp * 8 <= 256, "bf16 tensor-core brain: %d gradient columns exceed one MMA", DTp * 8);
  // ---- weight image
  const int hid[4] = {s.H1, s.H2, s.H3, CH};
  const int n_layers = S + 4;
  P.n_layers = n_layers;
  int w_elems = 0, bias_floats = 0;
  for (int l = 0; l < n_layers; ++l) {
    TtLayerImg& L = P.layers[l];
    for (int k = 0; k < kTtMaxK; ++k) L.kmap[k] = -1;
    L.pw_off = (int)s.w_off[l]; L.pb_off = (int)s.b_off[l];
    if (l < S) {
      L.No = F; L.Npad = r16(F);
      if (l == 0) {                    // rows of W: node (Dn), edge (De), neighbour (F: all-zero input, skipped)
        L.Kpad = XP * 8;
        for (int f = 0; f < Dn + De; ++f) L.kmap[f] = (short)f;
      } else {                         // rows of W: [h | node] (F + Dn), edge (De), aggregated (F); image K = [x0 | h | agg]
        L.Kpad = XP * 8 + 2 * F;
        for (int f = 0; f < Dn + De; ++f) L.kmap[f] = (short)(F + f);
        for (int j = 0; j < F; ++j) { L.kmap[XP * 8 + j] = (short)j; L.kmap[XP * 8 + F + j] = (short)(F + Dn + De + j); }
      }
    } else {
      const int j = l - S;
      L.No = hid[j]; L.Npad = r16(hid[j]);
      if (j == 0) {                    // rows of W: node (Dn), h (F), aggregated (F)   ([node | h | agg], :175)
        L.Kpad = XP * 8 + 2 * F;
        for (int f = 0; f < Dn; ++f) L.kmap[f] = (short)f;
        for (int i = 0; i < F; ++i) { L.kmap[XP * 8 + i] = (short)(Dn + i); L.kmap[XP * 8 + F + i] = (short)(Dn + F + i); }
      } else {
        L.Kpad = r16(planes_of(hid[j - 1]) * 8);
        for (int k = 0; k < hid[j - 1]; ++k) L.kmap[k] = (short)k;
      }
    }
    V2V_REQUIRE(L.Kpad <= kTtMaxK && L.Npad <= kTtWorkCols && L.Kpad <= kTtWorkCols, "bf16 tensor-core brain: layer %d too wide", l);
    L.w_off = w_elems; w_elems += L.Kpad * L.Npad;
    L.bias_off = bias_floats; bias_floats += L.Npad;
  }
  P.w_elems = w_elems;
  P.bias_floats = (bias_floats + 3) & ~3;
  // ---- shared memory carve-up
  int off = 0;
  auto take = [&](int bytes, int align) { off = (off + align - 1) & ~(align - 1); const int at = off; off += bytes; return at; };
  P.off_w = take(w_elems * 2, 128);
  P.off_bias = take(P.bias_floats * 4, 16);
  V2V_REQUIRE(P.off_bias == P.off_w + w_elems * 2, "bf16 tensor-core brain: weight image not contiguous");
  P.off_planes = take((XT + DTp) * kTtPlaneBytes, 128);
  P.off_scr = take(2 * F / 4 * kTtRows * 16, 128);            // [dh | dagg] fp32: 2F columns
  P.off_mask = take(2 * kTtRows * 4, 16);
  P.off_misc = take((32 + kTtRows) * 4, 16);
  P.off_tab = take(kTtMaxMma * (int)sizeof(TtMma) + kTtMaxSteps * (int)sizeof(TtStep), 16);
  P.smem_bytes = off;
  V2V_REQUIRE(P.smem_bytes <= 226 * 1024, "bf16 tensor-core brain: %d bytes of shared memory do not fit", P.smem_bytes);
  // ---- MMA and step tables
  int n_mma = 0, n_steps = 0;
  bool overflow = false;
  auto plane_off = [&](int plane) { return (uint32_t)(P.off_planes + plane * kTtPlaneBytes); };
  auto idesc = [](int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(kTtRows >> 4) << 24);
  };
  auto add_mma = [&](TtMma m) { if (n_mma >= kTtMaxMma) { overflow = true; return; } P.mma[n_mma++] = m; };
  auto add_step = [&](TtStep st) { if (n_steps >= kTtMaxSteps) { overflow = true; return; } P.steps[n_steps++] = st; };
  // forward contraction of layer l over a list of plane pairs
  auto fwd_mmas = [&](int l, const std::vector<int>& pair_first) {
    const TtLayerImg& L = P.layers[l];
    for (size_t ks = 0; ks < pair_first.size(); ++ks) {
      TtMma m{};
      m.a_off = plane_off(pair_first[ks]); m.a_lbo = kTtPlaneBytes; m.a_sbo = 128;
      m.b_off = (uint32_t)(P.off_w + (L.w_off + (int)(2 * ks) * L.Npad * 8) * 2); m.b_lbo = (uint32_t)(L.Npad * 16); m.b_sbo = 128;
      m.idesc = idesc(L.Npad, 0, 0); m.dcol = 0; m.acc = ks > 0 ? 1u : 0u;
      add_mma(m);
    }
  };
  // data gradient of layer l: dX[:, k-planes kp0 .. kp0 + nkp) = dZ (planes from dz_plane) . W^T
  auto dgrad_mmas = [&](int l, int dz_plane, int kp0, int nkp) {
    const TtLayerImg& L = P.layers[l];
    const int ksteps = L.Npad / 16;
    for (int ks = 0; ks < ksteps; ++ks) {
      TtMma m{};
      m.a_off = plane_off(dz_plane + 2 * ks); m.a_lbo = kTtPlaneBytes; m.a_sbo = 128;
      m.b_off = (uint32_t)(P.off_w + (L.w_off + kp0 * L.Npad * 8) * 2 + ks * 256); m.b_lbo = 128; m.b_sbo = (uint32_t)(L.Npad * 16);
      m.idesc = idesc(nkp * 8, 0, 1); m.dcol = 0; m.acc = ks > 0 ? 1u : 0u;
      add_mma(m);
    }
  };
  auto pairs = [](int first, int count) { std::vector<int> v; for (int i = 0; i < count; i += 2) v.push_back(first + i); return v; };
  auto cat = [](std::vector<int> a, const std::vector<int>& b) { a.insert(a.end(), b.begin(), b.end()); return a; };
  for (int st = 0; st < S; ++st) {
    TtStep e{};
    e.kind = TT_COMBINE; e.mma0 = n_mma;
    fwd_mmas(st, st == 0 ? pairs(px0, XP) : cat(cat(pairs(px0, XP), pairs(ph[st - 1], FP)), pairs(pa[st - 1], FP)));
    e.n_mma = n_mma - e.mma0; e.npad = P.layers[st].Npad; e.n_chunks = FP; e.out_plane = ph[st]; e.out2_plane = pa[st];
    e.gate_plane = -1; e.bias_off = P.layers[st].bias_off; e.relu = st < S - 1;
    add_step(e);
  }
  const int pm[3] = {pm1, pm2, pm3}, PM[3] = {P1, P2, P3};
  for (int j = 0; j < 4; ++j) {
    const int l = S + j;
    TtStep e{};
    e.kind = j < 3 ? TT_MLP : TT_Q; e.mma0 = n_mma;
    fwd_mmas(l, j == 0 ? cat(cat(pairs(px0, XP), pairs(ph[S - 1], FP)), pairs(pa[S - 1], FP)) : pairs(pm[j - 1], P.layers[l].Kpad / 8));
    e.n_mma = n_mma - e.mma0; e.npad = P.layers[l].Npad;
    e.n_chunks = j < 3 ? PM[j] : 1; e.out_plane = j < 3 ? pm[j] : pdq; e.out2_plane = -1; e.gate_plane = -1;
    e.bias_off = P.layers[l].bias_off; e.relu = j < 3;
    add_step(e);
  }
  P.n_steps_fwd = n_steps;
  {
    const int pdm[3] = {pdm1, pdm2, pdm3};
    for (int j = 3; j >= 1; --j) {          // MLP layer S + j: dz = its output gradient, result = gated gradient of its input m_j
      const int l = S + j;
      TtStep e{};
      e.kind = TT_DGRAD_MLP; e.mma0 = n_mma;
      dgrad_mmas(l, j == 3 ? pdq : pdm[j], 0, P.layers[l].Kpad / 8);
      e.n_mma = n_mma - e.mma0; e.npad = P.layers[l].Kpad; e.n_chunks = PM[j - 1]; e.out_plane = pdm[j - 1]; e.out2_plane = -1;
      e.gate_plane = pm[j - 1]; e.bias_off = 0; e.relu = 0;
      add_step(e);
    }
    {                                        // first MLP layer: [dh | dagg] of the last stage (linear: no gate)
      TtStep e{};
      e.kind = TT_DGRAD_AGG; e.mma0 = n_mma;
      dgrad_mmas(S, pdm1, XP, 2 * FP);
      e.n_mma = n_mma - e.mma0; e.npad = 2 * F; e.n_chunks = FP; e.out_plane = pdz[S - 1]; e.out2_plane = -1;
      e.gate_plane = -1; e.bias_off = 0; e.relu = 0;
      add_step(e);
    }
    for (int st = S - 1; st >= 1; --st) {    // stage st: [dh | dagg] of stage st - 1 (relu: gated by h(st - 1))
      TtStep e{};
      e.kind = TT_DGRAD_AGG; e.mma0 = n_mma;
      dgrad_mmas(st, pdz[st], XP, 2 * FP);
      e.n_mma = n_mma - e.mma0; e.npad = 2 * F; e.n_chunks = FP; e.out_plane = pdz[st - 1]; e.out2_plane = -1;
      e.gate_plane = ph[st - 1]; e.bias_off = 0; e.relu = 0;
      add_step(e);
    }
    // weight-gradient chains over the tile's 128 rows (8 k-steps of 16 rows), both operands MN-major
    const int nb2 = (P2 + P3 + 1) & ~1;
    const int tcol1 = kTtWorkCols, tcol2 = tcol1 + DTp * 8, tcol3 = tcol2 + nb2 * 8;
    V2V_REQUIRE(tcol3 + 16 <= (int)kTtTmemCols, "bf16 tensor-core brain: %d tensor-memory columns needed", tcol3 + 16);
    TtStep e{};
    e.kind = TT_WGRAD; e.mma0 = n_mma;
    struct Chain { int a_plane, b_plane, nb, tcol; } chains[3] = {{px0, XT, DTp, tcol1}, {pm1, pdm2, nb2, tcol2}, {pm3, pdq, 2, tcol3}};
    for (const Chain& c : chains)
      for (int ks = 0; ks < kTtRows / 16; ++ks) {
        TtMma m{};
        m.a_off = plane_off(c.a_plane) + ks * 256; m.a_lbo = 128; m.a_sbo = kTtPlaneBytes;
        m.b_off = plane_off(c.b_plane) + ks * 256; m.b_lbo = 128; m.b_sbo = kTtPlaneBytes;
        m.idesc = idesc(c.nb * 8, 1, 1); m.dcol = (uint32_t)c.tcol; m.acc = ks > 0 ? 1u : 2u;
        add_mma(m);
      }
    e.n_mma = n_mma - e.mma0; e.gate_plane = -1; e.out_plane = e.out2_plane = -1;
    add_step(e);
    // ---- read-out table: which accumulator lane / column block is which parameter
    int nb = 0;
    auto new_block = [&](int tcol, int n) -> TtBlock* {
      if (nb >= kTtMaxBlocks) { overflow = true; return &P.blocks[0]; }
      TtBlock* b = &P.blocks[nb++];
      b->tcol = tcol; b->n = n;
      for (int i = 0; i < kTtRows; ++i) b->dst[i] = -1;
      return b;
    };
    const int ones = P.ones_feature;
    for (int st = 0; st < S; ++st) {
      TtBlock* b = new_block(tcol1 + (pdz[st] - XT) * 8, F);
      const int w = (int)s.w_off[st];
      for (int f = 0; f < Dn + De; ++f) b->dst[f] = w + (st == 0 ? f : F + f) * F;
      if (st >= 1)
        for (int j = 0; j < F; ++j) {
          b->dst[ph[st - 1] * 8 + j] = w + j * F;
          b->dst[pa[st - 1] * 8 + j] = w + (F + Dn + De + j) * F;
        }
      b->dst[ones] = (int)s.b_off[st];
    }
    {
      TtBlock* b = new_block(tcol1 + (pdm1 - XT) * 8, s.H1);
      const int w = (int)s.w_off[S];
      for (int f = 0; f < Dn; ++f) b->dst[f] = w + f * s.H1;
      for (int j = 0; j < F; ++j) {
        b->dst[ph[S - 1] * 8 + j] = w + (Dn + j) * s.H1;
        b->dst[pa[S - 1] * 8 + j] = w + (Dn + F + j) * s.H1;
      }
      b->dst[ones] = (int)s.b_off[S];
    }
    { TtBlock* b = new_block(tcol1 + (pdm2 - XT) * 8, s.H2); b->dst[ones] = (int)s.b_off[S + 1]; }
    { TtBlock* b = new_block(tcol1 + (pdm3 - XT) * 8, s.H3); b->dst[ones] = (int)s.b_off[S + 2]; }
    { TtBlock* b = new_block(tcol1 + (pdq - XT) * 8, CH); b->dst[ones] = (int)s.b_off[S + 3]; }
    { TtBlock* b = new_block(tcol2, s.H2); for (int j = 0; j < s.H1; ++j) b->dst[j] = (int)s.w_off[S + 1] + j * s.H2; }
    { TtBlock* b = new_block(tcol2 + P2 * 8, s.H3); for (int j = 0; j < s.H2; ++j) b->dst[P1 * 8 + j] = (int)s.w_off[S + 2] + j * s.H3; }
    { TtBlock* b = new_block(tcol3, CH); for (int j = 0; j < s.H3; ++j) b->dst[j] = (int)s.w_off[S + 3] + j * CH; }
    P.n_blocks = nb;
  }
  P.n_steps_train = n_steps;
  P.n_mma = n_mma;
  V2V_REQUIRE(!overflow, "bf16 tensor-core brain: plan tables overflow");
  return 0;
}

int tt_grid(const TtPlan& p, int B) { return std::max(1, std::min(ceil_div(B, p.TG), sm_count())); }

int tt_launch(const TtPlan& ph, const TtPlan* plan_dev, const float* params, void* wimg, const float* node, const float* edge,
              const uint32_t* in_mask, const uint32_t* out_mask, const float* y, float* q_out, float* partial_dev, int B,
              int train, cudaStream_t st) {
  static int smem_set = 0;
  if (ph.smem_bytes > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(tt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ph.smem_bytes));
    smem_set = ph.smem_bytes;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  {
    cudaLaunchConfig_t ls{};
    ls.gridDim = dim3(16);
    ls.blockDim = dim3(256);
    ls.stream = st;
    ls.attrs = attr;
    ls.numAttrs = 1;
    V2V_CHECK_CUDA(cudaLaunchKernelEx(&ls, tt_stage_weights_kernel, plan_dev, params, reinterpret_cast<__nv_bfloat16*>(wimg)));
    if (int rc = launch_status("tt_stage_weights_kernel")) return rc;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(tt_grid(ph, B));
  lc.blockDim = dim3(kTtThreads);
  lc.dynamicSmemBytes = ph.smem_bytes;
  lc.stream = st;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const float inv_cnt = 1.f / ((float)B * (float)ph.CH);
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, tt_kernel, plan_dev, (const uint8_t*)wimg, node, edge, in_mask, out_mask, y, q_out,
                                    partial_dev, B, train, inv_cnt));
  return launch_status("tt_kernel");
}

}  // namespace v2v
