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
// The neighbour aggregation (AggLayer.call, :69-76) is a contraction as well once the tile's adjacency is written as
// ONE block-diagonal 0/1 operand  adj[m][n]  (bf16, 16 planes, rebuilt per tile from the in_mask words: <= 4 sixteen-byte
// entries per row):  agg = adj . h  reads it K-major with the h planes as the MN-major B operand, the backward
// dh += adj^T . dagg  reads the SAME bytes MN-major and accumulates onto the dh columns the data-gradient MMA just
// produced -- 0/1 times bf16 is exact, the sum is fp32 in TMEM.  No CUDA-core gather loop, no second mask orientation.
//
// Weight gradients never leave tensor memory during the kernel: three accumulator groups
//   chain 1: [x0 | h0 | a0 | h1 | a1 | h2 | a2 | ..]^T  x  [dz0 | dz1 | dz2 | dm1 | dm2 | dm3 | dq]   (128 x 208 at S = 3)
//   chain 2: [m1 | m2 | ..]^T x [dm2 | dm3],    chain 3: [m3 | ..]^T x [dq | 0]
// are accumulated over ALL tiles of the CTA and read out once, at the end, into the CTA's partial row; the wanted blocks
// (layer l: inputs of l x dz of l) are picked by a lane / column table, the rest of the cross product is ignored.  x0
// carries a ones feature, so the bias gradients are row 15 of chain 1.  Chain 1 is issued in two column ranges (the MLP's
// dz columns as soon as dm1 exists, the stages' dz columns at the very end): 32 MMAs per tile, 24 of them issued right
// after the commit of a backward step, so that the tensor pipe runs them under that step's epilogue.  The per-CTA
// partials (+ per-head Huber sums in the row tail) go through the same reduce + Keras-Adam kernel as the FP32 path.
//
// Rounding points (restated in oracle/bf16_emul.py): contraction operands are bf16 (RNE) -- weights, inputs, h, agg, the
// MLP activations, dq and every back-propagated dz / dagg -- accumulation is fp32 in TMEM; bias, ReLU, the output layer
// and the Huber head are fp32; master weights, gradients and Adam are fp32.  Every CTA converts the fp32 master weights
// into its bf16 image itself (35 KB read from L2 once per launch): no staging kernel, no global image.
// Warp roles: warp 12 issues MMAs from a per-tile table, 12 epilogue warps (4 TMEM lane quarters x 3 column groups)
// run the epilogues (TMEM -> registers -> bf16 planes, all lane-local); one tile in flight, the next tile's inputs wait
// in registers.
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

// optional clock trace of CTA 0's first two tiles (profiling aid): per step 8 stamps
//   [0] MMA thread: operands ready seen  [1] MMAs issued + committed
//   [2] epilogue thread 0: accumulator ready seen  [3] TMEM loads back  [4] planes written  [5] fenced + arrived
__device__ long long* g_tt_trace = nullptr;

__global__ void __launch_bounds__(kTtThreads, 1)
tt_kernel(const TtPlan* __restrict__ P, const float* __restrict__ params, const float* __restrict__ node,
          const float* __restrict__ edge, const uint32_t* __restrict__ in_mask, const float* __restrict__ y,
          float* __restrict__ q_out, float* __restrict__ partial, int B, int train, float inv_cnt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t ops_bar;          // operands of the next step are in place (16 warp arrivals)
  __shared__ __align__(8) uint64_t acc_bar;          // the step's accumulator is complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t wg_bar;           // the tile's weight-gradient MMAs have consumed the planes
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P->N, TG = P->TG, Dn = P->Dn, De = P->De, CH = P->CH, XP = P->XP;
  const int n_steps = train ? P->n_steps_train : P->n_steps_fwd;
  const int ones_f = P->ones_feature;
  float* bias_s = reinterpret_cast<float*>(smem + P->off_bias);
  uint8_t* planes = smem + P->off_planes;
  float* hl_s = reinterpret_cast<float*>(smem + P->off_misc);        // per-head Huber sums [32]
  float* rowloss = hl_s + 32;                                        // [128]
  TtMmaRt* mma_s = reinterpret_cast<TtMmaRt*>(smem + P->off_tab);
  TtStep* step_s = reinterpret_cast<TtStep*>(mma_s + kTtMaxMma);

  if (tid == 0) {
    mbar_init(&ops_bar, kTtEpiThreads / 32);
    mbar_init(&acc_bar, 1);
    mbar_init(&wg_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tt_tmem_alloc(&tmem_base_s, kTtTmemCols);
  {                                                           // planes, loss sums: finite (zero) everywhere; tables
    uint32_t* z = reinterpret_cast<uint32_t*>(planes);
    const int nz = (P->off_tab - P->off_planes) / 4;
    for (int i = tid; i < nz; i += kTtThreads) z[i] = 0u;
    const uint32_t sbase0 = smem_u32(smem);
    for (int i = tid; i < P->n_mma; i += kTtThreads) {        // MMA table: descriptors ready to issue
      const TtMma m = P->mma[i];
      TtMmaRt r;
      r.a_desc = tt_desc(sbase0 + m.a_off, m.a_lbo, m.a_sbo);
      r.b_desc = tt_desc(sbase0 + m.b_off, m.b_lbo, m.b_sbo);
      r.idesc = m.idesc; r.dcol = m.dcol; r.acc = m.acc; r.pad = 0u;
      mma_s[i] = r;
    }
    const int n_step_w = kTtMaxSteps * (int)(sizeof(TtStep) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P->steps);
    uint32_t* dst = reinterpret_cast<uint32_t*>(step_s);
    for (int i = tid; i < n_step_w; i += kTtThreads) dst[i] = src[i];
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");          // the parameters come from the preceding optimiser kernel
  // ---- fp32 master weights -> this CTA's bf16 image: entry (k-plane kp, output o) = 8 consecutive k of one o (16 bytes)
  {
    const int n_layers = P->n_layers;
    for (int l = 0; l < n_layers; ++l) {
      const TtLayerImg& L = P->layers[l];
      const int Npad = L.Npad, No = L.No, entries = (L.Kpad >> 3) * Npad;
      const float* Wp = params + L.pw_off;
      uint8_t* img = smem + P->off_w + (size_t)L.w_off * 2;
      for (int e = tid; e < entries; e += kTtThreads) {
        const int kp = e / Npad, o = e - kp * Npad;           // consecutive threads read consecutive o of the same W rows
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int wrow = L.kmap[kp * 8 + j];
          v[j] = (wrow >= 0 && o < No) ? __ldg(Wp + wrow * No + o) : 0.f;
        }
        *reinterpret_cast<uint4*>(img + (size_t)e * 16) = pack8(v);
      }
      for (int o = tid; o < Npad; o += kTtThreads) bias_s[L.bias_off + o] = (o < No) ? __ldg(params + L.pb_off + o) : 0.f;
    }
  }
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
      long long* trace = (blockIdx.x == 0) ? g_tt_trace : nullptr;
      int tile_i = 0;
      // one MMA: the table entry is already in registers when the barrier opens (loaded before the wait / one entry ahead)
      auto issue = [&](const uint4& lo, const uint4& hi) {
        const uint64_t da = ((uint64_t)lo.y << 32) | lo.x, db = ((uint64_t)lo.w << 32) | lo.z;
        const uint32_t acc = (hi.z == 0u) ? 0u : ((hi.z == 1u) ? 1u : (first ? 0u : 1u));
        tt_mma_bf16(tmem_base + hi.y, da, db, hi.x, acc);
      };
      auto run = [&](int m0, int n) {                         // MMAs m0 .. m0 + n - 1, next entry prefetched during the issue
        const uint4* tab = reinterpret_cast<const uint4*>(mma_s);
        uint4 lo = tab[2 * m0], hi = tab[2 * m0 + 1];
        for (int i = 0; i < n; ++i) {
          const int nx = (i + 1 < n) ? m0 + i + 1 : m0;
          const uint4 nlo = tab[2 * nx], nhi = tab[2 * nx + 1];
          issue(lo, hi);
          lo = nlo; hi = nhi;
        }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_i) {
        for (int s = 0; s < n_steps; ++s) {
          const int m0 = step_s[s].mma0, n = step_s[s].n_mma, p0 = step_s[s].post0, np = step_s[s].n_post;
          const int kind = step_s[s].kind;
          const uint4* tab = reinterpret_cast<const uint4*>(mma_s);
          uint4 lo = tab[2 * m0], hi = tab[2 * m0 + 1];       // first entry in registers before the barrier opens
          mbar_wait(&ops_bar, po);
          po ^= 1u;
          tt_fence_after();
          if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + s) * 8 + 0] = clock64();
          for (int i = 0; i < n; ++i) {
            const int nx = (i + 1 < n) ? m0 + i + 1 : m0;
            const uint4 nlo = tab[2 * nx], nhi = tab[2 * nx + 1];
            issue(lo, hi);
            lo = nlo; hi = nhi;
          }
          tt_commit(kind == TT_WGRAD ? &wg_bar : &acc_bar);
          if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + s) * 8 + 1] = clock64();
          if (np > 0) run(p0, np);                            // weight-gradient chains: the tensor pipe runs them under the epilogue
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
    // ---- this thread's entries of the block-diagonal adjacency operand: row m, k-planes kp0 + e, e = (tid >> 7) + kTtCQ i.
    //      They do not depend on the tile (same graph positions every tile): only the mask word does.
    const int am = tid & (kTtRows - 1);
    const int ag = am / N;
    const bool a_row = am < TG * N;
    const int a_kp0 = (ag * N) >> 3, a_kp1 = (ag * N + N - 1) >> 3;
    uint8_t* adj_planes = planes + (size_t)P->plane_adj * kTtPlaneBytes;
    // ---- input prefetch: the next tile's features / mask word / targets wait in registers
    float xin[8], ynext[8];
    uint32_t mk = 0u;
    auto fetch = [&](int tile) {
      const int g0 = tile * TG;
      const int R = (tile < num_tiles) ? min(TG, B - g0) * N : 0;
      const int pl = tid >> 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) { xin[j] = 0.f; ynext[j] = 0.f; }
      if (pl < XP && am < R) {
        const size_t gr = (size_t)g0 * N + am;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int f = pl * 8 + j;
          if (f < Dn) xin[j] = __ldg(node + gr * Dn + f);
          else if (f < Dn + De) xin[j] = __ldg(edge + gr * De + (f - Dn));
          else if (f == ones_f) xin[j] = 1.f;
        }
      }
      mk = (am < R) ? __ldg(in_mask + (size_t)g0 * N + am) : 0u;
      if (train && cq == 0 && row < R) {
        const float* yp = y + ((size_t)g0 * N + row) * CH;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < CH) ynext[j] = __ldg(yp + j);
      }
    };
    fetch(blockIdx.x);
    bool first = true;
    long long* trace = (blockIdx.x == 0 && tid == 0) ? g_tt_trace : nullptr;
    int tile_i = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_i) {
      const int g0 = tile * TG;
      const int R = min(TG, B - g0) * N;
      if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + kTtMaxSteps - 1) * 8 + 6] = clock64();      // tile start
      if (train && !first) {                                  // the previous tile's weight-gradient MMAs still read the planes
        mbar_wait(&wg_bar, pw);
        pw ^= 1u;
        tt_fence_after();
      }
      first = false;
      {
        const int pl = tid >> 7;
        if (pl < XP) *reinterpret_cast<uint4*>(planes + (size_t)pl * kTtPlaneBytes + am * 16) = pack8(xin);
        if (a_row) {
          for (int kp = a_kp0 + pl; kp <= a_kp1; kp += kTtCQ) {
            const int base = kp * 8 - ag * N;                 // source node of the entry's first element: in [-7, N)
            const uint32_t bits = (base >= 0) ? (mk >> base) : (mk << (-base));
            uint4 w;                                          // bit j -> bf16 1.0 (0x3F80) in element j
            w.x = ((bits & 1u) ? 0x3F80u : 0u) | ((bits & 2u) ? 0x3F800000u : 0u);
            w.y = ((bits & 4u) ? 0x3F80u : 0u) | ((bits & 8u) ? 0x3F800000u : 0u);
            w.z = ((bits & 16u) ? 0x3F80u : 0u) | ((bits & 32u) ? 0x3F800000u : 0u);
            w.w = ((bits & 64u) ? 0x3F80u : 0u) | ((bits & 128u) ? 0x3F800000u : 0u);
            *reinterpret_cast<uint4*>(adj_planes + (size_t)kp * kTtPlaneBytes + am * 16) = w;
          }
        }
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
        if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + s) * 8 + 2] = clock64();
        // ---- accumulator chunks of this thread: 8 columns each, chunk index = chunk0 + cq + 4 j
        uint32_t acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cq + kTtCQ * j;
          if (c < st.n_chunks) tt_ld8_issue(t_lane + (uint32_t)((st.chunk0 + c) * 8), acc[j]);
        }
        tt_wait_ld();
        if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + s) * 8 + 3] = clock64();
        if (st.kind == TT_PLANES) {
          // two chunks at a time: their shared-memory reads (bias vectors, gate planes) are issued before anything is consumed
          const bool has_bias = st.bias_off >= 0, has_gate = st.gate_plane >= 0;
#pragma unroll
          for (int jp = 0; jp < 4; jp += 2) {
            float4 b0[2], b1[2];
            uint4 gt[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int c = cq + kTtCQ * (jp + jj);
              b0[jj] = b1[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
              gt[jj] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);       // "positive": no gate
              if (c < st.n_chunks) {
                if (has_bias) {
                  const float4* bp = reinterpret_cast<const float4*>(bias_s + st.bias_off + (st.chunk0 + c) * 8);
                  b0[jj] = bp[0]; b1[jj] = bp[1];
                }
                if (has_gate) gt[jj] = *reinterpret_cast<const uint4*>(planes + (size_t)(st.gate_plane + c) * kTtPlaneBytes + row * 16);
              }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = jp + jj, c = cq + kTtCQ * j;
              if (c < st.n_chunks) {
                float v[8];
                v[0] = __uint_as_float(acc[j][0]) + b0[jj].x; v[1] = __uint_as_float(acc[j][1]) + b0[jj].y;
                v[2] = __uint_as_float(acc[j][2]) + b0[jj].z; v[3] = __uint_as_float(acc[j][3]) + b0[jj].w;
                v[4] = __uint_as_float(acc[j][4]) + b1[jj].x; v[5] = __uint_as_float(acc[j][5]) + b1[jj].y;
                v[6] = __uint_as_float(acc[j][6]) + b1[jj].z; v[7] = __uint_as_float(acc[j][7]) + b1[jj].w;
                if (st.relu) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                const uint32_t gw[4] = {gt[jj].x, gt[jj].y, gt[jj].z, gt[jj].w};   // relu'(saved activation): positive and non-zero
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = bf16_pos(gw[i >> 1], i & 1) ? v[i] : 0.f;
                *reinterpret_cast<uint4*>(planes + (size_t)(st.out_plane + c) * kTtPlaneBytes + row * 16) = pack8(v);
              }
            }
          }
        } else {   // TT_Q
          const float* bs = bias_s + st.bias_off;
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
        }
        if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + s) * 8 + 4] = clock64();
        if (s + 1 < n_steps) {                                // hand the next step's operands to the MMA warp
          fence_async_smem();
          tt_fence_before();
          __syncwarp();
          if (lane == 0) tt_mbar_arrive(&ops_bar);
        }
        if (trace && tile_i < 2) trace[(tile_i * kTtMaxSteps + s) * 8 + 5] = clock64();
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
        for (int c = cq; c < nch; c += kTtCQ) {
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
  V2V_REQUIRE(s.H1 >= 8 && s.H2 >= 8 && s.H3 >= 8 && s.H1 <= 32 * kTtCQ && s.H2 <= 32 * kTtCQ && s.H3 <= 32 * kTtCQ,
              "bf16 tensor-core brain: hidden widths outside [8,%d]", 32 * kTtCQ);
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
  // ---- extra planes behind the gradients: the transient dagg planes, then the block-diagonal adjacency operand
  const int pda = XT + DTp, padj = pda + FP, PT = padj + kTtRows / 8;
  P.plane_adj = padj;
  // ---- shared memory carve-up
  int off = 0;
  auto take = [&](int bytes, int align) { off = (off + align - 1) & ~(align - 1); const int at = off; off += bytes; return at; };
  P.off_w = take(w_elems * 2, 128);
  P.off_bias = take(P.bias_floats * 4, 16);
  P.off_planes = take(PT * kTtPlaneBytes, 128);
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
  auto add_step = [&](int kind, int mma0, int chunk0, int n_chunks, int out_plane, int gate_plane, int bias_off, int relu) {
    if (n_steps >= kTtMaxSteps) { overflow = true; return; }
    TtStep& e = P.steps[n_steps++];
    e = TtStep{};
    e.kind = kind; e.mma0 = mma0; e.n_mma = n_mma - mma0; e.chunk0 = chunk0; e.n_chunks = n_chunks; e.out_plane = out_plane;
    e.gate_plane = gate_plane; e.bias_off = bias_off; e.relu = relu;
  };
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
  // data gradient of layer l: dX[:, k-planes kp0 .. kp0 + nkp) = dZ (planes from dz_plane) . W^T  (weights read MN-major)
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
  // neighbour aggregation over the tile's block-diagonal adjacency: D[:, 0:F) (+)= adj (or adj^T) . (planes from src)
  auto agg_mmas = [&](int src_plane, bool transposed, bool accumulate) {
    for (int ks = 0; ks < kTtRows / 16; ++ks) {
      TtMma m{};
      if (!transposed) { m.a_off = plane_off(padj + 2 * ks); m.a_lbo = kTtPlaneBytes; m.a_sbo = 128; }
      else { m.a_off = plane_off(padj) + ks * 256; m.a_lbo = 128; m.a_sbo = kTtPlaneBytes; }
      m.b_off = plane_off(src_plane) + ks * 256; m.b_lbo = 128; m.b_sbo = kTtPlaneBytes;
      m.idesc = idesc(FP * 8, transposed ? 1 : 0, 1); m.dcol = 0; m.acc = (accumulate || ks > 0) ? 1u : 0u;
      add_mma(m);
    }
  };
  auto pairs = [](int first, int count) { std::vector<int> v; for (int i = 0; i < count; i += 2) v.push_back(first + i); return v; };
  auto cat = [](std::vector<int> a, const std::vector<int>& b) { a.insert(a.end(), b.begin(), b.end()); return a; };
  for (int st = 0; st < S; ++st) {
    int m0 = n_mma;
    fwd_mmas(st, st == 0 ? pairs(px0, XP) : cat(cat(pairs(px0, XP), pairs(ph[st - 1], FP)), pairs(pa[st - 1], FP)));
    add_step(TT_PLANES, m0, 0, FP, ph[st], -1, P.layers[st].bias_off, st < S - 1);       // h = act(combine + bias)
    m0 = n_mma;
    agg_mmas(ph[st], false, false);
    add_step(TT_PLANES, m0, 0, FP, pa[st], -1, -1, 0);                                     // agg = adj . h
  }
  const int pm[3] = {pm1, pm2, pm3}, PM[3] = {P1, P2, P3};
  for (int j = 0; j < 4; ++j) {
    const int l = S + j;
    const int m0 = n_mma;
    fwd_mmas(l, j == 0 ? cat(cat(pairs(px0, XP), pairs(ph[S - 1], FP)), pairs(pa[S - 1], FP)) : pairs(pm[j - 1], P.layers[l].Kpad / 8));
    if (j < 3) add_step(TT_PLANES, m0, 0, PM[j], pm[j], -1, P.layers[l].bias_off, 1);
    else add_step(TT_Q, m0, 0, 1, pdq, -1, P.layers[l].bias_off, 0);
  }
  P.n_steps_fwd = n_steps;
  {
    // weight-gradient chains over the tile's 128 rows (8 k-steps of 16 rows), both operands MN-major.  A chain is issued
    // right after the commit of the step that consumes its last-produced dz operand, so the tensor pipe runs it while the
    // epilogue warps work; only the chain of the GNN stages' dz (complete at the very end) is the tile's last step.
    const int nb2 = (P2 + P3 + 1) & ~1;
    const int tcol1 = kTtWorkCols, tcol2 = tcol1 + DTp * 8, tcol3 = tcol2 + nb2 * 8;
    V2V_REQUIRE(tcol3 + 16 <= (int)kTtTmemCols, "bf16 tensor-core brain: %d tensor-memory columns needed", tcol3 + 16);
    auto wgrad_chain = [&](int a_plane, int b_plane, int nbp, int tcol) {
      for (int ks = 0; ks < kTtRows / 16; ++ks) {
        TtMma m{};
        m.a_off = plane_off(a_plane) + ks * 256; m.a_lbo = 128; m.a_sbo = kTtPlaneBytes;
        m.b_off = plane_off(b_plane) + ks * 256; m.b_lbo = 128; m.b_sbo = kTtPlaneBytes;
        m.idesc = idesc(nbp * 8, 1, 1); m.dcol = (uint32_t)tcol; m.acc = ks > 0 ? 1u : 2u;
        add_mma(m);
      }
    };
    auto set_post = [&](int post0) { if (n_steps > 0) { P.steps[n_steps - 1].post0 = post0; P.steps[n_steps - 1].n_post = n_mma - post0; } };
    const int pdm[3] = {pdm1, pdm2, pdm3};
    for (int j = 3; j >= 1; --j) {          // MLP layer S + j: dz = its output gradient, result = gated gradient of its input m_j
      const int l = S + j;
      const int m0 = n_mma;
      dgrad_mmas(l, j == 3 ? pdq : pdm[j], 0, P.layers[l].Kpad / 8);
      add_step(TT_PLANES, m0, 0, PM[j - 1], pdm[j - 1], pm[j - 1], -1, 0);
      const int p0 = n_mma;
      if (j == 3) wgrad_chain(pm3, pdq, 2, tcol3);                                          // chain 3: [m3 ..]^T x [dq | 0]
      if (j == 1) wgrad_chain(pm1, pdm2, nb2, tcol2);                                       // chain 2: [m1 | m2 ..]^T x [dm2 | dm3]
      set_post(p0);
    }
    // layers that feed on [h | agg]: the first MLP layer (last stage is linear: no gate), then stages S-1 .. 1
    for (int st = S; st >= 1; --st) {
      int m0 = n_mma;
      dgrad_mmas(st, st == S ? pdm1 : pdz[st], XP, 2 * FP);                                // accumulator = [dh | dagg]
      add_step(TT_PLANES, m0, FP, FP, pda, -1, -1, 0);                                      // dagg -> bf16 planes
      if (st == S) {                                                                         // chain 1a: [x0 | h | a ..]^T x [dm1 | dm2 | dm3 | dq | 0]
        const int p0 = n_mma;
        wgrad_chain(px0, pdm1, XT + DTp - pdm1, tcol1 + (pdm1 - XT) * 8);
        set_post(p0);
      }
      m0 = n_mma;
      agg_mmas(pda, true, true);                                                             // dh += adj^T . dagg
      add_step(TT_PLANES, m0, 0, FP, pdz[st - 1], (st - 1 < S - 1) ? ph[st - 1] : -1, -1, 0);
    }
    {                                                                                        // chain 1b: [x0 | h | a ..]^T x [dz0 | .. | dz(S-1)]
      const int m0 = n_mma;
      wgrad_chain(px0, XT, FP * S, tcol1);
      add_step(TT_WGRAD, m0, 0, 0, -1, -1, -1, 0);
    }
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

int tt_set_trace(long long* dev_buf) {
  V2V_CHECK_CUDA(cudaMemcpyToSymbol(g_tt_trace, &dev_buf, sizeof(dev_buf)));
  return 0;
}

int tt_grid(const TtPlan& p, int B) { return std::max(1, std::min(ceil_div(B, p.TG), sm_count())); }

int tt_launch(const TtPlan& ph, const TtPlan* plan_dev, const float* params, const float* node, const float* edge,
              const uint32_t* in_mask, const float* y, float* q_out, float* partial_dev, int B, int train, cudaStream_t st) {
  static int smem_set = 0;
  if (ph.smem_bytes > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(tt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ph.smem_bytes));
    smem_set = ph.smem_bytes;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(tt_grid(ph, B));
  lc.blockDim = dim3(kTtThreads);
  lc.dynamicSmemBytes = ph.smem_bytes;
  lc.stream = st;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const float inv_cnt = 1.f / ((float)B * (float)ph.CH);
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, tt_kernel, plan_dev, params, node, edge, in_mask, y, q_out, partial_dev, B, train, inv_cnt));
  return launch_status("tt_kernel");
}

}  // namespace v2v
