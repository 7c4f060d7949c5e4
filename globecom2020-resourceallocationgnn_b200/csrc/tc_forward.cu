// Tensor-core forward of the shared-weight brain (BS.predict, BS_brain.py:147-200, :225-235) on tcgen05 / TMEM.
//
// The combine layers (GNNLayer.call, :44-51) and the decision MLP (:176-200) are genuine dense contractions
// [rows x K] x [K x O]; here they run on the 5th-generation tensor cores:
//   * a tile = TG = floor(128 / N) whole graphs = up to 128 node rows = the 128 TMEM lanes of one accumulator (UMMA M = 128);
//   * shared-memory operands are 4-feature PLANES  plane[k / 4][row][k % 4]  (2 KB per plane): the no-swizzle K-major
//     core-matrix layout (8 rows x 16 bytes contiguous, SBO = 128 B, LBO = one plane), so the reference's concatenations
//     ([h | agg | node,edge], :154-175) are plane lists; the weights W[k][o] are staged once per CTA in the same form with
//     o as the row, hi and lo halves of a k-plane adjacent (plane[k / 4][o | Npad + o][k % 4]).  (MN-major TF32 operands
//     without swizzle read as zeros -- scratch/tc_probe.cu -- so both operands are K-major);
//   * fp32 parity (1e-4, north star) rules out single-pass TF32, so every operand is split x = hi + lo (hi = TF32
//     truncation: one LOP; lo = x - hi: one FADD) and each k-step of 8 issues TWO tcgen05.mma.kind::tf32 into one fp32
//     accumulator: A_hi * [W_hi | W_lo] (hi*hi and hi*lo share an instruction through N-concatenation) and A_lo * W_hi;
//     the epilogue adds the two column halves.  The per-instruction cost (~75 cycles) does not depend on N here;
//   * the decision MLP never leaves tensor memory: the epilogue of layer j writes relu(acc + bias), split, back over its
//     own accumulator columns with tcgen05.st and layer j+1 takes its A operand from TMEM (TS form);
//   * warp specialisation with TWO tiles in flight: warp 16 only issues MMAs (waits on an "operands ready" mbarrier per
//     tile slot, commits to an "accumulator ready" mbarrier), the 16 epilogue warps alternate between the two slots
//     (tcgen05.ld with all loads in flight and one wait, bias, ReLU, hi/lo, the in-shared-memory neighbour aggregation
//     after a combine stage), so the tensor pipe works on one tile while the CUDA cores finish the other's epilogue;
//     each slot owns 256 of the 512 TMEM columns and its own operand planes; the next tiles' inputs wait in registers.
// Used for predict at batch sizes that give every SM several tiles; smaller batches stay on the FP32-pipe fused
// kernel, which has finer tiles (see brain.cu).
#include <algorithm>

#include "tc_forward.cuh"

namespace v2v {

namespace {

constexpr int kPlaneBytes = kTcRows * 16;      // one plane: 128 rows x 4 floats
constexpr int kPlaneFloats = kTcRows * 4;
constexpr uint32_t kTmemCols = 512;            // two tile slots x 256 columns
constexpr int kSlotCols = 256, kRegionY = 160;   // a slot: region X = columns [0, 160), region Y = [160, 256)

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
// same with the A operand in tensor memory (lane = row, column = k): the MLP chain
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float4& a, const float4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(a.x), "f"(a.y),
               "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }
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
// x = hi + lo with hi the TF32 truncation of x (one LOP) and lo = x - hi exactly (one FADD); the tensor core reads only
// the upper 19 bits of lo, so hi + tf32(lo) carries >= 21 significand bits.  (cvt.rna.tf32 is a quarter-rate
// conversion and dominated the epilogues.)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
  hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
  lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
}

// issue a TMEM load of 8 columns without waiting (pair with tc_wait_ld before the registers are read)
__device__ __forceinline__ void tc_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void padd4(float4& a, const float4& v) {
  float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(v.x, v.y));
  float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(v.z, v.w));
  a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
}

// Builds the shared-memory image of all weights (hi/lo k-planes, see below) and zero-padded biases ONCE per call in
// global memory; every CTA of the forward kernel then pulls it with one bulk-async copy instead of re-deriving it.
__global__ void __launch_bounds__(256)
tc_stage_weights_kernel(const TcPlan* __restrict__ P, const float* __restrict__ params, float* __restrict__ wimg) {
  asm volatile("griddepcontrol.wait;" ::: "memory");          // parameters may come from the preceding optimiser kernel
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int n_layers = P->n_layers;
  float* bias_img = wimg + P->w_floats;
  for (int l = 0; l < n_layers; ++l) {
    const TcLayer& L = P->layers[l];
    const int Kpad = L.Kpad, Npad = L.Npad, No = L.N;
    float* Wc = wimg + L.w_off;         // k-plane kc: rows [0, Npad) = hi(W[k][o]), rows [Npad, 2 Npad) = lo, 4 k's per row
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < Kpad * Npad; idx += gridDim.x * blockDim.x) {
      const int o = idx % Npad, k = idx / Npad;               // consecutive threads read consecutive o of one W row
      const int wrow = L.kmap[k];
      const float w = (wrow >= 0 && o < No) ? params[L.pw_off + wrow * No + o] : 0.f;
      const float h = tf32_rn(w);
      const int at = (k >> 2) * (2 * Npad * 4) + o * 4 + (k & 3);
      Wc[at] = h;
      Wc[at + Npad * 4] = tf32_rn(w - h);
    }
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < Npad; o += gridDim.x * blockDim.x)
      bias_img[L.bias_off + o] = (o < No) ? params[L.pb_off + o] : 0.f;
  }
}

struct TcSlot {                          // per tile slot: shared-memory operand planes, fp32 h planes, masks
  float* hs;
  float* stage_hi;
  float* stage_lo;
  uint32_t* mask_s;
};

__global__ void __launch_bounds__(kTcThreads, 1)
tc_forward_kernel(const TcPlan* __restrict__ P, const float* __restrict__ wimg, const float* __restrict__ node,
                  const float* __restrict__ edge, const uint32_t* __restrict__ in_mask, float* __restrict__ q_out, int B,
                  float* __restrict__ dbg, int dbg_layer) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t ops_bar[2];        // operands of the slot's next layer are in place (16 warp arrivals)
  __shared__ __align__(8) uint64_t acc_bar[2];        // the slot's accumulator is complete (tcgen05.commit)
  __shared__ __align__(8) uint64_t w_bar;             // the weight image has landed
  __shared__ uint32_t tmem_base_s;
  struct LayerRt { int Kpad, Npad, relu, a_src, out_kind, dcol, acol, w_off, bias_off; };
  __shared__ LayerRt lay[kTcMaxLayers];               // the per-layer scalars of the plan (the plan itself is in global memory)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = P->N, TG = P->TG, Dn = P->Dn, De = P->De, CH = P->CH, XP = P->x_planes, dn_pad = P->dn_pad;
  const int FP = P->F >> 2;                                   // planes of h / agg
  const int n_layers = P->n_layers;
  const int SP = P->stage_planes;                             // shared-memory operand planes per slot
  if (tid < n_layers) {
    const TcLayer& L = P->layers[tid];
    lay[tid] = LayerRt{L.Kpad, L.Npad, L.relu, L.a_src, L.out_kind, L.dcol, L.acol, L.w_off, L.bias_off};
  }

  float* Wsm = reinterpret_cast<float*>(smem);
  float* bias_s = Wsm + P->w_floats;
  TcSlot slot[2];
  {
    float* p = bias_s + P->bias_floats;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      slot[s].hs = p; p += FP * kPlaneFloats;
      slot[s].stage_hi = p; p += SP * kPlaneFloats;
      slot[s].stage_lo = p; p += SP * kPlaneFloats;
      slot[s].mask_s = reinterpret_cast<uint32_t*>(p); p += kTcRows;
    }
  }
  if (tid == 0) {
    mbar_init(&ops_bar[0], kTcEpiThreads / 32); mbar_init(&ops_bar[1], kTcEpiThreads / 32);   // one arrival per epilogue warp
    mbar_init(&acc_bar[0], 1); mbar_init(&acc_bar[1], 1);
    mbar_init(&w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, kTmemCols);
  {                                                           // operand planes finite everywhere (padding rows)
    float* z = slot[0].hs;
    const int nz = 2 * ((FP + 2 * SP) * kPlaneFloats + kTcRows);
    for (int i = tid; i < nz; i += kTcThreads) z[i] = 0.f;
  }
  __syncthreads();                                            // the mbarrier inits are visible to every waiting thread
  asm volatile("griddepcontrol.wait;" ::: "memory");          // the weight image comes from tc_stage_weights_kernel
  // ---- weights + biases: one bulk-async copy of the pre-built image (written through the async proxy: no fence needed
  //      before the tensor core reads it)
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)((P->w_floats + P->bias_floats) * sizeof(float));
    mbar_arrive_expect_tx(&w_bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u)         // in 32 KB pieces
      bulk_g2s(reinterpret_cast<uint8_t*>(Wsm) + off, reinterpret_cast<const uint8_t*>(wimg) + off, min(32768u, bytes - off), &w_bar);
  }
  mbar_wait(&w_bar, 0u);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int num_tiles = (B + TG - 1) / TG;
  const int G = gridDim.x;
  const int x_plane0 = 2 * FP;                                // x0 lives behind h and agg in the operand planes

  if (warp == kTcEpiThreads / 32) {
    // =============================== MMA warp: one thread issues everything ===============================
    if (lane == 0) {
      uint32_t po[2] = {0u, 0u};
      for (int tA = blockIdx.x; tA < num_tiles; tA += 2 * G) {
        const bool validB = tA + G < num_tiles;
        for (int l = 0; l < n_layers; ++l) {
          const int Kpad = lay[l].Kpad, Npad = lay[l].Npad, a_src = lay[l].a_src, dcol = lay[l].dcol, acol = lay[l].acol;
          const uint32_t idesc2 = umma_idesc_tf32(2 * Npad), idesc1 = umma_idesc_tf32(Npad);
          const uint32_t b_plane = (uint32_t)(2 * Npad * 16);             // one k-plane of W: 2 Npad rows x 16 bytes
          const uint64_t db0 = umma_desc(smem_u32(Wsm + lay[l].w_off), b_plane, 128);
          const uint64_t b_step = (uint64_t)((2 * b_plane) >> 4);
          const int ksteps = Kpad >> 3;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (s == 1 && !validB) break;
            mbar_wait(&ops_bar[s], po[s]);
            po[s] ^= 1u;
            tc_fence_after();
            const uint32_t tslot = tmem_base + (uint32_t)(s * kSlotCols);
            const uint32_t d_t = tslot + (uint32_t)dcol;
            uint64_t db = db0;
            // per k-step (8 of K): D[:, 0:2Npad] += A_hi * [W_hi | W_lo]  and  D[:, 0:Npad] += A_lo * W_hi
            if (a_src == 2) {                                 // operand in tensor memory (written by the previous epilogue)
              uint32_t ah = tslot + (uint32_t)acol, al = ah + (uint32_t)Kpad;
              for (int ks = 0; ks < ksteps; ++ks, ah += 8u, al += 8u, db += b_step) {
                tc_mma_tf32_ts(d_t, ah, db, idesc2, ks > 0 ? 1u : 0u);
                tc_mma_tf32_ts(d_t, al, db, idesc1, 1u);
              }
            } else {
              const uint32_t a_off = (uint32_t)((a_src == 0 ? x_plane0 : 0) * kPlaneBytes);
              uint64_t dah = umma_desc(smem_u32(slot[s].stage_hi) + a_off, kPlaneBytes, 128);
              uint64_t dal = umma_desc(smem_u32(slot[s].stage_lo) + a_off, kPlaneBytes, 128);
              const uint64_t a_step = (uint64_t)((2 * kPlaneBytes) >> 4);
              for (int ks = 0; ks < ksteps; ++ks, dah += a_step, dal += a_step, db += b_step) {
                tc_mma_tf32(d_t, dah, db, idesc2, ks > 0 ? 1u : 0u);
                tc_mma_tf32(d_t, dal, db, idesc1, 1u);
              }
            }
            tc_commit(&acc_bar[s]);
          }
        }
      }
    }
  } else {
    // =============================== 16 epilogue warps ===============================
    const int q4 = warp & 3, cq = warp >> 2;                  // TMEM lane quarter, column quarter
    const int row = q4 * 32 + lane;                           // this thread's accumulator lane = tile row
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int MS = N;                                         // aggregation: one target per item (F/4 * TG * N <= 512)
    uint32_t pa[2] = {0u, 0u};
    // ---- input prefetch: each slot's next tile (node / edge features, mask word) waits in registers
    float xin[2][4];
    uint32_t min_[2] = {0u, 0u};
    auto fetch = [&](int s, int tile) {
      const int g0 = tile * TG;
      const int R = (tile < num_tiles) ? min(TG, B - g0) * N : 0;
      const int r = tid & (kTcRows - 1), pl = tid >> 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) xin[s][j] = 0.f;
      if (pl < XP && r < R) {
        const size_t gr = (size_t)g0 * N + r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = pl * 4 + j;
          if (f < Dn) xin[s][j] = __ldg(node + gr * Dn + f);
          else if (f >= dn_pad && f - dn_pad < De) xin[s][j] = __ldg(edge + gr * De + (f - dn_pad));
        }
      }
      min_[s] = (tid < TG * N && tid < R) ? __ldg(in_mask + (size_t)g0 * N + tid) : 0u;
    };
    // prefetched inputs -> the slot's x0 operand planes (hi / lo) and masks; signals layer 0; fetches the slot's next tile
    auto start_tile = [&](int s, int next_tile) {
      const int r = tid & (kTcRows - 1), pl = tid >> 7;
      if (pl < XP) {
        float4 hi, lo;
        split4(make_float4(xin[s][0], xin[s][1], xin[s][2], xin[s][3]), hi, lo);
        *reinterpret_cast<float4*>(slot[s].stage_hi + (x_plane0 + pl) * kPlaneFloats + r * 4) = hi;
        *reinterpret_cast<float4*>(slot[s].stage_lo + (x_plane0 + pl) * kPlaneFloats + r * 4) = lo;
      }
      if (tid < TG * N) slot[s].mask_s[tid] = min_[s];
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ops_bar[s]);
      fetch(s, next_tile);
    };
    fetch(0, blockIdx.x);
    fetch(1, blockIdx.x + G);
    for (int tA = blockIdx.x; tA < num_tiles; tA += 2 * G) {
      const bool validB = tA + G < num_tiles;
      start_tile(0, tA + 2 * G);
      if (validB) start_tile(1, tA + 3 * G);
      for (int l = 0; l < n_layers; ++l) {
        const int Npad = lay[l].Npad, relu = lay[l].relu, out_kind = lay[l].out_kind, dcol = lay[l].dcol;
        const float* bs = bias_s + lay[l].bias_off;
        // columns are dealt to the 4 column quarters in 8-column chunks: chunk j of this thread = 8 * (cq + 4 j)
        const int nch = ((Npad >> 3) - cq + 3) >> 2;          // <= 3 for Npad <= 96
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s == 1 && !validB) break;
          const int tile = tA + s * G;
          const int g0 = tile * TG;
          const int R = min(TG, B - g0) * N;
          const TcSlot& S = slot[s];
          const uint32_t t_d = t_lane + (uint32_t)(s * kSlotCols + dcol);
          mbar_wait(&acc_bar[s], pa[s]);
          pa[s] ^= 1u;
          tc_fence_after();
          // ---- TMEM -> registers (all loads in flight, one wait), bias, activation, then the consumer's layout
          uint32_t acc[3][8], acc2[3][8];                    // hi*hi + lo*hi columns, hi*lo columns
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (j < nch) {
              tc_ld8_issue(t_d + (uint32_t)(8 * (cq + 4 * j)), acc[j]);
              tc_ld8_issue(t_d + (uint32_t)(Npad + 8 * (cq + 4 * j)), acc2[j]);
            }
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < nch) {
              const int c0 = 8 * (cq + 4 * j);
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[j][i]) + __uint_as_float(acc2[j][i]);
              if (dbg && dbg_layer >= 0 && l == dbg_layer && tile == 0) {   // debugging aid: raw accumulator, [128][Npad]
#pragma unroll
                for (int i = 0; i < 8; ++i) dbg[row * Npad + c0 + i] = v[i];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[i] += bs[c0 + i];
                if (relu) v[i] = fmaxf(v[i], 0.f);
              }
              const float4 x0 = make_float4(v[0], v[1], v[2], v[3]), x1 = make_float4(v[4], v[5], v[6], v[7]);
              if (out_kind == 2) {
                if (row < R) {
                  float* qd = q_out + ((size_t)g0 * N + row) * CH;
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    if (c0 + i < CH) qd[c0 + i] = v[i];
                }
              } else {
                float4 h0, l0, h1, l1;
                split4(x0, h0, l0);
                split4(x1, h1, l1);
                if (out_kind == 1) {                         // MLP chain: back into tensor memory, over the accumulator
                  tc_st8(t_d + (uint32_t)c0, h0, h1);
                  tc_st8(t_d + (uint32_t)(Npad + c0), l0, l1);
                } else {                                     // combine stage: fp32 copy for the aggregation + operand planes
                  const int pl = c0 >> 2;
                  *reinterpret_cast<float4*>(S.hs + pl * kPlaneFloats + row * 4) = x0;
                  *reinterpret_cast<float4*>(S.hs + (pl + 1) * kPlaneFloats + row * 4) = x1;
                  *reinterpret_cast<float4*>(S.stage_hi + pl * kPlaneFloats + row * 4) = h0;
                  *reinterpret_cast<float4*>(S.stage_lo + pl * kPlaneFloats + row * 4) = l0;
                  *reinterpret_cast<float4*>(S.stage_hi + (pl + 1) * kPlaneFloats + row * 4) = h1;
                  *reinterpret_cast<float4*>(S.stage_lo + (pl + 1) * kPlaneFloats + row * 4) = l1;
                }
              }
            }
          }
          // ---- neighbour aggregation after a combine stage: agg[pc][g, m] = sum_n Adj[g][n][m] * h[pc][g, n]  -> hi / lo planes
          if (out_kind == 0) {
            epi_barrier();
            const int items = FP * TG * MS;
            for (int item = tid; item < items; item += kTcEpiThreads) {
              const int ms = item % MS;                      // lanes of a warp share (plane, graph): broadcast reads
              const int rest = item / MS;
              const int g = rest % TG, pc = rest / TG;
              const uint32_t k0 = S.mask_s[g * N + ms];
              const float* src = S.hs + pc * kPlaneFloats + (g * N) * 4;
              float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f);
              int n = 0;
              for (; n + 4 <= N; n += 4) {
                const float4 y0 = *reinterpret_cast<const float4*>(src + n * 4), y1 = *reinterpret_cast<const float4*>(src + n * 4 + 4);
                const float4 y2 = *reinterpret_cast<const float4*>(src + n * 4 + 8), y3 = *reinterpret_cast<const float4*>(src + n * 4 + 12);
                const uint32_t b0 = k0 >> n;
                if (b0 & 1u) padd4(a0, y0);
                if (b0 & 2u) padd4(a0, y1);
                if (b0 & 4u) padd4(a0, y2);
                if (b0 & 8u) padd4(a0, y3);
              }
              for (; n < N; ++n)
                if (k0 & (1u << n)) padd4(a0, *reinterpret_cast<const float4*>(src + n * 4));
              float4 hi, lo;
              const int at = (FP + pc) * kPlaneFloats + (g * N + ms) * 4;
              split4(a0, hi, lo);
              *reinterpret_cast<float4*>(S.stage_hi + at) = hi;
              *reinterpret_cast<float4*>(S.stage_lo + at) = lo;
            }
          }
          if (out_kind != 2) {                                // hand the next layer's operands to the MMA warp
            if (out_kind == 1) tc_wait_st(); else fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ops_bar[s]);
          }
        }
      }
    }
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
  int prev_dcol = kRegionY;              // the first layer's accumulator goes to region X
  auto finish = [&](TcLayer& L) {
    V2V_REQUIRE(L.Kpad % 8 == 0 && L.Kpad <= kTcMaxK, "tensor-core forward: contraction length %d unsupported", L.Kpad);
    V2V_REQUIRE(L.Npad >= 16 && L.Npad <= 96, "tensor-core forward: layer width %d unsupported", L.Npad);
    L.w_off = w_floats; w_floats += 2 * L.Kpad * L.Npad;
    L.bias_off = bias_floats; bias_floats += L.Npad;
    if (L.a_src != 2) stage_planes = std::max(stage_planes, L.Kpad / 4);
    // accumulators alternate between the slot's two TMEM regions; a chain layer reads its operand where the previous
    // layer's epilogue left it (over that layer's accumulator)
    L.acol = prev_dcol;
    L.dcol = (L.a_src == 2) ? (prev_dcol == 0 ? kRegionY : 0) : 0;
    const int room = L.dcol == 0 ? kRegionY : kSlotCols - kRegionY;
    V2V_REQUIRE(2 * L.Npad <= room, "tensor-core forward: layer width %d does not fit its tensor-memory region", L.Npad);
    prev_dcol = L.dcol;
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
  V2V_REQUIRE(P.x_planes * kTcRows <= kTcEpiThreads, "tensor-core forward: %d input planes unsupported", P.x_planes);
  V2V_REQUIRE((F / 4) * P.TG * s.N <= 4 * kTcEpiThreads, "tensor-core forward: aggregation items");
  const size_t bytes = (size_t)(P.w_floats + P.bias_floats) * 4 +
                       2 * ((size_t)(F / 4 + 2 * stage_planes) * kPlaneBytes + (size_t)kTcRows * 4) + 64;
  V2V_REQUIRE(bytes <= 227 * 1024, "tensor-core forward: %zu bytes of shared memory do not fit", bytes);
  P.smem_bytes = (int)bytes;
  return 0;
}

int tc_grid(const TcPlan& p, int B) { return std::max(1, std::min(ceil_div(B, p.TG), sm_count())); }

int tc_forward_launch(const TcPlan& ph, const TcPlan* plan_dev, const float* params, float* wimg, const float* node,
                      const float* edge, const uint32_t* in_mask, float* q_out, int B, cudaStream_t st, float* dbg, int dbg_layer) {
  static int smem_set = 0;
  if (ph.smem_bytes > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ph.smem_bytes));
    smem_set = ph.smem_bytes;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  {
    cudaLaunchConfig_t ls{};
    ls.gridDim = dim3(24);
    ls.blockDim = dim3(256);
    ls.stream = st;
    ls.attrs = attr;
    ls.numAttrs = 1;
    V2V_CHECK_CUDA(cudaLaunchKernelEx(&ls, tc_stage_weights_kernel, plan_dev, params, wimg));
    if (int rc = launch_status("tc_stage_weights_kernel")) return rc;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(tc_grid(ph, B));
  lc.blockDim = dim3(kTcThreads);
  lc.dynamicSmemBytes = ph.smem_bytes;
  lc.stream = st;
  lc.attrs = attr;
  lc.numAttrs = 1;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, tc_forward_kernel, plan_dev, (const float*)wimg, node, edge, in_mask, q_out, B, dbg, dbg_layer));
  return launch_status("tc_forward_kernel");
}

}  // namespace v2v
