// Whole-network fused kernel for the shared-weight brain (G == 1, N <= 32):
//   forward (BS_brain.py:147-200) + Huber head (:86-87) + full backward in ONE launch.
//
// A CTA owns tiles of TG whole graphs (R = TG*N node rows).  Every activation of the tile lives in a
// feature-major shared-memory arena  arena[feature_row][RP]  (row index fastest), so
//   * a layer is out[o][r] = sum_k W[k][o] * in[k][r]: the thread tile (4 rows x 4..8 columns) reads ONE
//     128-bit shared-memory vector of activations and one/two broadcast vectors of weights per k and
//     issues 16..32 FFMA -- true fp32 (the 1e-4 parity bar rules out single-pass TF32);
//   * the concatenations of the reference ([h|node|edge|agg], [node|h|agg]) are just row lists;
//   * the neighbour aggregation is a register gather-reduce over the 20 rows of a graph inside the
//     arena: no HBM traffic at all (the standalone kernel in agg.cu serves the AggLayer surface);
//   * backward re-uses dead forward rows for the data gradients; the weight-gradient 4x4 blocks of
//     all layers are spread over the threads, accumulated in registers across the CTA's tiles and
//     written once as a per-CTA partial (deterministic; reduced by fused_reduce_adam).
// HBM traffic per graph (N=20): 1,040 B features + 80 B mask + 320 B targets in; nothing out but the
// 35 KB gradient partial per CTA.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "fused.cuh"

#ifndef V2V_TF32_SPLIT_RN
#define V2V_TF32_SPLIT_RN 0
#endif

namespace v2v {

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
template <int TR>
struct RowVec;
template <>
struct RowVec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <>
struct RowVec<2> {
  float v[2];
  __device__ __forceinline__ void load(const float* p) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};

// Packed fp32 FMA (fma.rn.f32x2, sm_100): two IEEE fused multiply-adds per instruction -- bit-identical to two fmaf, half the
// FMA issue slots.  The register pairs of a 128-bit shared-memory load are free operands; a broadcast operand is duplicated.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ float2 unpk2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

template <int TR, int TC, bool SLOT>
__device__ __forceinline__ void gemm_phase(const FusedOp& op, float* arena, const float* Ws, const int* tab, int RP,
                                           int TGp, int tid) {
  const int row_groups = RP / TR;
  const int items = row_groups * (op.O / TC);
  const int4* in_rows4 = reinterpret_cast<const int4*>(tab + op.in_tab);     // padded to a multiple of 4 with zero_row
  const int* out_rows = tab + op.out_tab;
  const int K4 = (op.K + 3) >> 2;
  const int O = op.O;
  for (int item = tid; item < items; item += kFusedThreads) {
    const int cg = item / row_groups, rg = item - cg * row_groups;
    const int r0 = rg * TR, o0 = cg * TC;
    const int slot = SLOT ? r0 / TGp : 0;                    // slot-major rows: the TR rows of a tile share a weight set
    float acc[TC][TR];
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const float b = Ws[op.b_off + slot * op.slot_b + o0 + j];
#pragma unroll
      for (int i = 0; i < TR; ++i) acc[j][i] = b;
    }
    const float* wp = Ws + op.w_off + slot * op.slot_w + o0;
    const float* xb = arena + r0;
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
      const int4 rows = in_rows4[k4];
      RowVec<TR> x[4];
      x[0].load(xb + rows.x * RP); x[1].load(xb + rows.y * RP); x[2].load(xb + rows.z * RP); x[3].load(xb + rows.w * RP);
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
        for (int j = 0; j < TC; ++j)
#pragma unroll
          for (int i = 0; i < TR; ++i) acc[j][i] = fmaf(x[kk].v[i], w[kk][j], acc[j][i]);
    }
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      RowVec<TR> o;
#pragma unroll
      for (int i = 0; i < TR; ++i) o.v[i] = op.relu ? fmaxf(acc[j][i], 0.f) : acc[j][i];
      o.store(arena + out_rows[o0 + j] * RP + r0);
    }
  }
}

// Neighbour aggregation inside the arena.  item = (feature f, graph g, target split ms).
template <int NMAX, bool SLOT>
__device__ __forceinline__ void agg_phase(const FusedOp& op, float* arena, const int* tab, const uint32_t* mask_s, int RP,
                                          int N, int TG, int TGp, int tid) {
  const int row_g = SLOT ? 1 : N, row_n = SLOT ? TGp : 1;      // node (g, n) is arena row g * row_g + n * row_n
  constexpr int MS = 3;
  constexpr int MT = (NMAX + MS - 1) / MS;
  const int F = op.K;
  const int items = F * TG * MS;
  for (int item = tid; item < items; item += kFusedThreads) {
    const int f = item % F;
    const int rest = item / F;
    const int g = rest % TG, ms = rest / TG;
    const float* src = arena + tab[op.in_tab + f] * RP + g * row_g;
    float v[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) v[n] = (n < N) ? src[n * row_n] : 0.f;
    float acc[MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) acc[j] = 0.f;
    // forward: in_mask[m] bit n = Adj[n][m]; transposed: out_mask[n] bit m = Adj[n][m] -- same gather-reduce
    const uint32_t* mk = (op.transposed ? mask_s + TG * N : mask_s) + g * N;
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int m = ms + j * MS;
      const uint32_t msk = (m < N) ? mk[m] : 0u;
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (msk & (1u << n)) acc[j] += v[n];
    }
    if (g == 0 && ms == 0) {                                // padding rows of the tile: exact zeros
      if (SLOT) {
        for (int n = 0; n < N; ++n)
          for (int gp = TG; gp < TGp; ++gp) arena[tab[op.out_tab + f] * RP + n * TGp + gp] = 0.f;
      } else {
        for (int r = TG * N; r < RP; ++r) arena[tab[op.out_tab + f] * RP + r] = 0.f;
      }
    }
    float* dst = arena + tab[op.out_tab + f] * RP + g * row_g;
    const float* add = op.add_tab >= 0 ? arena + tab[op.add_tab + f] * RP + g * row_g : nullptr;
    const float* gate = op.gate_tab >= 0 ? arena + tab[op.gate_tab + f] * RP + g * row_g : nullptr;
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int m = ms + j * MS;
      if (m < N) {
        float r = acc[j];
        if (add) r += add[m * row_n];
        if (gate) r = gate[m * row_n] > 0.f ? r : 0.f;
        dst[m * row_n] = r;
      }
    }
  }
}


template <bool SLOT>
__device__ __forceinline__ void loss_phase(const FusedProgram* P, float* arena, const int* tab, float* hl_s, int RP,
                                           int ng, float inv_cnt, int tid) {
  const int N = P->N, TGp = P->TGp;
  const int valid_rows = ng * N;
  const int items = P->CH * RP;
  for (int item = tid; item < items; item += kFusedThreads) {
    const int c = item / RP, r = item - c * RP;
    float* q = arena + tab[P->q_tab + c] * RP + r;
    float* yp = arena + tab[P->y_tab + c] * RP + r;
    float dq = 0.f, hub = 0.f;
    const bool valid = SLOT ? (r % TGp) < ng : r < valid_rows;
    if (valid) {
      const float e = *q - *yp;
      const float ae = fabsf(e);
      const float quad = fminf(ae, 1.f);
      hub = 0.5f * quad * quad + (ae - quad);
      dq = fminf(fmaxf(e, -1.f), 1.f) * inv_cnt;
    }
    *q = dq;
    *yp = hub;                    // the target is dead from here on: its slot carries the element's Huber value
  }
  __syncthreads();
  // per-head sums in a fixed order (one thread per head, graphs then channels): bit-stable loss, no float atomics
  if (tid < N) {
    float s = 0.f;
    for (int g = 0; g < ng; ++g) {
      const int r = SLOT ? tid * TGp + g : g * N + tid;
      for (int c = 0; c < P->CH; ++c) s += arena[tab[P->y_tab + c] * RP + r];
    }
    hl_s[tid] += s;
  }
}

__device__ __forceinline__ void bias_grad(const FusedOp& op, int op_idx, const float* arena, const int* tab, int RP,
                                          int my_bias, float& bacc) {
  if (my_bias >= 0 && (my_bias >> 16) == op_idx) {
    const float* dz = arena + tab[op.dz_tab + (my_bias & 0xffff)] * RP;
    float s = 0.f;
    for (int r = 0; r < RP; r += 4) {
      const float4 t = *reinterpret_cast<const float4*>(dz + r);
      s += (t.x + t.y) + (t.z + t.w);
    }
    bacc += s;
  }
}

// 4x4 weight-gradient block: a[i*4+j] += sum over rows r = r_begin, r_begin + r_step, ... (4 rows each)
__device__ __forceinline__ void wgrad_block(const FusedOp& op, const float* arena, const int* tab, int RP, int zero_row,
                                            int kb, int ob, int r_begin, int r_step, int r_end, float (&a)[16]) {
  const int OB = op.O >> 2;
  const float* xr[4];
  const float* dr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = kb * 4 + i;
    xr[i] = arena + (k < op.K ? tab[op.in_tab + k] : zero_row) * RP;
    dr[i] = arena + tab[op.dz_tab + ob + i * OB] * RP;      // strided columns: the lanes of a warp hit consecutive rows
  }
#pragma unroll 2
  for (int r = r_begin; r < r_end; r += r_step) {
    float4 xv[4], dv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(xr[i] + r);
      dv[i] = *reinterpret_cast<const float4*>(dr[i] + r);
    }
    // 16 independent accumulators per component: dependent FFMAs are 16 instructions apart
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i * 4 + j] = fmaf(xv[i].x, dv[j].x, a[i * 4 + j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i * 4 + j] = fmaf(xv[i].y, dv[j].y, a[i * 4 + j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i * 4 + j] = fmaf(xv[i].z, dv[j].z, a[i * 4 + j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a[i * 4 + j] = fmaf(xv[i].w, dv[j].w, a[i * 4 + j]);
  }
}

template <bool SLOT>
__device__ __forceinline__ void bwd_phase(const FusedOp& op, int op_idx, float* arena, const float* Ws, const int* tab, int RP,
                                          int zero_row, int tid, const int (&my_blk)[kFusedBlkPerThread],
                                          float (&wacc)[kFusedBlkPerThread][16], int my_bias, float& bacc, float* dWs,
                                          int* dx_ctr, int G, int TGp, float* part, bool first_tile) {
  // ---- weight gradient.  Large layers: 4x4 blocks owned by fixed threads, accumulated in registers across
  // tiles.  Small layers (op.nwt = RS > 0): every block is split over RS adjacent lanes by row groups, reduced
  // with shuffles, and the first lane adds the block into the CTA's shared accumulator (one owner per
  // address: deterministic) -- otherwise a 5..48-block layer would leave most of the CTA idle.
  int owners_begin = 0, owners_n = 0;
  if (SLOT) {
    // per-slot weights: block (slot, kb, ob) sums over the TGp rows of its slot and goes straight into this CTA's partial
    // row in global memory (one owner per address and tile: deterministic).  The CTA's first tile stores (nothing to
    // read, nothing to zero beforehand), later tiles load all 16 values, add, store: no dependent round trips.
    const int OB = op.O >> 2, KB = (op.K + 3) >> 2;
    const int per_slot = KB * OB;
    for (int blk = tid; blk < G * per_slot; blk += kFusedThreads) {
      const int slot = blk / per_slot, rem = blk - slot * per_slot;
      const int kb = rem / OB, ob = rem - kb * OB;
      float a[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = 0.f;
      wgrad_block(op, arena, tab, RP, zero_row, kb, ob, slot * TGp, 4, (slot + 1) * TGp, a);
      float* dst = part + op.w_off + slot * op.slot_w + (kb * 4) * op.O + ob;
      if (!first_tile) {
        float old[16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) old[i * 4 + j] = (kb * 4 + i < op.K) ? __ldcg(dst + i * op.O + j * OB) : 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += old[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (kb * 4 + i < op.K) {
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[i * op.O + j * OB] = a[i * 4 + j];
        }
      }
    }
    for (int idx = tid; idx < G * op.O; idx += kFusedThreads) {
      const int slot = idx / op.O, o = idx - slot * op.O;
      const float* dz = arena + tab[op.dz_tab + o] * RP + slot * TGp;
      float sb = 0.f;
      for (int r = 0; r < TGp; r += 4) {
        const float4 t = *reinterpret_cast<const float4*>(dz + r);
        sb += (t.x + t.y) + (t.z + t.w);
      }
      float* bd = part + op.b_off + slot * op.slot_b + o;
      *bd = first_tile ? sb : (__ldcg(bd) + sb);
    }
    if (first_tile) {                                        // parameter rows without an input (stage-0 neighbour rows): exact zeros
      const int Kp = op.slot_w / op.O, dead = (Kp - op.K) * op.O;
      for (int idx = tid; idx < G * dead; idx += kFusedThreads) {
        const int slot = idx / dead, r = idx - slot * dead;
        part[op.w_off + slot * op.slot_w + op.K * op.O + r] = 0.f;
      }
    }
  } else if (op.nwt > 0) {
    const int RS = op.nwt, OB = op.O >> 2;
    const int tasks = op.nblk * RS;
    const int blk = tid / RS, split = tid - blk * RS;
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.f;
    const int kb = blk / OB, ob = blk - kb * OB;
    if (tid < tasks) wgrad_block(op, arena, tab, RP, zero_row, kb, ob, split * 4, RS * 4, RP, a);
    for (int off = RS >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
    }
    if (tid < tasks && split == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = kb * 4 + i;
        if (k < op.K) {
#pragma unroll
          for (int j = 0; j < 4; ++j) dWs[op.wt0 + k * op.O + ob + j * OB] += a[i * 4 + j];
        }
      }
    }
    owners_begin = 0;
    owners_n = min(tasks, kFusedThreads);
  } else {
#pragma unroll
    for (int s = 0; s < kFusedBlkPerThread; ++s) {
      const int b = tid + s * kFusedThreads;
      if (b >= op.blk0 && b < op.blk0 + op.nblk) {
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = 0.f;
        wgrad_block(op, arena, tab, RP, zero_row, (my_blk[s] >> 8) & 0xff, my_blk[s] & 0xff, 0, 4, RP, a);
#pragma unroll
        for (int i = 0; i < 16; ++i) wacc[s][i] += a[i];
      }
    }
    owners_begin = op.blk0 % kFusedThreads;
    owners_n = min(op.nblk, kFusedThreads);
  }
  if (!SLOT) bias_grad(op, op_idx, arena, tab, RP, my_bias, bacc);
  // ---- data gradient of the requested input columns (4 rows x 4 columns per item)
  // Item shape of the data gradient.  On the register-owned weight-gradient path (nwt == 0) the block owners are busy with
  // one long block each, so the data gradient should be finished by the other threads in ONE round: 8 input columns x 4
  // rows per item when that many items fit the non-owners (half the items, 12 shared-memory vectors per 128 FMAs instead
  // of 8 per 64, packed FMAs on row pairs).  Every output is the same sum in the same order as in the 4-column form:
  // bit-identical results.  Traced on the configs[1] tile: `bwd mlp1` 19.2k -> 15.5k cycles; where every thread owns
  // weight-gradient work (the row-split small layers) the smaller 4-column items balance better and stay.
  const bool wide = !SLOT && op.nwt == 0 && op.n_dx > 0 && (op.n_dx & 7) == 0 &&
                    (RP / 4) * (op.n_dx / 8) <= kFusedThreads - min(op.nblk, kFusedThreads);
  if (wide) {
    const int row_groups = RP / 4;
    const int items = row_groups * (op.n_dx / 8);
    const int4* dz_rows4 = reinterpret_cast<const int4*>(tab + op.dz_tab);
    for (;;) {
      const int item = atomicAdd(dx_ctr, 1);
      if (item >= items) break;
      const int kg = item / row_groups, rg = item - kg * row_groups;
      const int r0 = rg * 4;
      const float* wr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wr[i] = Ws + op.w_off + tab[op.dxk_tab + kg * 8 + i] * op.O;
      unsigned long long acc[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = pk2(0.f, 0.f);
      const float* ab = arena + r0;
#pragma unroll 2
      for (int o = 0; o < op.O; o += 4) {
        const int4 zr = dz_rows4[o >> 2];
        const float4 d0 = *reinterpret_cast<const float4*>(ab + zr.x * RP), d1 = *reinterpret_cast<const float4*>(ab + zr.y * RP);
        const float4 d2 = *reinterpret_cast<const float4*>(ab + zr.z * RP), d3 = *reinterpret_cast<const float4*>(ab + zr.w * RP);
        const unsigned long long z[4][2] = {{pk2(d0.x, d0.y), pk2(d0.z, d0.w)}, {pk2(d1.x, d1.y), pk2(d1.z, d1.w)},
                                            {pk2(d2.x, d2.y), pk2(d2.z, d2.w)}, {pk2(d3.x, d3.y), pk2(d3.z, d3.w)}};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(wr[i] + o);
          const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int oo = 0; oo < 4; ++oo) {
            const unsigned long long w2 = pk2(ws[oo], ws[oo]);
            fma2(acc[i][0], z[oo][0], w2);
            fma2(acc[i][1], z[oo][1], w2);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 lo = unpk2(acc[i][0]), hi = unpk2(acc[i][1]);
        float4 v = make_float4(lo.x, lo.y, hi.x, hi.y);
        if (op.gate_tab >= 0) {
          const float4 g = *reinterpret_cast<const float4*>(arena + tab[op.gate_tab + kg * 8 + i] * RP + r0);
          v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f;
          v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
        }
        *reinterpret_cast<float4*>(arena + tab[op.dx_tab + kg * 8 + i] * RP + r0) = v;
      }
    }
  } else if (op.n_dx > 0) {
    const int row_groups = RP / 4;
    const int items = row_groups * (op.n_dx / 4);
    // Work queue: the threads without weight-gradient work start on the items at once, the block owners join as they
    // finish (static splits left the two groups 20-40 % apart: profiles trace of bwd mlp2 / mlp1).  Every item writes
    // its own outputs, so the result does not depend on who takes which item.
    (void)owners_begin; (void)owners_n;
    for (;;) {
      const int item = atomicAdd(dx_ctr, 1);
      if (item >= items) break;
      const int kg = item / row_groups, rg = item - kg * row_groups;
      const int r0 = rg * 4;
      const float* wr[4];
      const int wslot = SLOT ? (r0 / TGp) * op.slot_w : 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) wr[i] = Ws + op.w_off + wslot + tab[op.dxk_tab + kg * 4 + i] * op.O;
      float4 acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int4* dz_rows4 = reinterpret_cast<const int4*>(tab + op.dz_tab);
      const float* ab = arena + r0;
#pragma unroll 2
      for (int o = 0; o < op.O; o += 4) {
        const int4 zr = dz_rows4[o >> 2];
        float4 dz[4];
        dz[0] = *reinterpret_cast<const float4*>(ab + zr.x * RP); dz[1] = *reinterpret_cast<const float4*>(ab + zr.y * RP);
        dz[2] = *reinterpret_cast<const float4*>(ab + zr.z * RP); dz[3] = *reinterpret_cast<const float4*>(ab + zr.w * RP);
        float4 w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(wr[i] + o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i].x = fmaf(dz[0].x, w[i].x, acc[i].x); acc[i].y = fmaf(dz[0].y, w[i].x, acc[i].y);
          acc[i].z = fmaf(dz[0].z, w[i].x, acc[i].z); acc[i].w = fmaf(dz[0].w, w[i].x, acc[i].w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i].x = fmaf(dz[1].x, w[i].y, acc[i].x); acc[i].y = fmaf(dz[1].y, w[i].y, acc[i].y);
          acc[i].z = fmaf(dz[1].z, w[i].y, acc[i].z); acc[i].w = fmaf(dz[1].w, w[i].y, acc[i].w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i].x = fmaf(dz[2].x, w[i].z, acc[i].x); acc[i].y = fmaf(dz[2].y, w[i].z, acc[i].y);
          acc[i].z = fmaf(dz[2].z, w[i].z, acc[i].z); acc[i].w = fmaf(dz[2].w, w[i].z, acc[i].w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i].x = fmaf(dz[3].x, w[i].w, acc[i].x); acc[i].y = fmaf(dz[3].y, w[i].w, acc[i].y);
          acc[i].z = fmaf(dz[3].z, w[i].w, acc[i].z); acc[i].w = fmaf(dz[3].w, w[i].w, acc[i].w);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = acc[i];
        if (op.gate_tab >= 0) {
          const float4 g = *reinterpret_cast<const float4*>(arena + tab[op.gate_tab + kg * 4 + i] * RP + r0);
          v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f;
          v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
        }
        *reinterpret_cast<float4*>(arena + tab[op.dx_tab + kg * 4 + i] * RP + r0) = v;
      }
    }
  }
}


// ---------------------------------------------------------------------------
// Tensor-core form of the backward contractions of the shared-weight kernel (template parameter MMA of fused_brain_kernel).
//
// The tile's operands sit in the feature-major arena (row index fastest) and the weights in shared memory as W[k][o];
// tcgen05 cannot read the weight gradient's operands in that form for fp32-grade products (no-swizzle MN-major TF32
// operands read as zeros, scratch/tc_probe_mn.cu) and hi/lo copies of a whole tile do not fit shared memory, so these
// phases use the register-operand tensor-core instruction instead: mma.sync m16n8k8 TF32 (SASS HMMA.1688.F32.TF32; measured
// on B200, scratch/hmma_probe.cu: 23 cycles dependent, one per 2.17 cycles per SM) with every product done in three passes
// (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi; hi = the upper 19 bits, lo = the exact remainder, which the tensor core reads through
// its upper 19 bits: the dropped lo*lo term and the truncation of the two lo operands are each <= 2^-20 of |a||b| -- emulated
// in tests/test_properties.py -- i.e. fp32-grade dot products; fp32 accumulation).  A fragment element is ONE
// scalar shared-memory load whatever the orientation of the operand, so both gradients of a layer are the same routine:
//   data grad   dx[kq][r] = gate * sum_o W[kq][o] dz[o][r]      M = kq, N = r (tile rows),   reduction over o
//   weight grad dW[k][o]  = sum_r in[k][r] dz[o][r]             M = k,  N = o,               reduction over r
// A warp task is MT x NT blocks of 16 x 8 outputs: every loaded value is split once and used by MT (or NT) blocks, which
// is what keeps the split arithmetic (5 integer/float instructions per value) below the tensor pipe's time.  Every output
// element has one owner and a fixed summation order (deterministic).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
#if V2V_TF32_SPLIT_RN
  // round-to-nearest (ties away) on the bit pattern: finite inputs only, which the arena guarantees
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  const float r = x - __uint_as_float(hi);
  lo = (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
#else
  // truncating split, two instructions per value: hi = the upper 19 bits, lo = the exact remainder (13 bits), handed to
  // the tensor core as it is (the TF32 datapath reads the upper 19 bits of a 32-bit operand)
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
#endif
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int warp_next(int* ctr, int lane) {
  int v = 0;
  if (lane == 0) v = atomicAdd(ctr, 1);
  return __shfl_sync(0xffffffffu, v, 0);
}

// One reduction step of a 16 x (8 NT) task: a = the 4 A-fragment values, b[j] = the 2 B-fragment values of column block j.
// The three passes run pass-major over the NT blocks (volatile keeps that order): consecutive tensor-core instructions are
// independent, the next pass on the same accumulator is NT instructions away (HMMA latency 23 cycles, issue one per ~9).
__device__ __forceinline__ void mma_tf32_ordered(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int NT>
__device__ __forceinline__ void mma_step(float (&c)[NT][4], const float (&a)[4], const float (&b)[NT][2]) {
  uint32_t ah[4], al[4], bh[NT][2], bl[NT][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) tf32_split(a[q], ah[q], al[q]);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    tf32_split(b[j][0], bh[j][0], bl[j][0]);
    tf32_split(b[j][1], bh[j][1], bl[j][1]);
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) mma_tf32_ordered(c[j], al, bh[j][0], bh[j][1]);      // small terms first
#pragma unroll
  for (int j = 0; j < NT; ++j) mma_tf32_ordered(c[j], ah, bl[j][0], bl[j][1]);
#pragma unroll
  for (int j = 0; j < NT; ++j) mma_tf32_ordered(c[j], ah, bh[j][0], bh[j][1]);
}

// data gradient: 16 requested input columns (row block mt) x NT column blocks of 8 tile rows from nt0
template <int NT>
__device__ __forceinline__ void mma_dgrad_task(const FusedOp& op, float* arena, const float* Ws, const int* tab, int RP,
                                               int zero_row, int mt, int nt0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int O = op.O, n_dx = op.n_dx;
  const int m_lo = mt * 16 + g, m_hi = m_lo + 8;
  const bool v_lo = m_lo < n_dx, v_hi = m_hi < n_dx;
  const float* w_lo = Ws + op.w_off + (v_lo ? tab[op.dxk_tab + m_lo] : 0) * O + t;
  const float* w_hi = Ws + op.w_off + (v_hi ? tab[op.dxk_tab + m_hi] : 0) * O + t;
  const int* dz_rows = tab + op.dz_tab + t;
  float c[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  const float* bcol = arena + nt0 * 8 + g;
  // RP is a multiple of 4, not of 8: in the tile's last column block the lanes of columns >= RP read column RP - 1 instead
  // (their results are never stored; reading past the row would touch a row another warp may be writing)
  const int last_off = min((nt0 + NT - 1) * 8 + g, RP - 1) - (nt0 * 8 + g);
  // operands of step s + 1 are loaded before the tensor-core instructions of step s are issued (register double buffer)
  auto load = [&](int o0, float (&a)[4], float (&b)[NT][2]) {
    const bool pa = o0 + t < O, pb = o0 + t + 4 < O;
    a[0] = (pa && v_lo) ? w_lo[o0] : 0.f; a[1] = (pa && v_hi) ? w_hi[o0] : 0.f;
    a[2] = (pb && v_lo) ? w_lo[o0 + 4] : 0.f; a[3] = (pb && v_hi) ? w_hi[o0 + 4] : 0.f;
    const float* pa_ = bcol + (pa ? dz_rows[o0] : zero_row) * RP;
    const float* pb_ = bcol + (pb ? dz_rows[o0 + 4] : zero_row) * RP;
#pragma unroll
    for (int j = 0; j < NT - 1; ++j) { b[j][0] = pa_[j * 8]; b[j][1] = pb_[j * 8]; }
    b[NT - 1][0] = pa_[last_off]; b[NT - 1][1] = pb_[last_off];
  };
  float a0[4], b0[NT][2];
  load(0, a0, b0);
#pragma unroll 2
  for (int o0 = 8; o0 < O; o0 += 8) {
    float a1[4], b1[NT][2];
    load(o0, a1, b1);
    mma_step<NT>(c, a0, b0);
#pragma unroll
    for (int q = 0; q < 4; ++q) a0[q] = a1[q];
#pragma unroll
    for (int j = 0; j < NT; ++j) { b0[j][0] = b1[j][0]; b0[j][1] = b1[j][1]; }
  }
  mma_step<NT>(c, a0, b0);
  const bool gated = op.gate_tab >= 0;
  float* out_lo = arena + (v_lo ? tab[op.dx_tab + m_lo] : zero_row) * RP;
  float* out_hi = arena + (v_hi ? tab[op.dx_tab + m_hi] : zero_row) * RP;
  const float* g_lo = arena + ((gated && v_lo) ? tab[op.gate_tab + m_lo] : zero_row) * RP;
  const float* g_hi = arena + ((gated && v_hi) ? tab[op.gate_tab + m_hi] : zero_row) * RP;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int r = (nt0 + j) * 8 + 2 * t;
    if (r < RP) {
      float2 lo = make_float2(c[j][0], c[j][1]), hi = make_float2(c[j][2], c[j][3]);
      if (gated) {
        const float2 gl = *reinterpret_cast<const float2*>(g_lo + r), gh = *reinterpret_cast<const float2*>(g_hi + r);
        lo.x = gl.x > 0.f ? lo.x : 0.f; lo.y = gl.y > 0.f ? lo.y : 0.f;
        hi.x = gh.x > 0.f ? hi.x : 0.f; hi.y = gh.y > 0.f ? hi.y : 0.f;
      }
      if (v_lo) *reinterpret_cast<float2*>(out_lo + r) = lo;
      if (v_hi) *reinterpret_cast<float2*>(out_hi + r) = hi;
    }
  }
}

// weight gradient: 16 input features (row block mt) x NT column blocks of 8 output features from nt0, summed over the
// tile's RP rows; the block goes straight into this CTA's partial row (stored on the CTA's first tile, load-add-store
// afterwards)
template <int NT>
__device__ __forceinline__ void mma_wgrad_task(const FusedOp& op, const float* arena, const int* tab, int RP, int zero_row,
                                               int mt, int nt0, int lane, float* part, bool first_tile) {
  const int g = lane >> 2, t = lane & 3;
  const int K = op.K, O = op.O;
  const int k_lo = mt * 16 + g, k_hi = k_lo + 8;
  const float* x_lo = arena + (k_lo < K ? tab[op.in_tab + k_lo] : zero_row) * RP + t;
  const float* x_hi = arena + (k_hi < K ? tab[op.in_tab + k_hi] : zero_row) * RP + t;
  const float* dz[NT];
  float c[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int o = (nt0 + j) * 8 + g;
    dz[j] = arena + (o < O ? tab[op.dz_tab + o] : zero_row) * RP + t;
    c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
  }
  // the CTA's later tiles add to what its earlier tiles left in the partial row: fetch those values now, use them at the end
  float2 old_lo[NT], old_hi[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int o = (nt0 + j) * 8 + 2 * t;
    old_lo[j] = old_hi[j] = make_float2(0.f, 0.f);
    if (!first_tile && o < O) {
      if (k_lo < K) old_lo[j] = __ldcg(reinterpret_cast<const float2*>(part + op.w_off + k_lo * O + o));
      if (k_hi < K) old_hi[j] = __ldcg(reinterpret_cast<const float2*>(part + op.w_off + k_hi * O + o));
    }
  }
  // operands of step s + 1 are loaded before the tensor-core instructions of step s are issued (register double buffer);
  // RP is a multiple of 4: the last step may hold 4 rows
  auto load = [&](int r0, float (&a)[4], float (&b)[NT][2]) {
    const bool full = r0 + 8 <= RP;
    a[0] = x_lo[r0]; a[1] = x_hi[r0];
    a[2] = full ? x_lo[r0 + 4] : 0.f; a[3] = full ? x_hi[r0 + 4] : 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) { b[j][0] = dz[j][r0]; b[j][1] = full ? dz[j][r0 + 4] : 0.f; }
  };
  float a0[4], b0[NT][2];
  load(0, a0, b0);
#pragma unroll 2
  for (int r0 = 8; r0 < RP; r0 += 8) {
    float a1[4], b1[NT][2];
    load(r0, a1, b1);
    mma_step<NT>(c, a0, b0);
#pragma unroll
    for (int q = 0; q < 4; ++q) a0[q] = a1[q];
#pragma unroll
    for (int j = 0; j < NT; ++j) { b0[j][0] = b1[j][0]; b0[j][1] = b1[j][1]; }
  }
  mma_step<NT>(c, a0, b0);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int o = (nt0 + j) * 8 + 2 * t;
    if (o < O) {
      if (k_lo < K) {
        float2* d = reinterpret_cast<float2*>(part + op.w_off + k_lo * O + o);
        *d = make_float2(c[j][0] + old_lo[j].x, c[j][1] + old_lo[j].y);
      }
      if (k_hi < K) {
        float2* d = reinterpret_cast<float2*>(part + op.w_off + k_hi * O + o);
        *d = make_float2(c[j][2] + old_hi[j].x, c[j][3] + old_hi[j].y);
      }
    }
  }
}

// backward phase of one layer on the tensor cores: weight-gradient tasks (the long ones) first, then the data-gradient
// tasks, handed to the warps through the phase's work counter.  A task = one 16-row block x up to kMmaNT column blocks
// (balanced group sizes; the exact count is a template parameter, so no tensor-core instruction is predicated).
#ifndef V2V_MMA_NT
#define V2V_MMA_NT 6
#endif
constexpr int kMmaNT = V2V_MMA_NT;
#define V2V_MMA_DISPATCH(n, CALL)                 \
  switch (n) {                                    \
    case 1: { constexpr int NT_ = 1; CALL; break; } \
    case 2: { constexpr int NT_ = kMmaNT >= 2 ? 2 : kMmaNT; CALL; break; } \
    case 3: { constexpr int NT_ = kMmaNT >= 3 ? 3 : kMmaNT; CALL; break; } \
    case 4: { constexpr int NT_ = kMmaNT >= 4 ? 4 : kMmaNT; CALL; break; } \
    case 5: { constexpr int NT_ = kMmaNT >= 5 ? 5 : kMmaNT; CALL; break; } \
    case 6: { constexpr int NT_ = kMmaNT >= 6 ? 6 : kMmaNT; CALL; break; } \
    case 7: { constexpr int NT_ = kMmaNT >= 7 ? 7 : kMmaNT; CALL; break; } \
    default: { constexpr int NT_ = kMmaNT; CALL; break; } \
  }
__device__ __forceinline__ void mma_bwd_phase(const FusedOp& op, int op_idx, float* arena, const float* Ws, const int* tab,
                                              int RP, int zero_row, int tid, int my_bias, float& bacc, int* ctr, float* part,
                                              bool first_tile) {
  bias_grad(op, op_idx, arena, tab, RP, my_bias, bacc);
  __syncwarp();
  const int lane = tid & 31;
  // weight gradient: K / 16 row blocks x O / 8 column blocks; data gradient: n_dx / 16 row blocks x RP / 8 column blocks
  const int wM = (op.K + 15) >> 4, wN = (op.O + 7) >> 3;
  const int wNg = (wN + kMmaNT - 1) / kMmaNT, wNper = (wN + wNg - 1) / wNg;
  const int n_w = wM * wNg;
  const int dM = (op.n_dx + 15) >> 4, dN = (RP + 7) >> 3;
  const int dNg = (dN + kMmaNT - 1) / kMmaNT, dNper = (dN + dNg - 1) / dNg;
  const int n_d = dM * dNg;
  for (;;) {
    int task = warp_next(ctr, lane);
    if (task >= n_w + n_d) break;
#ifdef V2V_MMA_DGRAD_FIRST
    task = task < n_d ? task + n_w : task - n_d;
#endif
    if (task < n_w) {
      const int mt = task / wNg, nt0 = (task - mt * wNg) * wNper;
      V2V_MMA_DISPATCH(min(wNper, wN - nt0), mma_wgrad_task<NT_>(op, arena, tab, RP, zero_row, mt, nt0, lane, part, first_tile));
    } else {
      task -= n_w;
      const int mt = task / dNg, nt0 = (task - mt * dNg) * dNper;
      V2V_MMA_DISPATCH(min(dNper, dN - nt0), mma_dgrad_task<NT_>(op, arena, Ws, tab, RP, zero_row, mt, nt0, lane));
    }
  }
}
#undef V2V_MMA_DISPATCH

// Shared weights: the tile's node / edge / target rows -> feature rows of the arena.  The three spans are contiguous in
// global memory; with `bulk` one thread fetches them by bulk-async copies into arena rows that are dead at tile start (all
// bytes in flight at once: one memory latency instead of one per round of a load loop) and the CTA transposes from shared
// memory.  Kept out of line: inlined, its temporaries cost the phase loops ten registers.
__device__ __noinline__ void load_tile_inputs(const FusedProgram* P, float* arena, const int* tab, const float* node,
                                              const float* edge, const float* y, int g0, int valid_rows, bool bulk,
                                              uint32_t parity, uint64_t* bar, int tid) {
  const int N = P->N, RP = P->RP, Dn = P->Dn, De = P->De, CH = P->CH;
  const float* nsrc = node + (size_t)g0 * N * Dn;
  const float* esrc = edge + (size_t)g0 * N * De;
  const float* ysrc = y ? y + (size_t)g0 * N * CH : nullptr;
  if (bulk) {
    float* stage = arena + (size_t)P->stage_row0 * RP;
    if (tid == 0) {
      fence_async_smem();                  // the rows were written by the previous tile's phases (generic proxy)
      const uint32_t nb = (uint32_t)valid_rows * Dn * 4u, eb = (uint32_t)valid_rows * De * 4u;
      const uint32_t yb = y ? (uint32_t)valid_rows * CH * 4u : 0u;
      mbar_arrive_expect_tx(bar, nb + eb + yb);
      bulk_g2s(stage, nsrc, nb, bar);
      bulk_g2s(stage + RP * Dn, esrc, eb, bar);
      if (y) bulk_g2s(stage + RP * (Dn + De), ysrc, yb, bar);
    }
    mbar_wait(bar, parity);                // one completion per tile
    nsrc = stage; esrc = stage + RP * Dn; ysrc = stage + RP * (Dn + De);
  }
  for (int idx = tid; idx < RP * Dn; idx += kFusedThreads) {
    const int r = idx / Dn, f = idx - r * Dn;
    arena[(P->x0_row0 + f) * RP + r] = (r < valid_rows) ? nsrc[idx] : 0.f;
  }
  for (int idx = tid; idx < RP * De; idx += kFusedThreads) {
    const int r = idx / De, f = idx - r * De;
    arena[(P->x0_row0 + Dn + f) * RP + r] = (r < valid_rows) ? esrc[idx] : 0.f;
  }
  if (y) {
    for (int idx = tid; idx < RP * CH; idx += kFusedThreads) {
      const int r = idx / CH, c = idx - r * CH;
      arena[tab[P->y_tab + c] * RP + r] = (r < valid_rows) ? ysrc[idx] : 0.f;
    }
  }
}

__device__ long long* g_fused_trace = nullptr;       // optional per-phase clock trace of CTA 0 (profiling aid)

// MMA: 0 = every contraction on the FP32 pipe, 1 = backward contractions (weight and data gradients) on the tensor cores
// (shared weights only)
template <int NMAX, bool SLOT, int MMA>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_brain_kernel(const FusedProgram* __restrict__ P, const float* __restrict__ params, const float* __restrict__ node,
                   const float* __restrict__ edge, const uint32_t* __restrict__ in_mask,
                   const uint32_t* __restrict__ out_mask, const float* __restrict__ y, float* __restrict__ q_out, float* __restrict__ partial, float* __restrict__ head_loss, int B,
                   float inv_cnt) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int dx_ctr[2];                            // data-gradient work queues of two consecutive backward phases
  const int tid = threadIdx.x;
  if (tid < 2) dx_ctr[tid] = 0;
  const int N = P->N, TG = P->TG, RP = P->RP, CH = P->CH, Dn = P->Dn, De = P->De;
  const int G = P->G, TGp = P->TGp, row_g = P->row_g, row_n = P->row_n;
  const int n_params = P->n_params, n_ops = P->n_ops, n_tab = P->n_tab, train = P->train;
  // shared weights live in shared memory; per-slot weights (G x as many: 151 KB at the reference's N = 4) stay in global
  // memory / L2 and are read through the read-only path (a warp reads one broadcast vector per k)
  const int np_pad = SLOT ? 0 : ((n_params + 3) & ~3);
  float* Ws_s = smem;
  const float* Ws = SLOT ? params : Ws_s;
  int* tab = reinterpret_cast<int*>(Ws_s + np_pad);
  FusedOp* ops = reinterpret_cast<FusedOp*>(tab + ((n_tab + 3) & ~3));
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(ops + n_ops);
  float* hl_s = reinterpret_cast<float*>(mask_s + ((2 * TG * N + 3) & ~3));
  float* dWs = hl_s + 32;                                   // shared weight-gradient accumulator of the small layers
  float* wbuf = dWs + P->n_small;                           // per-slot weights: two layer buffers filled by bulk-async copies
  const int wstage = SLOT ? P->wstage_floats : 0;
  float* arena = wbuf + 2 * wstage;
  __shared__ __align__(8) uint64_t wbar[2];
  if (tid == 0) {
    mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1);
    fence_mbar_init();
  }
  // Shared weights: a tile's node / edge / target rows are three contiguous spans of global memory.  One thread fetches them
  // with bulk-async copies into arena rows that are dead at tile start (all bytes in flight at once: one memory latency
  // instead of one per round of a load loop), then the CTA transposes them into feature rows from shared memory.
  const bool bulk_in = !SLOT && P->stage_row0 >= 0 && (N * Dn) % 4 == 0 && (N * De) % 4 == 0 && (!train || (N * CH) % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(node) | reinterpret_cast<uintptr_t>(edge) |
                         (train ? reinterpret_cast<uintptr_t>(y) : (uintptr_t)0)) & 15u) == 0;
  float* part = (train && partial) ? partial + (size_t)blockIdx.x * (n_params + kFusedPartialTail) : nullptr;

  // programmatic dependent launch: everything up to griddepcontrol.wait overlaps the tail of the preceding kernel
  // (normally the previous step's reduce/Adam); it only reads the immutable program and writes shared memory
  for (int i = tid; i < n_tab; i += kFusedThreads) tab[i] = P->tab[i];
  {
    const int words = n_ops * (int)(sizeof(FusedOp) / 4);
    const int* src = reinterpret_cast<const int*>(P->ops);
    int* dst = reinterpret_cast<int*>(ops);
    for (int i = tid; i < words; i += kFusedThreads) dst[i] = src[i];
  }
  if (tid < 32) hl_s[tid] = 0.f;
  for (int i = tid; i < P->n_small; i += kFusedThreads) dWs[i] = 0.f;
  for (int i = tid; i < P->n_rows * RP / 4; i += kFusedThreads)      // whole arena: finite everywhere, zero row included
    reinterpret_cast<float4*>(arena)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!SLOT) {
    for (int i = tid; i < n_params / 4; i += kFusedThreads)
      reinterpret_cast<float4*>(Ws_s)[i] = reinterpret_cast<const float4*>(params)[i];
    for (int i = (n_params & ~3) + tid; i < n_params; i += kFusedThreads) Ws_s[i] = params[i];
  }

  // the whole grid is resident (one CTA per SM): let the dependent reduce/Adam kernel be scheduled now, its CTAs
  // park in griddepcontrol.wait until this grid has completed and flushed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = tid >> 5, lane = tid & 31;
  int my_blk[kFusedBlkPerThread];
  float wacc[kFusedBlkPerThread][16];                  // two 4x4 weight-gradient blocks per thread (large layers)
#pragma unroll
  for (int s = 0; s < kFusedBlkPerThread; ++s) {
    const int b = tid + s * kFusedThreads;
    my_blk[s] = (!SLOT && MMA == 0 && train && b < P->n_blocks) ? P->blk_info[b] : 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) wacc[s][i] = 0.f;
  }
  const int bias_slot = kFusedThreads - 1 - tid;
  const int my_bias = (!SLOT && train && bias_slot < P->n_bias) ? P->bias_info[bias_slot] : -1;
  float bacc = 0.f;

  const int n_tiles = (B + TG - 1) / TG;
  // per-slot weights: the layers' [W | bias] spans stream through two shared-memory buffers, one layer ahead of its use
  int wcur = 0;
  uint32_t wpar[2] = {0u, 0u};
  auto next_weight_op = [&](int from) {
    for (int o2 = from; o2 < n_ops; ++o2)
      if (ops[o2].type == FOP_GEMM || ops[o2].type == FOP_BWD) return o2;
    return -1;
  };
  auto issue_weights = [&](int o2, int buf) {               // one thread
    const uint32_t bytes = (uint32_t)ops[o2].w_span * 4u;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(params + ops[o2].w_off);
    uint8_t* dstb = reinterpret_cast<uint8_t*>(wbuf + (size_t)buf * wstage);
    mbar_arrive_expect_tx(&wbar[buf], bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(dstb + off, src + off, min(32768u, bytes - off), &wbar[buf]);
  };
  __syncthreads();                           // ops / barriers visible before the first prefetch
  if (SLOT && wstage > 0 && tid == 0 && blockIdx.x < n_tiles) {
    const int o2 = next_weight_op(0);
    if (o2 >= 0) issue_weights(o2, 0);
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int g0 = tile * TG;
    const int ng = min(TG, B - g0);
    const int valid_rows = ng * N;
    __syncthreads();                         // previous tile fully consumed (and the prologue stores)
    // adjacency bit masks: plain loads, issued first so that they are in flight together with the bulk copies below
    for (int idx = tid; idx < TG * N; idx += kFusedThreads) {
      mask_s[idx] = (idx < valid_rows) ? in_mask[(size_t)g0 * N + idx] : 0u;
      if (train) mask_s[TG * N + idx] = (idx < valid_rows) ? out_mask[(size_t)g0 * N + idx] : 0u;
    }
    if (!SLOT) {
      load_tile_inputs(P, arena, tab, node, edge, train ? y : nullptr, g0, valid_rows, bulk_in,
                       (uint32_t)((tile - (int)blockIdx.x) / (int)gridDim.x) & 1u, &wbar[0], tid);
    } else {                                 // slot-major rows: node (g, n) -> row n * TGp + g
      const float* nsrc = node + (size_t)g0 * N * Dn;
      for (int idx = tid; idx < TG * N * Dn; idx += kFusedThreads) {
        const int gn = idx / Dn, f = idx - gn * Dn;
        const int g = gn / N, n = gn - g * N;
        arena[(P->x0_row0 + f) * RP + g * row_g + n * row_n] = (g < ng) ? nsrc[idx] : 0.f;
      }
      const float* esrc = edge + (size_t)g0 * N * De;
      for (int idx = tid; idx < TG * N * De; idx += kFusedThreads) {
        const int gn = idx / De, f = idx - gn * De;
        const int g = gn / N, n = gn - g * N;
        arena[(P->x0_row0 + Dn + f) * RP + g * row_g + n * row_n] = (g < ng) ? esrc[idx] : 0.f;
      }
      if (train) {
        const float* ysrc = y + (size_t)g0 * N * CH;
        for (int idx = tid; idx < TG * N * CH; idx += kFusedThreads) {
          const int gn = idx / CH, c = idx - gn * CH;
          const int g = gn / N, n = gn - g * N;
          arena[tab[P->y_tab + c] * RP + g * row_g + n * row_n] = (g < ng) ? ysrc[idx] : 0.f;
        }
      }
    }
    __syncthreads();
    // profiling aid: lane 0 of every warp of CTA 0 records [phase][warp] = {work done, barrier released}
    long long* trace = (blockIdx.x == 0 && lane == 0 && tile == 0) ? g_fused_trace : nullptr;
    if (trace) trace[warp * 2 + 1] = clock64();
    for (int oi = 0; oi < n_ops; ++oi) {
      const FusedOp& op = ops[oi];
      if (tid == 0) dx_ctr[(oi + 1) & 1] = 0;              // the NEXT phase's work queue (idle: its last user ended a barrier ago)
      const float* Wop = Ws;
      if (SLOT && wstage > 0 && (op.type == FOP_GEMM || op.type == FOP_BWD)) {
        mbar_wait(&wbar[wcur], wpar[wcur]);                // this layer's span has landed
        wpar[wcur] ^= 1u;
        if (tid == 0) {                                    // the other buffer's last reader ended a barrier ago: refill it
          int nx = next_weight_op(oi + 1);
          if (nx < 0 && tile + (int)gridDim.x < n_tiles) nx = next_weight_op(0);
          if (nx >= 0) issue_weights(nx, wcur ^ 1);
        }
        Wop = wbuf + (size_t)wcur * wstage - op.w_off;      // so that Wop + op.w_off (+ slot strides) and Wop + op.b_off index the span
        wcur ^= 1;
      }
      switch (op.type) {
        case FOP_GEMM: {
          const int rg4 = RP / 4;
          if (op.O % 8 == 0 && rg4 * (op.O / 8) >= (kFusedThreads * 3) / 4) gemm_phase<4, 8, SLOT>(op, arena, Wop, tab, RP, TGp, tid);
          else if (SLOT || rg4 * (op.O / 4) >= kFusedThreads / 2) gemm_phase<4, 4, SLOT>(op, arena, Wop, tab, RP, TGp, tid);
          else gemm_phase<2, 4, SLOT>(op, arena, Wop, tab, RP, TGp, tid);
          break;
        }
        case FOP_AGG: agg_phase<NMAX, SLOT>(op, arena, tab, mask_s, RP, N, TG, TGp, tid); break;
        case FOP_LOSS: loss_phase<SLOT>(P, arena, tab, hl_s, RP, ng, inv_cnt, tid); break;
        case FOP_BWD:
          if (MMA != 0) mma_bwd_phase(op, oi, arena, Wop, tab, RP, P->zero_row, tid, my_bias, bacc, &dx_ctr[oi & 1], part, tile == (int)blockIdx.x);
          else bwd_phase<SLOT>(op, oi, arena, Wop, tab, RP, P->zero_row, tid, my_blk, wacc, my_bias, bacc, dWs, &dx_ctr[oi & 1], G, TGp, part, tile == (int)blockIdx.x); break;
        default: break;
      }
      if (trace) trace[((oi + 1) * kFusedWarps + warp) * 2] = clock64();
      __syncthreads();
      if (trace) trace[((oi + 1) * kFusedWarps + warp) * 2 + 1] = clock64();
    }
    if (!train) {
      float* qdst = q_out + (size_t)g0 * N * CH;
      for (int idx = tid; idx < valid_rows * CH; idx += kFusedThreads) {
        const int gn = idx / CH, c = idx - gn * CH;
        const int g = gn / N, n = gn - g * N;
        qdst[idx] = arena[tab[P->q_tab + c] * RP + g * row_g + n * row_n];
      }
    }
  }

  if (train) {
    float* dst = partial + (size_t)blockIdx.x * (n_params + kFusedPartialTail);
    if (!SLOT && MMA != 0) {            // the weight gradients went to the partial row phase by phase
      if (my_bias >= 0) dst[ops[my_bias >> 16].b_off + (my_bias & 0xffff)] = bacc;
    } else if (!SLOT) {
#pragma unroll
      for (int s = 0; s < kFusedBlkPerThread; ++s) {
        const int b = tid + s * kFusedThreads;
        if (b < P->n_blocks) {
          const FusedOp& op = ops[my_blk[s] >> 16];
          const int kb = (my_blk[s] >> 8) & 0xff, ob = my_blk[s] & 0xff;
          const int OB = op.O >> 2;
          for (int i = 0; i < 4; ++i) {
            const int k = kb * 4 + i;
            if (k < op.K) {
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[op.w_off + k * op.O + ob + j * OB] = wacc[s][i * 4 + j];
            }
          }
        }
      }
      if (my_bias >= 0) dst[ops[my_bias >> 16].b_off + (my_bias & 0xffff)] = bacc;
      __syncthreads();
      for (int oi = 0; oi < n_ops; ++oi) {
        const FusedOp& op = ops[oi];
        if (op.type == FOP_BWD && op.nwt > 0)
          for (int i = tid; i < op.K * op.O; i += kFusedThreads) dst[op.w_off + i] = dWs[op.wt0 + i];
      }
    }
    __syncthreads();
    if (tid < kFusedPartialTail) dst[n_params + tid] = (tid < N) ? hl_s[tid] * inv_cnt : 0.f;   // per-head Huber sums
  }
}

// grad[i] = sum_c partial[c][i] (+ Keras-Adam).  The sum over the n_cta per-CTA partials is latency-bound (every
// addend is an L2 round trip), so a CTA covers only kRedCols columns and splits the partials over kRedSlices warps:
// each thread has <= ceil(n_cta / kRedSlices) independent loads in flight, the slices meet in shared memory and are
// added in slice order (deterministic).  Programmatic dependent launch: the prologue overlaps the producer's tail.
constexpr int kRedCols = 32, kRedSlices = 16;
__global__ void __launch_bounds__(kRedCols * kRedSlices)
reduce_adam_kernel(const float* __restrict__ partial, int n_cta, long stride, float* __restrict__ grad,
                   float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, long n, int n_tail,
                   float* __restrict__ tail_out, int do_adam, float lr_t, float b1, float b2, float eps, float gscale) {
  __shared__ float red[kRedSlices][kRedCols];
  const int col = threadIdx.x & (kRedCols - 1), slice = threadIdx.x / kRedCols;
  const long i = (long)blockIdx.x * kRedCols + col;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // every CTA is running: the next step may queue up
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  if (i < n + n_tail) {
    const float* src = partial + i;
    int c = slice;
    for (; c + 3 * kRedSlices < n_cta; c += 4 * kRedSlices) {
      const float a0 = __ldcg(src + (long)c * stride), a1 = __ldcg(src + (long)(c + kRedSlices) * stride);
      const float a2 = __ldcg(src + (long)(c + 2 * kRedSlices) * stride), a3 = __ldcg(src + (long)(c + 3 * kRedSlices) * stride);
      g0 += a0; g1 += a1; g2 += a2; g3 += a3;
    }
    for (; c < n_cta; c += kRedSlices) g0 += __ldcg(src + (long)c * stride);
  }
  red[slice][col] = (g0 + g1) + (g2 + g3);
  __syncthreads();
  if (slice == 0 && i < n + n_tail) {
    float g = 0.f;
#pragma unroll
    for (int s = 0; s < kRedSlices; ++s) g += red[s][col];
    if (i >= n) {
      if (tail_out) tail_out[i - n] = g;
      return;
    }
    grad[i] = g;
    if (do_adam) {
      const float gi = g * gscale;
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
  }
}

// ---------------------------------------------------------------------------
// host side: program builder
// ---------------------------------------------------------------------------
namespace {

struct Builder {
  FusedProgram* P;
  std::vector<int> free_rows;
  int next_row = 0;
  bool overflow = false;

  int tab_put(const std::vector<int>& rows) {          // 16-B aligned, padded with the zero row
    const int off = P->n_tab;
    const int n = ((int)rows.size() + 3) & ~3;
    if (off + n > kFusedMaxTab) { overflow = true; return 0; }
    for (int r : rows) P->tab[P->n_tab++] = r;
    while (P->n_tab < off + n) P->tab[P->n_tab++] = P->zero_row;
    return off;
  }
  std::vector<int> fresh(int n) {
    std::vector<int> v(n);
    for (int i = 0; i < n; ++i) v[i] = next_row++;
    return v;
  }
  std::vector<int> take(int n) {                 // re-use dead rows first
    std::vector<int> v;
    while ((int)v.size() < n && !free_rows.empty()) { v.push_back(free_rows.back()); free_rows.pop_back(); }
    while ((int)v.size() < n) v.push_back(next_row++);
    return v;
  }
  void release(const std::vector<int>& rows) { free_rows.insert(free_rows.end(), rows.begin(), rows.end()); }
  FusedOp* add_op() {
    if (P->n_ops >= kFusedMaxOps) { overflow = true; return &P->ops[0]; }
    FusedOp* op = &P->ops[P->n_ops++];
    *op = FusedOp{};
    op->add_tab = op->gate_tab = -1;
    return op;
  }
};

std::vector<int> cat(std::initializer_list<std::vector<int>> parts) {
  std::vector<int> v;
  for (auto& p : parts) v.insert(v.end(), p.begin(), p.end());
  return v;
}
std::vector<int> iota(int first, int n) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = first + i;
  return v;
}

}  // namespace

int fused_build_program(const FusedShape& s, int TG, int train, FusedProgram* out) {
  V2V_REQUIRE(s.N <= 32 && s.F % 4 == 0 && s.CH % 4 == 0 && s.H1 % 4 == 0 && s.H2 % 4 == 0 && s.H3 % 4 == 0,
              "fused path: unsupported dimensions");
  V2V_REQUIRE(s.G == 1 || s.G == s.N, "fused path: weight groups must be 1 (shared) or N (per slot)");
  const bool slot = s.G > 1;
  V2V_REQUIRE(s.n_layers == s.S + 4 && s.S >= 1, "fused path: unexpected layer count");
  for (int l = 0; l < s.n_layers; ++l)
    V2V_REQUIRE(s.w_off[l] % 4 == 0 && s.b_off[l] % 4 == 0 && s.layer_K[l] <= 255 && s.layer_O[l] <= 1020,
                "fused path: unaligned parameter layout");
  FusedProgram* P = out;
  *P = FusedProgram{};
  P->N = s.N; P->TG = TG; P->R = TG * s.N; P->CH = s.CH; P->F = s.F;
  P->G = s.G;
  P->TGp = slot ? ((TG + 3) & ~3) : 0;
  P->RP = slot ? s.N * P->TGp : ((P->R + 3) & ~3);
  P->row_g = slot ? 1 : s.N;
  P->row_n = slot ? P->TGp : 1;
  P->Dn = s.Dn; P->De = s.De; P->n_params = (int)s.n_params; P->train = train;
  Builder b; b.P = P;
  const int F = s.F, Dn = s.Dn, De = s.De, S = s.S;

  P->zero_row = b.fresh(1)[0];
  std::vector<int> x0 = b.fresh(Dn + De);
  P->x0_row0 = x0[0];
  std::vector<int> nodeR(x0.begin(), x0.begin() + Dn), edgeR(x0.begin() + Dn, x0.end());
  std::vector<std::vector<int>> h(S), a(S);
  std::vector<std::vector<int>> gemm_in(s.n_layers);
  // ---- forward
  for (int st = 0; st < S; ++st) {
    h[st] = b.fresh(F);
    a[st] = b.fresh(F);
    gemm_in[st] = (st == 0) ? x0 : cat({h[st - 1], nodeR, edgeR, a[st - 1]});
    FusedOp* g = b.add_op();
    g->type = FOP_GEMM; g->in_tab = b.tab_put(gemm_in[st]); g->K = (int)gemm_in[st].size();
    g->out_tab = b.tab_put(h[st]); g->O = F; g->w_off = (int)s.w_off[st]; g->b_off = (int)s.b_off[st];
    g->slot_w = s.layer_K[st] * s.layer_O[st]; g->slot_b = s.layer_O[st];
    g->w_span = s.G * (g->slot_w + g->slot_b);
    g->relu = st < S - 1;
    FusedOp* ag = b.add_op();
    ag->type = FOP_AGG; ag->in_tab = b.tab_put(h[st]); ag->K = F; ag->out_tab = b.tab_put(a[st]); ag->O = F;
  }
  const int widths[4] = {s.H1, s.H2, s.H3, s.CH};
  std::vector<std::vector<int>> mrows(4);
  for (int j = 0; j < 4; ++j) {
    const int l = S + j;
    mrows[j] = b.fresh(widths[j]);
    if (j == 0) P->stage_row0 = (!slot && Dn + De + s.CH <= widths[0]) ? mrows[0][0] : -1;
    gemm_in[l] = (j == 0) ? cat({nodeR, h[S - 1], a[S - 1]}) : mrows[j - 1];
    V2V_REQUIRE((int)gemm_in[l].size() == s.layer_K[l], "fused path: layer %d input width mismatch", l);
    FusedOp* g = b.add_op();
    g->type = FOP_GEMM; g->in_tab = b.tab_put(gemm_in[l]); g->K = (int)gemm_in[l].size();
    g->out_tab = b.tab_put(mrows[j]); g->O = widths[j]; g->w_off = (int)s.w_off[l]; g->b_off = (int)s.b_off[l];
    g->slot_w = s.layer_K[l] * s.layer_O[l]; g->slot_b = s.layer_O[l];
    g->w_span = s.G * (g->slot_w + g->slot_b);
    g->relu = j < 3;
  }
  P->q_tab = b.tab_put(mrows[3]);
  if (train) {
    std::vector<int> yR = b.fresh(s.CH);
    P->y_tab = b.tab_put(yR);
    FusedOp* lo = b.add_op();
    lo->type = FOP_LOSS;
    b.release(yR);   // dead after the loss phase
    int blk = 0, bias = 0;
    // the layer with the most 4x4 blocks (<= one per thread) keeps them in registers across tiles; every other
    // layer row-splits its blocks and accumulates in the CTA's shared buffer
    int reg_layer = -1, reg_blocks = 0;
    for (int l = 0; l < s.n_layers; ++l) {
      const int K_act = (l == 0) ? (Dn + De) : s.layer_K[l];
      const int nb = ((K_act + 3) / 4) * (s.layer_O[l] / 4);
      if (nb <= kFusedThreads * kFusedBlkPerThread && nb > reg_blocks) { reg_blocks = nb; reg_layer = l; }
    }
    auto add_bwd = [&](int l, const std::vector<int>& dz, const std::vector<int>& dxk, const std::vector<int>& dx,
                       const std::vector<int>& gate) -> int {
      FusedOp* o = b.add_op();
      const int op_idx = P->n_ops - 1;
      o->type = FOP_BWD;
      o->in_tab = b.tab_put(gemm_in[l]); o->K = (int)gemm_in[l].size();
      o->O = s.layer_O[l]; o->w_off = (int)s.w_off[l]; o->b_off = (int)s.b_off[l];
      o->slot_w = s.layer_K[l] * s.layer_O[l]; o->slot_b = s.layer_O[l];
      o->w_span = s.G * (o->slot_w + o->slot_b);
      o->dz_tab = b.tab_put(dz);
      o->n_dx = (int)dxk.size();
      if (o->n_dx) { o->dxk_tab = b.tab_put(dxk); o->dx_tab = b.tab_put(dx); }
      o->gate_tab = gate.empty() ? -1 : b.tab_put(gate);
      o->blk0 = blk;
      const int kb_n = (o->K + 3) / 4, ob_n = o->O / 4;
      o->nblk = kb_n * ob_n;
      if (slot) {                  // per-slot weights: blocks go straight to the CTA's partial row (bwd_phase<SLOT>)
        o->blk0 = -1; o->nblk = 0; o->wt0 = 0; o->nwt = 0; o->bias0 = 0;
        return op_idx;
      }
      const bool small = (l != reg_layer);
      if (small) {                 // row-split weight gradient into the shared accumulator (wt0 = offset, nwt = RS)
        int rs = 1;
        while (rs < 8 && o->nblk * rs * 2 <= kFusedThreads) rs *= 2;
        o->wt0 = P->n_small;
        o->nwt = rs;
        P->n_small += (o->K * o->O + 3) & ~3;
        o->blk0 = -1;
      }
      for (int kb = 0; kb < kb_n && !small; ++kb)
        for (int ob = 0; ob < ob_n; ++ob) {
          if (blk >= kFusedThreads * kFusedBlkPerThread) { b.overflow = true; break; }
          P->blk_info[blk++] = (op_idx << 16) | (kb << 8) | ob;
        }
      o->bias0 = bias;
      for (int oo = 0; oo < o->O; ++oo) {
        if (bias >= kFusedThreads) { b.overflow = true; break; }
        P->bias_info[bias++] = (op_idx << 16) | oo;
      }
      return op_idx;
    };
    // decision MLP, last layer first; dz of layer j lives in `dz`
    std::vector<int> dz = mrows[3];                               // dq in place over q
    for (int j = 3; j >= 1; --j) {
      std::vector<int> dx = b.take(widths[j - 1]);
      add_bwd(S + j, dz, iota(0, widths[j - 1]), dx, mrows[j - 1]);
      b.release(dz);
      b.release(mrows[j - 1]);
      dz = dx;
    }
    // first MLP layer: gradients w.r.t. the h and agg columns of [node | h | agg]
    std::vector<int> dh = b.take(F), da = b.take(F);
    add_bwd(S, dz, iota(Dn, 2 * F), cat({dh, da}), {});
    b.release(dz);
    b.release(a[S - 1]);
    b.release(h[S - 1]);
    {
      FusedOp* ag = b.add_op();                                   // dh += Agg^T(da)   (last stage is linear: no gate)
      ag->type = FOP_AGG; ag->transposed = 1; ag->in_tab = b.tab_put(da); ag->K = F;
      ag->add_tab = b.tab_put(dh); ag->out_tab = ag->add_tab; ag->O = F;
    }
    b.release(da);
    for (int st = S - 1; st >= 1; --st) {
      std::vector<int> dxh = b.take(F), dxa = b.take(F);
      std::vector<int> kk = cat({iota(0, F), iota(F + Dn + De, F)});
      add_bwd(st, dh, kk, cat({dxh, dxa}), {});
      b.release(dh);
      b.release(a[st - 1]);
      FusedOp* ag = b.add_op();                                   // dh(st-1) = relu'(h(st-1)) * (dxh + Agg^T(dxa))
      ag->type = FOP_AGG; ag->transposed = 1; ag->in_tab = b.tab_put(dxa); ag->K = F;
      ag->add_tab = b.tab_put(dxh); ag->out_tab = ag->add_tab; ag->O = F;
      ag->gate_tab = b.tab_put(h[st - 1]);
      b.release(dxa);
      b.release(h[st - 1]);
      dh = dxh;
    }
    add_bwd(0, dh, {}, {}, {});
    P->n_blocks = blk;
    P->n_bias = bias;
  }
  P->n_rows = b.next_row;
  if (slot) {          // stream the weights through shared memory when two buffers of the largest layer span fit beside a useful arena
    int span = 0;
    for (int l = 0; l < s.n_layers; ++l) {
      V2V_REQUIRE(s.b_off[l] == s.w_off[l] + (size_t)s.G * s.layer_K[l] * s.layer_O[l], "fused path: a layer's bias must follow its weights");
      span = std::max(span, s.G * (s.layer_K[l] * s.layer_O[l] + s.layer_O[l]));
    }
    span = (span + 3) & ~3;
    P->wstage_floats = (2 * span * 4 <= 112 * 1024) ? span : 0;
  }
  V2V_REQUIRE(!b.overflow, "fused path: program tables overflow");
  return 0;
}

size_t fused_smem_bytes(const FusedProgram& p) {
  size_t words = 0;
  if (p.G == 1) words += (p.n_params + 3) & ~3;       // per-slot weights stay in global memory
  words += (p.n_tab + 3) & ~3;
  words += (size_t)p.n_ops * (sizeof(FusedOp) / 4);
  words += (2 * p.TG * p.N + 3) & ~3;
  words += 32;
  words += p.n_small;
  words += 2 * (size_t)p.wstage_floats;
  words += (size_t)p.n_rows * p.RP;
  return words * 4;
}

int fused_pick_tg(const FusedShape& s, int B, int train) {
  // largest tile that fits shared memory, then the TG that minimises the per-SM critical path
  FusedProgram* tmp = new FusedProgram();
  int fit = 0;
  for (int tg = 1; tg <= 64; ++tg) {
    if (fused_build_program(s, tg, train, tmp) != 0) break;
    if (fused_smem_bytes(*tmp) > 226 * 1024) break;
    fit = tg;
  }
  delete tmp;
  if (fit == 0) return 0;
  const int sms = sm_count();
  int best = 1;
  double best_cost = 1e30;
  for (int tg = 1; tg <= fit; ++tg) {
    const int tiles = ceil_div(B, tg);
    const int rounds = ceil_div(tiles, sms);
    const int rp = s.G > 1 ? s.N * ((tg + 3) & ~3) : ((tg * s.N + 3) & ~3);
    const double cost = rounds * (rp + 24.0);        // rows per tile + fixed per-tile overhead (barriers, loads)
    if (cost < best_cost - 1e-9) { best_cost = cost; best = tg; }
  }
  return best;
}

int fused_grid(const FusedProgram& p, int B) { return std::max(1, std::min(ceil_div(B, p.TG), sm_count())); }

template <int NMAX, bool SLOT, int MMA>
static int fused_launch_t(const FusedProgram& ph, const FusedProgram* prog_dev, const float* params, const float* node,
                          const float* edge, const uint32_t* in_mask, const uint32_t* out_mask, const float* y, float* q_out,
                          float* partial_dev, float* head_loss, int B, int grid, cudaStream_t st) {
  const size_t smem = fused_smem_bytes(ph);
  const float inv_cnt = 1.f / ((float)B * (float)ph.CH);
  static size_t smem_set = 0;                  // per instantiation (NMAX): every kernel needs its own opt-in
  if (smem > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(fused_brain_kernel<NMAX, SLOT, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(kFusedThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, fused_brain_kernel<NMAX, SLOT, MMA>, prog_dev, params, node, edge, in_mask, out_mask, y, q_out,
                                    partial_dev, head_loss, B, inv_cnt));
  return launch_status("fused_brain_kernel");
}

int fused_launch(const FusedProgram& ph, const FusedProgram* prog_dev, const float* params, const float* node,
                 const float* edge, const uint32_t* in_mask, const uint32_t* out_mask, const float* y, float* q_out,
                 float* partial_dev, float* head_loss, int B, int grid, cudaStream_t st) {
#define V2V_FUSED_ARGS ph, prog_dev, params, node, edge, in_mask, out_mask, y, q_out, partial_dev, head_loss, B, grid, st
  if (ph.G > 1) {                                        // per-slot weights (the reference's model; N <= 8 here)
    if (ph.N <= 4) return fused_launch_t<4, true, 0>(V2V_FUSED_ARGS);
    return fused_launch_t<8, true, 0>(V2V_FUSED_ARGS);
  }
  const int mma = fused_get_mma();
#define V2V_FUSED_BY_N(M)                                                  \
  do {                                                                     \
    if (ph.N <= 4) return fused_launch_t<4, false, M>(V2V_FUSED_ARGS);     \
    if (ph.N <= 8) return fused_launch_t<8, false, M>(V2V_FUSED_ARGS);     \
    if (ph.N <= 20) return fused_launch_t<20, false, M>(V2V_FUSED_ARGS);   \
    return fused_launch_t<32, false, M>(V2V_FUSED_ARGS);                   \
  } while (0)
  if (mma == 1) V2V_FUSED_BY_N(1);
  V2V_FUSED_BY_N(0);
#undef V2V_FUSED_BY_N
#undef V2V_FUSED_ARGS
}

// Which pipe runs the backward contractions of the shared-weight kernel: 0 FP32 pipe, 1 tensor cores.
// Process-wide (the choice does not change any result beyond rounding); V2V_FUSED_MMA overrides the default at first use.
static int g_fused_mma = -1;
int fused_get_mma() {
  if (g_fused_mma < 0) {
    int m = kFusedMmaDefault;
    if (const char* e = getenv("V2V_FUSED_MMA")) m = atoi(e);
    g_fused_mma = (m < 0 || m > 1) ? kFusedMmaDefault : m;
  }
  return g_fused_mma;
}
int fused_set_mma(int mode) {
  V2V_REQUIRE(mode >= 0 && mode <= 1, "fused_set_mma: mode %d outside [0,1]", mode);
  g_fused_mma = mode;
  return 0;
}

int fused_set_trace(long long* dev_buf) {
  V2V_CHECK_CUDA(cudaMemcpyToSymbol(g_fused_trace, &dev_buf, sizeof(dev_buf)));
  return 0;
}

int fused_reduce_adam(const float* partial, int n_cta, long stride, float* grad, float* p, float* m, float* v, long n,
                      int n_tail, float* tail_out, int t, float lr, float b1, float b2, float eps, float gscale,
                      cudaStream_t st) {
  double lr_t = 0.0;
  if (t >= 1) lr_t = (double)lr * (sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)((n + n_tail + kRedCols - 1) / kRedCols));
  lc.blockDim = dim3(kRedCols * kRedSlices);
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const int do_adam = t >= 1;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, reduce_adam_kernel, partial, n_cta, stride, grad, p, m, v, n, n_tail, tail_out,
                                    do_adam, (float)lr_t, b1, b2, eps, gscale));
  return launch_status("reduce_adam_kernel");
}

}  // namespace v2v
