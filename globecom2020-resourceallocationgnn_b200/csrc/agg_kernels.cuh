// Device code of the neighbour-aggregation fast path (see agg.cu for the operator contract).
//
//   out[b][m][:] = sum_n Adj[b][n][m] * H[b][n][:]   (+ addend[b][m][:]),   F == 16, N <= 32
//
// Structure (sm_100a):
//   * every CTA owns a contiguous range of graphs, every WARP is an autonomous pipeline over
//     tiles of TG consecutive graphs of that range (a tile of H is one contiguous span);
//   * the span, the tile's mask words and (optionally) the addend span are fetched by 1-D
//     bulk-async copies (TMA engine; SASS UBLKCP) into a per-warp shared-memory ring, each stage
//     signalled through its own mbarrier -- no thread ever issues a global load;
//   * lane = (feature quad c, target partition mp, graph gl).  A lane keeps MT float4
//     accumulators in registers, walks the N source rows once (one conflict-free LDS.128 each)
//     and adds the row into every accumulator whose mask bit is set: one LOP3 producing the
//     predicate + two predicated packed FADD2 per (target, source) pair;
//   * results overwrite the consumed stage in place (after a __syncwarp) and leave through a
//     bulk-async store, so global traffic is exactly read-H-once + write-out-once;
//   * launched with programmatic dependent launch: the prologue (barrier init, descriptor math)
//     overlaps the tail of the previous kernel; `griddepcontrol.wait` guards the first global
//     access unless the caller declares the operands independent of the preceding kernel.
#pragma once
#include "v2v_common.cuh"

namespace v2v {

constexpr int kAggStages = 2;       // smem ring depth per warp

struct AggTileSizes {
  int h_bytes;       // TG * N * 16 * sizeof(T)
  int mask_off;      // offset of the mask words inside a stage
  int stage_bytes;   // h (+ addend) + mask, 128-B aligned
  int warp_bytes;    // stages * stage
};

template <typename T>
__host__ __device__ inline AggTileSizes agg_tile_sizes(int N, int TG, bool has_addend) {
  AggTileSizes s;
  s.h_bytes = TG * N * 16 * (int)sizeof(T);
  s.mask_off = s.h_bytes * (has_addend ? 2 : 1);
  s.stage_bytes = (s.mask_off + TG * N * 4 + 127) & ~127;
  s.warp_bytes = kAggStages * s.stage_bytes;
  return s;
}

__device__ __forceinline__ float4 ld_row4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld_row4(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float4 r;
  r.x = __uint_as_float(u.x << 16);
  r.y = __uint_as_float(u.x & 0xffff0000u);
  r.z = __uint_as_float(u.y << 16);
  r.w = __uint_as_float(u.y & 0xffff0000u);
  return r;
}
__device__ __forceinline__ void st_row4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_row4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ void add4(float4& a, const float4& v) {
  float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(v.x, v.y));
  float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(v.z, v.w));
  a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
}

__device__ __forceinline__ void sub4(float4& a, const float4& v);          // a = a - v, packed (defined below)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// T: storage type. MT: targets per lane. MP: target partitions (lanes per graph = 4*MP).
// WARPS: warps per CTA.  COMPUTE=false turns the kernel into its own data-movement floor (bench only).
// NC > 0: N is the compile-time constant NC (source walk fully unrolled: row offsets and mask bits become immediates).
// WALK: per target, walk only the set bits of its mask -- or, when more than half are set (the reference adjacency has
// in-degree N - 2: everyone but the node itself and its own receiver, BS_brain.py:441-445), the CLEAR bits, subtracted
// from the graph's column total (which the MP target partitions of a graph build together: N / MP rows each + shuffles).
// ~90 instead of ~320 instructions per lane and tile at N = 20, for the reference-dense and the sparse variant alike.
template <typename T, int MT, int MP, bool ADD, int WARPS, bool COMPUTE = true, int NC = 0, bool WALK = false>
__global__ void __launch_bounds__(WARPS * 32)
agg_mask_f16_kernel(const T* __restrict__ H, const uint32_t* __restrict__ mask,
                    const T* __restrict__ addend, T* __restrict__ out, int B, int N_rt, int dep_wait) {
  const int N = NC ? NC : N_rt;
  constexpr int TG = 32 / (4 * MP);           // graphs per warp tile
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[WARPS][kAggStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane & 3, mp = (lane >> 2) % MP, gl = lane / (4 * MP);
  const AggTileSizes ts = agg_tile_sizes<T>(N, TG, ADD);
  uint8_t* wbase = smem_raw + (size_t)warp * ts.warp_bytes;

  // contiguous graph range of this CTA, tiles of TG graphs inside it, warps interleave over tiles
  const int g_begin = (int)(((long)B * blockIdx.x) / gridDim.x);
  const int g_end = (int)(((long)B * (blockIdx.x + 1)) / gridDim.x);
  const int num_tiles = (g_end - g_begin + TG - 1) / TG;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kAggStages; ++s) mbar_init(&bars[warp][s], 1);
    fence_async_smem();                        // make the inits visible to the async (TMA) proxy
  }
  __syncwarp();
  pdl_launch_dependents();
  if (dep_wait) pdl_wait();

  const size_t graph_elems = (size_t)N * 16;

  auto issue = [&](int t, int s) {     // lane 0 only
    const int g0 = g_begin + t * TG;
    const int ng = min(TG, g_end - g0);
    const uint32_t hb = (uint32_t)(ng * N * 16 * sizeof(T));
    const uint32_t mb = (uint32_t)(ng * N * 4);
    const bool mask_bulk = ((((uint32_t)g0 * (uint32_t)N * 4u) | mb) & 15u) == 0;
    uint8_t* st = wbase + s * ts.stage_bytes;
    const uint32_t tx = hb + (ADD ? hb : 0) + (mask_bulk ? mb : 0);
    mbar_arrive_expect_tx(&bars[warp][s], tx);
    bulk_g2s(st, H + (size_t)g0 * graph_elems, hb, &bars[warp][s]);
    if (ADD) bulk_g2s(st + ts.h_bytes, addend + (size_t)g0 * graph_elems, hb, &bars[warp][s]);
    if (mask_bulk) bulk_g2s(st + ts.mask_off, mask + (size_t)g0 * N, mb, &bars[warp][s]);
  };

  int tile = warp;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kAggStages; ++s) {
      const int t = tile + s * WARPS;
      if (t < num_tiles) issue(t, s);
    }
  }

  uint32_t phase = 0;
  int stage = 0;
  for (; tile < num_tiles; tile += WARPS) {
    const int g0 = g_begin + tile * TG;
    const int ng = min(TG, g_end - g0);
    uint8_t* st = wbase + stage * ts.stage_bytes;
    T* Hs = reinterpret_cast<T*>(st);
    T* As = reinterpret_cast<T*>(st + ts.h_bytes);
    uint32_t* Ms = reinterpret_cast<uint32_t*>(st + ts.mask_off);

    if (((((uint32_t)g0 * (uint32_t)N * 4u) | (uint32_t)(ng * N * 4)) & 15u) != 0) {
      // ragged / unaligned tile: mask words by plain loads (the span itself is always 16-B aligned)
      for (int i = lane; i < ng * N; i += 32) Ms[i] = mask[(size_t)g0 * N + i];
      __syncwarp();
    }
    mbar_wait(&bars[warp][stage], (phase >> stage) & 1u);

    float4 acc[MT];
    if (COMPUTE) {
      uint32_t msk[MT];
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const int m = j * MP + mp;
        const bool ok = (m < N) && (gl < ng);
        msk[j] = ok ? Ms[gl * N + m] : 0u;
        if (ADD) {
          acc[j] = ok ? ld_row4(As + ((size_t)(gl * N + m) * 16 + c * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      const T* hrow = Hs + (size_t)gl * N * 16 + c * 4;
      if (WALK) {
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NC) {
          const T* hmp = hrow + mp * 16;                       // rows mp, mp + MP, ...: immediates off one base register
#pragma unroll
          for (int i = 0; i < (NC + MP - 1) / MP; ++i)
            if (NC % MP == 0 || i * MP + mp < NC) add4(tot, ld_row4(hmp + i * MP * 16));
        } else {
          for (int n = mp; n < N; n += MP) add4(tot, ld_row4(hrow + n * 16));
        }
#pragma unroll
        for (int off = 4; off < 4 * MP; off <<= 1) {          // the MP partitions of a graph sit in lane bits 2..: same c, same gl
          tot.x += __shfl_xor_sync(0xffffffffu, tot.x, off); tot.y += __shfl_xor_sync(0xffffffffu, tot.y, off);
          tot.z += __shfl_xor_sync(0xffffffffu, tot.z, off); tot.w += __shfl_xor_sync(0xffffffffu, tot.w, off);
        }
        const uint32_t valid = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const bool dense = 2 * __popc(msk[j]) > N;
          uint32_t bits = dense ? (~msk[j] & valid) : msk[j];
          if (j * MP + mp >= N || gl >= ng) bits = 0u;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          while (bits) {
            const int n = __ffs(bits) - 1;
            bits &= bits - 1;
            add4(a, ld_row4(hrow + n * 16));
          }
          if (dense) { float4 r = tot; sub4(r, a); a = r; }
          add4(acc[j], a);
        }
      } else if (gl < ng) {
        if (NC) {
#pragma unroll
          for (int n = 0; n < NC; ++n) {
            const float4 v = ld_row4(hrow + n * 16);
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              if (msk[j] & (1u << n)) add4(acc[j], v);
            }
          }
        } else {
#pragma unroll 2
          for (int n = 0; n < N; ++n) {
            const float4 v = ld_row4(hrow + n * 16);
            const uint32_t bit = 1u << n;
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              if (msk[j] & bit) add4(acc[j], v);
            }
          }
        }
      }
      __syncwarp();                      // every lane is done reading the tile: overwrite it in place
      T* Os = ADD ? As : Hs;
#pragma unroll
      for (int j = 0; j < MT; ++j) {
        const int m = j * MP + mp;
        if (m < N && gl < ng) st_row4(Os + ((size_t)(gl * N + m) * 16 + c * 4), acc[j]);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(out + (size_t)g0 * graph_elems, ADD ? (void*)As : (void*)Hs, (uint32_t)(ng * N * 16 * sizeof(T)));
      bulk_commit();
      const int nt = tile + kAggStages * WARPS;   // refill this stage once its store has drained smem
      if (nt < num_tiles) {
        bulk_wait_read<0>();
        issue(nt, stage);
      }
    }
    phase ^= (1u << stage);
    stage = (stage + 1 == kAggStages) ? 0 : stage + 1;
  }
  if (lane == 0) bulk_wait_read<0>();          // smem must outlive the last store's reads
}

// ---------------------------------------------------------------------------
// Large graphs (20 < N <= 256): same TMA ring, but the N^2 predicated adds of the dense form would make the kernel
// FP32-bound (26 adds per byte at N = 256).  Per target the kernel walks only the set bits of its mask -- or, when
// more than half of the bits are set (the reference adjacency has in-degree N-2, BS_brain.py:441-445), the CLEAR
// bits, subtracting them from the graph's column total.  Work is O(N * min(deg, N - deg)) and the kernel stays
// HBM-bound for both the reference-dense and the sparse variants.  (fp32 rounding differs from the sequential sum
// by a few ulp of sum|H|, far inside the 1e-4 parity bar.)
// The walk is the issue-bound part (ncu: 65 % issue-active, 36 M warp instructions per launch at N = 128 before this
// form), so W = ceil(N/32) is a template parameter (mask words live in registers, loops unroll), every shared-memory
// access uses a 32-bit shared-window address computed once per graph, and the adds are packed FADD2.
// ---------------------------------------------------------------------------
template <typename T> struct RowBytes { static constexpr int v = 16 * (int)sizeof(T); };

__device__ __forceinline__ float4 lds_row4(uint32_t addr, const float*) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_row4(uint32_t addr, const __nv_bfloat16*) {
  uint2 u;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(u.x), "=r"(u.y) : "r"(addr));
  float4 r;
  r.x = __uint_as_float(u.x << 16);
  r.y = __uint_as_float(u.x & 0xffff0000u);
  r.z = __uint_as_float(u.y << 16);
  r.w = __uint_as_float(u.y & 0xffff0000u);
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sub4(float4& a, const float4& v) {          // a = a - v, packed
  float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(-v.x, -v.y));
  float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(-v.z, -v.w));
  a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
}

// One target: sum of the H rows selected by its mask words (set bits, or total minus the clear bits).
// h_addr: shared address of (row 0, this lane's feature quad); m_addr: shared address of the target's W mask words.
template <typename T, int W>
__device__ __forceinline__ float4 walk_target(uint32_t h_addr, uint32_t m_addr, int N, uint32_t last_valid, const float4& tot) {
  uint32_t wds[W];
  int cnt = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) { wds[w] = lds_u32(m_addr + 4 * w); cnt += __popc(wds[w]); }
  const bool dense = 2 * cnt > N;
  if (dense) {
#pragma unroll
    for (int w = 0; w < W; ++w) wds[w] = ~wds[w] & (w == W - 1 ? last_valid : 0xffffffffu);
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t bits = wds[w];
    const uint32_t base = h_addr + (uint32_t)(w * 32 * RowBytes<T>::v);
    while (bits) {
      const int n = __ffs(bits) - 1;
      bits &= bits - 1;
      add4(acc, lds_row4(base + (uint32_t)n * RowBytes<T>::v, (const T*)nullptr));
    }
  }
  if (dense) { float4 r = tot; sub4(r, acc); return r; }
  return acc;
}

struct AggSparseSizes {
  int h_bytes, mask_off, stage_bytes, out_bytes, warp_bytes;
};
template <typename T>
__host__ __device__ inline AggSparseSizes agg_sparse_sizes(int N, int W, int TG, bool has_addend) {
  AggSparseSizes s;
  s.h_bytes = TG * N * 16 * (int)sizeof(T);
  s.mask_off = s.h_bytes * (has_addend ? 2 : 1);
  s.stage_bytes = (s.mask_off + TG * N * W * 4 + 127) & ~127;
  s.out_bytes = (s.h_bytes + 127) & ~127;
  s.warp_bytes = kAggStages * s.stage_bytes + s.out_bytes;
  return s;
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

template <typename T, bool ADD, int WARPS, int W>
__global__ void __launch_bounds__(WARPS * 32)
agg_sparse_f16_kernel(const T* __restrict__ H, const uint32_t* __restrict__ mask, const T* __restrict__ addend,
                      T* __restrict__ out, int B, int N, int TG, int dep_wait) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[WARPS][kAggStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane & 3, tl = lane >> 2;                       // feature quad, target lane (8 targets in flight)
  const AggSparseSizes ts = agg_sparse_sizes<T>(N, W, TG, ADD);
  uint8_t* wbase = smem_raw + (size_t)warp * ts.warp_bytes;
  uint8_t* out_s = wbase + kAggStages * ts.stage_bytes;
  const int g_begin = (int)(((long)B * blockIdx.x) / gridDim.x);
  const int g_end = (int)(((long)B * (blockIdx.x + 1)) / gridDim.x);
  const int num_tiles = (g_end - g_begin + TG - 1) / TG;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kAggStages; ++s) mbar_init(&bars[warp][s], 1);
    fence_async_smem();
  }
  __syncwarp();
  pdl_launch_dependents();
  if (dep_wait) pdl_wait();
  const size_t graph_elems = (size_t)N * 16;
  constexpr int RB = RowBytes<T>::v;
  auto issue = [&](int t, int s) {     // lane 0 only
    const int g0 = g_begin + t * TG;
    const int ng = min(TG, g_end - g0);
    const uint32_t hb = (uint32_t)(ng * N * 16 * sizeof(T));
    const uint32_t mb = (uint32_t)(ng * N * W * 4);
    const bool mask_bulk = ((((uint32_t)g0 * (uint32_t)(N * W) * 4u) | mb) & 15u) == 0;
    uint8_t* st = wbase + s * ts.stage_bytes;
    mbar_arrive_expect_tx(&bars[warp][s], hb + (ADD ? hb : 0) + (mask_bulk ? mb : 0));
    bulk_g2s(st, H + (size_t)g0 * graph_elems, hb, &bars[warp][s]);
    if (ADD) bulk_g2s(st + ts.h_bytes, addend + (size_t)g0 * graph_elems, hb, &bars[warp][s]);
    if (mask_bulk) bulk_g2s(st + ts.mask_off, mask + (size_t)g0 * N * W, mb, &bars[warp][s]);
  };
  int tile = warp;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kAggStages; ++s) {
      const int t = tile + s * WARPS;
      if (t < num_tiles) issue(t, s);
    }
  }
  const uint32_t last_valid = (N & 31) ? ((1u << (N & 31)) - 1u) : 0xffffffffu;
  const uint32_t out_addr = smem_u32(out_s) + (uint32_t)(c * 4 * sizeof(T));
  uint32_t phase = 0;
  int stage = 0;
  for (; tile < num_tiles; tile += WARPS) {
    const int g0 = g_begin + tile * TG;
    const int ng = min(TG, g_end - g0);
    uint8_t* st = wbase + stage * ts.stage_bytes;
    uint32_t* Ms = reinterpret_cast<uint32_t*>(st + ts.mask_off);
    if (((((uint32_t)g0 * (uint32_t)(N * W) * 4u) | (uint32_t)(ng * N * W * 4)) & 15u) != 0) {
      for (int i = lane; i < ng * N * W; i += 32) Ms[i] = mask[(size_t)g0 * N * W + i];
      __syncwarp();
    }
    mbar_wait(&bars[warp][stage], (phase >> stage) & 1u);
    if (lane == 0) bulk_wait_read<0>();            // the previous tile's store has drained out_s
    __syncwarp();
    const uint32_t st_addr = smem_u32(st);
    for (int g = 0; g < ng; ++g) {
      const uint32_t h_addr = st_addr + (uint32_t)(g * N * RB + c * 4 * (int)sizeof(T));
      const uint32_t a_addr = h_addr + (uint32_t)ts.h_bytes;
      const uint32_t m_addr = st_addr + (uint32_t)(ts.mask_off + g * N * W * 4);
      float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int n = tl; n < N; n += 8) add4(tot, lds_row4(h_addr + (uint32_t)(n * RB), (const T*)nullptr));
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        tot.x += __shfl_xor_sync(0xffffffffu, tot.x, off); tot.y += __shfl_xor_sync(0xffffffffu, tot.y, off);
        tot.z += __shfl_xor_sync(0xffffffffu, tot.z, off); tot.w += __shfl_xor_sync(0xffffffffu, tot.w, off);
      }
      for (int m = tl; m < N; m += 8) {
        float4 r = walk_target<T, W>(h_addr, m_addr + (uint32_t)(m * W * 4), N, last_valid, tot);
        if (ADD) add4(r, lds_row4(a_addr + (uint32_t)(m * RB), (const T*)nullptr));
        st_row4(reinterpret_cast<T*>(out_s) + ((size_t)(g * N + m) * 16 + c * 4), r);
      }
    }
    (void)out_addr;
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(out + (size_t)g0 * graph_elems, out_s, (uint32_t)(ng * N * 16 * sizeof(T)));
      bulk_commit();
      const int nt = tile + kAggStages * WARPS;
      if (nt < num_tiles) issue(nt, stage);          // the stage itself was only read: refill immediately
    }
    phase ^= (1u << stage);
    stage = (stage + 1 == kAggStages) ? 0 : stage + 1;
  }
  if (lane == 0) bulk_wait_read<0>();
}

// ---------------------------------------------------------------------------
// Very large graphs (N >= 96): one CTA per graph at a time.  The graph's H rows and mask words arrive by bulk-async
// copy into a 2-stage ring (the next graph loads while this one is walked); 256 threads = 64 targets x 4 feature quads
// walk the set / clear bits as above; results go straight to global memory (a warp writes 8 consecutive 64-byte rows).
// ---------------------------------------------------------------------------
template <typename T, bool ADD, int W>
__global__ void __launch_bounds__(256)
agg_block_f16_kernel(const T* __restrict__ H, const uint32_t* __restrict__ mask, const T* __restrict__ addend,
                     T* __restrict__ out, int B, int N, int dep_wait) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ float4 tot_s[8][4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = tid & 3, part = tid >> 2;                         // 64 targets in flight
  constexpr int RB = RowBytes<T>::v;
  const size_t graph_elems = (size_t)N * 16;
  const uint32_t hb = (uint32_t)(graph_elems * sizeof(T));
  const uint32_t mb = (uint32_t)(N * W * 4);
  const uint32_t stage_bytes = (hb + mb + 127u) & ~127u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_async_smem(); }
  __syncthreads();
  pdl_launch_dependents();
  if (dep_wait) pdl_wait();
  const uint32_t last_valid = (N & 31) ? ((1u << (N & 31)) - 1u) : 0xffffffffu;
  // masks whose global offset is not 16-byte aligned cannot use the bulk copy: the whole CTA copies them instead
  const bool mask_bulk_all = (mb & 15u) == 0;       // g * mb stays 16-byte aligned for every g
  auto issue = [&](int g, int s) {                  // thread 0 only
    uint8_t* st = smem_raw + (size_t)s * stage_bytes;
    mbar_arrive_expect_tx(&bar[s], hb + (mask_bulk_all ? mb : 0));
    bulk_g2s(st, H + (size_t)g * graph_elems, hb, &bar[s]);
    if (mask_bulk_all) bulk_g2s(st + hb, mask + (size_t)g * N * W, mb, &bar[s]);
  };
  int g = blockIdx.x;
  if (tid == 0 && g < B) {
    issue(g, 0);
    if (g + (int)gridDim.x < B) issue(g + gridDim.x, 1);
  }
  uint32_t phase = 0;
  int stage = 0;
  for (; g < B; g += gridDim.x) {
    uint8_t* st = smem_raw + (size_t)stage * stage_bytes;
    if (!mask_bulk_all) {
      uint32_t* Ms = reinterpret_cast<uint32_t*>(st + hb);
      for (int i = tid; i < N * W; i += 256) Ms[i] = mask[(size_t)g * N * W + i];
      __syncthreads();
    }
    mbar_wait(&bar[stage], (phase >> stage) & 1u);
    const uint32_t h_addr = smem_u32(st) + (uint32_t)(c * 4 * sizeof(T));
    const uint32_t m_addr = smem_u32(st) + hb;
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = part; n < N; n += 64) add4(tot, lds_row4(h_addr + (uint32_t)(n * RB), (const T*)nullptr));
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      tot.x += __shfl_xor_sync(0xffffffffu, tot.x, off); tot.y += __shfl_xor_sync(0xffffffffu, tot.y, off);
      tot.z += __shfl_xor_sync(0xffffffffu, tot.z, off); tot.w += __shfl_xor_sync(0xffffffffu, tot.w, off);
    }
    if (lane < 4) tot_s[warp][lane] = tot;
    __syncthreads();
    tot = tot_s[0][c];
#pragma unroll
    for (int w8 = 1; w8 < 8; ++w8) add4(tot, tot_s[w8][c]);
    const size_t obase = (size_t)g * graph_elems + c * 4;
    for (int m = part; m < N; m += 64) {
      float4 r = walk_target<T, W>(h_addr, m_addr + (uint32_t)(m * W * 4), N, last_valid, tot);
      const size_t o = obase + (size_t)m * 16;
      if (ADD) r = f4_add(r, ld_row4(addend + o));
      st_row4(out + o, r);
    }
    __syncthreads();                     // everyone is done with this stage (and tot_s): refill it two graphs ahead
    const int gn = g + 2 * (int)gridDim.x;
    if (tid == 0 && gn < B) issue(gn, stage);
    phase ^= (1u << stage);
    stage ^= 1;
  }
}

template <typename T, bool ADD, int W>
static int launch_agg_block_w(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N, bool dep_wait,
                              cudaStream_t st) {
  const size_t stage = (((size_t)N * 16 * sizeof(T) + (size_t)N * W * 4) + 127) & ~(size_t)127;
  const size_t smem = 2 * stage;
  if (smem > 200 * 1024) return -1;
  const int per_sm = std::max(1, std::min(8, (int)((220 * 1024) / (smem + 1024))));
  const int grid = std::max(1, std::min(B, sm_count() * per_sm));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(256);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const int dep = dep_wait ? 1 : 0;
  auto k = agg_block_f16_kernel<T, ADD, W>;
  static size_t smem_set = 0;
  if (smem > smem_set) { V2V_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, k, H, mask, addend, out, B, N, dep));
  return launch_status("agg_block_f16_kernel");
}

template <typename T>
static int launch_agg_block(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N, bool dep_wait,
                            cudaStream_t st) {
  const int W = ceil_div(N, 32);
#define V2V_BLOCK_CASE(WW)                                                                                         \
  case WW:                                                                                                         \
    return addend ? launch_agg_block_w<T, true, WW>(H, mask, addend, out, B, N, dep_wait, st)                      \
                  : launch_agg_block_w<T, false, WW>(H, mask, addend, out, B, N, dep_wait, st);
  switch (W) {
    V2V_BLOCK_CASE(1) V2V_BLOCK_CASE(2) V2V_BLOCK_CASE(3) V2V_BLOCK_CASE(4)
    V2V_BLOCK_CASE(5) V2V_BLOCK_CASE(6) V2V_BLOCK_CASE(7) V2V_BLOCK_CASE(8)
    default: return -1;
  }
#undef V2V_BLOCK_CASE
}

// ---------------------------------------------------------------------------
// launch helper (programmatic dependent launch)
// ---------------------------------------------------------------------------
struct AggLaunchCfg {
  int ctas_per_sm = 2;
  bool pdl = true;
  bool dep_wait = true;
};

template <typename T, int MT, int MP, bool ADD, int WARPS, bool COMPUTE = true, int NC = 0, bool WALK = false>
static int launch_agg_fast(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N,
                           const AggLaunchCfg& cfg, cudaStream_t st) {
  constexpr int TG = 32 / (4 * MP);
  AggTileSizes ts = agg_tile_sizes<T>(N, TG, ADD);
  const size_t smem = (size_t)ts.warp_bytes * WARPS;
  const int num_tiles = ceil_div(B, TG);
  const int grid = std::max(1, std::min(ceil_div(num_tiles, WARPS), sm_count() * cfg.ctas_per_sm));
  auto k = agg_mask_f16_kernel<T, MT, MP, ADD, WARPS, COMPUTE, NC, WALK>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(WARPS * 32);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = cfg.pdl ? 1 : 0;
  const int dep = cfg.dep_wait ? 1 : 0;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, k, H, mask, addend, out, B, N, dep));
  return launch_status("agg_mask_f16_kernel");
}

template <typename T, bool ADD, int WARPS, int W>
static int launch_agg_sparse_t(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N, int TG,
                               bool dep_wait, cudaStream_t st) {
  AggSparseSizes ts = agg_sparse_sizes<T>(N, W, TG, ADD);
  const size_t smem = (size_t)ts.warp_bytes * WARPS;
  const int num_tiles = ceil_div(B, TG);
  const int ctas_per_sm = std::max(1, (int)((227 * 1024) / (smem + 1024)));
  // Grid: a launch that waits for its predecessor wants many CTAs (short ramp: every byte in flight at once); a stream of
  // independent launches wants ONE CTA per SM, so that the next launch's CTAs are not queued behind this one's second
  // wave (measured at B = 8192, profiles/agg_variants_r02.txt: N = 64 0.685 -> 0.787, N = 40 0.707 -> 0.848 of the roofline
  // in the throughput regime; the dependent regime loses 25 % with one CTA per SM, hence the switch).
  const int cap = dep_wait ? 4 : 1;
  const int grid = std::max(1, std::min(ceil_div(num_tiles, WARPS), sm_count() * std::min(ctas_per_sm, cap)));
  auto k = agg_sparse_f16_kernel<T, ADD, WARPS, W>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    V2V_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(WARPS * 32);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const int dep = dep_wait ? 1 : 0;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, k, H, mask, addend, out, B, N, TG, dep));
  return launch_status("agg_sparse_f16_kernel");
}

template <typename T, int W>
static int launch_agg_sparse_w(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N, bool dep_wait,
                               cudaStream_t st) {
  const bool add = addend != nullptr;
  // ~2 KB of H per warp tile (one graph from N = 32 on): smaller tiles = more of them in flight per SM and a shorter
  // per-tile latency; 4 KB tiles were 15-20 % slower in both launch regimes at N = 32 (profiles/agg_variants_r02.txt)
  int TG = std::max(1, 2048 / (N * 16 * (int)sizeof(T)));
  while (TG > 1 && ((TG * N * W * 4) & 15)) --TG;               // keep the mask span 16-byte sized for the bulk copy
  AggSparseSizes ts = agg_sparse_sizes<T>(N, W, TG, add);
  const size_t budget = 224 * 1024;
  if ((size_t)ts.warp_bytes * 8 <= budget)
    return add ? launch_agg_sparse_t<T, true, 8, W>(H, mask, addend, out, B, N, TG, dep_wait, st)
               : launch_agg_sparse_t<T, false, 8, W>(H, mask, addend, out, B, N, TG, dep_wait, st);
  if ((size_t)ts.warp_bytes * 4 <= budget)
    return add ? launch_agg_sparse_t<T, true, 4, W>(H, mask, addend, out, B, N, TG, dep_wait, st)
               : launch_agg_sparse_t<T, false, 4, W>(H, mask, addend, out, B, N, TG, dep_wait, st);
  return -1;   // does not fit: caller falls back
}

// warp-per-tile bit walk (graphs per warp tile ~4 KB of H); W = ceil(N/32) selects the instantiation
template <typename T>
static int launch_agg_sparse(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N, bool dep_wait,
                             cudaStream_t st) {
  switch (ceil_div(N, 32)) {
    case 1: return launch_agg_sparse_w<T, 1>(H, mask, addend, out, B, N, dep_wait, st);
    case 2: return launch_agg_sparse_w<T, 2>(H, mask, addend, out, B, N, dep_wait, st);
    case 3: return launch_agg_sparse_w<T, 3>(H, mask, addend, out, B, N, dep_wait, st);
    default: return -1;
  }
}

template <typename T>
static bool agg_fast_fits(int N, int TG, bool add, int warps, int ctas_per_sm) {
  AggTileSizes ts = agg_tile_sizes<T>(N, TG, add);
  const size_t per_cta = (size_t)ts.warp_bytes * warps + 1024;      // + static barriers / reserve
  return per_cta * ctas_per_sm <= 227 * 1024;
}

}  // namespace v2v
