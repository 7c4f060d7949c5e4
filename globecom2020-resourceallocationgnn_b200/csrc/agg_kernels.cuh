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

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// T: storage type. MT: targets per lane. MP: target partitions (lanes per graph = 4*MP).
// WARPS: warps per CTA.  COMPUTE=false turns the kernel into its own data-movement floor (bench only).
template <typename T, int MT, int MP, bool ADD, int WARPS, bool COMPUTE = true>
__global__ void __launch_bounds__(WARPS * 32)
agg_mask_f16_kernel(const T* __restrict__ H, const uint32_t* __restrict__ mask,
                    const T* __restrict__ addend, T* __restrict__ out, int B, int N, int dep_wait) {
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
      if (gl < ng) {
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
// launch helper (programmatic dependent launch)
// ---------------------------------------------------------------------------
struct AggLaunchCfg {
  int ctas_per_sm = 2;
  bool pdl = true;
  bool dep_wait = true;
};

template <typename T, int MT, int MP, bool ADD, int WARPS, bool COMPUTE = true>
static int launch_agg_fast(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N,
                           const AggLaunchCfg& cfg, cudaStream_t st) {
  constexpr int TG = 32 / (4 * MP);
  AggTileSizes ts = agg_tile_sizes<T>(N, TG, ADD);
  const size_t smem = (size_t)ts.warp_bytes * WARPS;
  const int num_tiles = ceil_div(B, TG);
  const int grid = std::max(1, std::min(ceil_div(num_tiles, WARPS), sm_count() * cfg.ctas_per_sm));
  auto k = agg_mask_f16_kernel<T, MT, MP, ADD, WARPS, COMPUTE>;
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

template <typename T>
static bool agg_fast_fits(int N, int TG, bool add, int warps, int ctas_per_sm) {
  AggTileSizes ts = agg_tile_sizes<T>(N, TG, add);
  const size_t per_cta = (size_t)ts.warp_bytes * warps + 1024;      // + static barriers / reserve
  return per_cta * ctas_per_sm <= 227 * 1024;
}

}  // namespace v2v
