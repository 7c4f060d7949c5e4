// Neighbour aggregation over the V2V adjacency  (AggLayer.call, BS_brain.py:69-76)
//
//     out[b][m][:] = sum_n Adj[b][n][m] * H[b][n][:]   (+ addend[b][m][:])
//
// The reference evaluates this as (B,NF) x (B,NF,NF) against kron(Adj, I_F)
// (409,600 B/graph at N=20); here the adjacency is N bitmask words per graph and
// the kernel moves the algorithmic minimum: read H once, write out once.
//
// Fast path (fp32/bf16 storage, F == 16, N <= 32, 16-byte aligned tensors):
//   * persistent grid, one CTA per SM slot, every WARP is an autonomous pipeline
//     over tiles of TG consecutive graphs (their H rows are one contiguous span);
//   * the span is fetched by a 1-D bulk-async copy (TMA engine, UBLKCP) into a
//     per-warp 2-stage shared-memory ring signalled through an mbarrier, results
//     are staged in shared memory and written back by a bulk-async store, so the
//     LSU only sees conflict-free 128-bit shared-memory traffic;
//   * lane = (feature quad c, target partition mp, graph gl); a lane keeps MT
//     float4 accumulators in registers, walks the N source rows once (one
//     LDS.128 each) and adds the row into the accumulators whose mask bit is set
//     (predicated packed FADD2), i.e. a register-tiled gather-reduce.
// Generic path: any N <= 256 / F, one thread per output element.
#include "v2v_common.cuh"

namespace v2v {

// ---------------------------------------------------------------------------
// fast path
// ---------------------------------------------------------------------------
constexpr int kAggWarps = 8;        // warps per CTA
constexpr int kAggStages = 2;       // smem ring depth per warp

struct AggTileSizes {
  int h_bytes;       // TG * N * 16 * sizeof(T)
  int mask_bytes;    // TG * N * 4
  int stage_bytes;   // h + mask (+ addend), 128-B aligned
  int out_bytes;     // == h_bytes
  int warp_bytes;    // stages * stage + out
};

template <typename T>
__host__ __device__ inline AggTileSizes agg_tile_sizes(int N, int TG, bool has_addend) {
  AggTileSizes s;
  s.h_bytes = TG * N * 16 * (int)sizeof(T);
  s.mask_bytes = TG * N * 4;
  int st = s.h_bytes + (has_addend ? s.h_bytes : 0) + s.mask_bytes;
  s.stage_bytes = (st + 127) & ~127;
  s.out_bytes = (s.h_bytes + 127) & ~127;
  s.warp_bytes = kAggStages * s.stage_bytes + s.out_bytes;
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

// T: storage type. MT: targets per lane. MP: target partitions (lanes per graph = 4*MP).
template <typename T, int MT, int MP, bool ADD>
__global__ void __launch_bounds__(kAggWarps * 32, 1)
agg_mask_f16_kernel(const T* __restrict__ H, const uint32_t* __restrict__ mask,
                    const T* __restrict__ addend, T* __restrict__ out, int B, int N) {
  constexpr int TG = 32 / (4 * MP);           // graphs per warp tile
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[kAggWarps][kAggStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane & 3, mp = (lane >> 2) % MP, gl = lane / (4 * MP);
  const AggTileSizes ts = agg_tile_sizes<T>(N, TG, ADD);
  uint8_t* wbase = smem_raw + (size_t)warp * ts.warp_bytes;
  uint8_t* out_s = wbase + kAggStages * ts.stage_bytes;

  const int num_tiles = (B + TG - 1) / TG;
  const int warp_stride = gridDim.x * kAggWarps;
  int tile = blockIdx.x * kAggWarps + warp;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kAggStages; ++s) mbar_init(&bars[warp][s], 1);
    fence_mbar_init();
  }
  __syncwarp();

  const size_t graph_elems = (size_t)N * 16;

  auto issue = [&](int t, int s) {     // lane 0 only
    int ng = min(TG, B - t * TG);
    uint32_t hb = (uint32_t)(ng * N * 16 * sizeof(T));
    uint32_t mb = (uint32_t)(ng * N * 4);
    bool mask_bulk = (mb & 15u) == 0;
    uint8_t* st = wbase + s * ts.stage_bytes;
    uint32_t tx = hb + (ADD ? hb : 0) + (mask_bulk ? mb : 0);
    mbar_arrive_expect_tx(&bars[warp][s], tx);
    bulk_g2s(st, H + (size_t)t * TG * graph_elems, hb, &bars[warp][s]);
    if (ADD) bulk_g2s(st + ts.h_bytes, addend + (size_t)t * TG * graph_elems, hb, &bars[warp][s]);
    if (mask_bulk)
      bulk_g2s(st + ts.h_bytes * (ADD ? 2 : 1), mask + (size_t)t * TG * N, mb, &bars[warp][s]);
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kAggStages; ++s) {
      int t = tile + s * warp_stride;
      if (t < num_tiles) issue(t, s);
    }
  }

  uint32_t phase = 0;
  int stage = 0;
  for (; tile < num_tiles; tile += warp_stride) {
    const int ng = min(TG, B - tile * TG);
    uint8_t* st = wbase + stage * ts.stage_bytes;
    const T* Hs = reinterpret_cast<const T*>(st);
    const T* As = reinterpret_cast<const T*>(st + ts.h_bytes);
    uint32_t* Ms = reinterpret_cast<uint32_t*>(st + ts.h_bytes * (ADD ? 2 : 1));

    if (((ng * N * 4) & 15) != 0) {    // ragged last tile: masks by plain loads
      for (int i = lane; i < ng * N; i += 32) Ms[i] = mask[(size_t)tile * TG * N + i];
      __syncwarp();
    }
    mbar_wait(&bars[warp][stage], (phase >> stage) & 1u);

    uint32_t msk[MT];
    float4 acc[MT];
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
    // previous tile's store must have finished reading out_s
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    T* Os = reinterpret_cast<T*>(out_s);
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int m = j * MP + mp;
      if (m < N) st_row4(Os + ((size_t)(gl * N + m) * 16 + c * 4), acc[j]);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(out + (size_t)tile * TG * graph_elems, out_s, (uint32_t)(ng * N * 16 * sizeof(T)));
      bulk_commit();
      const int nt = tile + kAggStages * warp_stride;   // refill the stage just consumed
      if (nt < num_tiles) issue(nt, stage);
    }
    phase ^= (1u << stage);
    stage = (stage + 1 == kAggStages) ? 0 : stage + 1;
  }
  if (lane == 0) bulk_wait<0>();
}

// ---------------------------------------------------------------------------
// generic paths (any N <= 256, any F): one thread per output element
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void agg_mask_generic_kernel(const T* __restrict__ H, const uint32_t* __restrict__ mask,
                                        const T* __restrict__ addend, T* __restrict__ out,
                                        long total, int N, int F, int W) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = (int)(idx % F);
  const long bm = idx / F;              // b*N + m
  const long b = bm / N;
  const T* Hb = H + b * (long)N * F + f;
  float acc = addend ? to_f<T>(addend[idx]) : 0.f;
  for (int w = 0; w < W; ++w) {
    uint32_t bits = mask[bm * W + w];
    while (bits) {
      int n = w * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      acc += to_f<T>(Hb[(long)n * F]);
    }
  }
  out[idx] = from_f<T>(acc);
}

__global__ void agg_dense_kernel(const float* __restrict__ H, const float* __restrict__ adj,
                                 const float* __restrict__ addend, float* __restrict__ out,
                                 long total, int N, int F, int transpose) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = (int)(idx % F);
  const long bm = idx / F;
  const int m = (int)(bm % N);
  const long b = bm / N;
  const float* Hb = H + b * (long)N * F + f;
  const float* Ab = adj + b * (long)N * N;
  float acc = addend ? addend[idx] : 0.f;
  for (int n = 0; n < N; ++n) {
    float w = transpose ? Ab[(long)m * N + n] : Ab[(long)n * N + m];
    acc = fmaf(w, Hb[(long)n * F], acc);
  }
  out[idx] = acc;
}

__global__ void adj_pack_kernel(const float* __restrict__ adj, int B, int N, int W,
                                uint32_t* __restrict__ in_mask, uint32_t* __restrict__ out_mask,
                                int* __restrict__ nonbinary) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, i, w)
  long total = (long)B * N * W;
  if (idx >= total) return;
  const int w = (int)(idx % W);
  const long bi = idx / W;
  const int i = (int)(bi % N);
  const long b = bi / N;
  const float* Ab = adj + b * (long)N * N;
  uint32_t im = 0, om = 0;
  bool bad = false;
  for (int j = 0; j < 32; ++j) {
    int k = w * 32 + j;
    if (k >= N) break;
    float a_in = Ab[(long)k * N + i];    // Adj[n=k][m=i]
    float a_out = Ab[(long)i * N + k];   // Adj[n=i][m=k]
    if (a_in != 0.f) im |= 1u << j;
    if (a_out != 0.f) om |= 1u << j;
    bad |= (a_in != 0.f && a_in != 1.f);
  }
  if (in_mask) in_mask[idx] = im;
  if (out_mask) out_mask[idx] = om;
  if (bad && nonbinary) *nonbinary = 1;
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
template <typename T, int MT, int MP>
static int launch_fast(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N,
                       cudaStream_t st) {
  constexpr int TG = 32 / (4 * MP);
  const bool add = addend != nullptr;
  AggTileSizes ts = agg_tile_sizes<T>(N, TG, add);
  size_t smem = (size_t)ts.warp_bytes * kAggWarps;
  int num_tiles = ceil_div(B, TG);
  int grid = std::min(ceil_div(num_tiles, kAggWarps), sm_count());
  if (add) {
    auto k = agg_mask_f16_kernel<T, MT, MP, true>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
      V2V_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    k<<<grid, kAggWarps * 32, smem, st>>>(H, mask, addend, out, B, N);
  } else {
    auto k = agg_mask_f16_kernel<T, MT, MP, false>;
    static size_t smem_set = 0;
    if (smem > smem_set) {
      V2V_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    k<<<grid, kAggWarps * 32, smem, st>>>(H, mask, addend, out, B, N);
  }
  return launch_status("agg_mask_f16_kernel");
}

template <typename T>
static bool fast_fits(int N, int TG, bool add) {
  AggTileSizes ts = agg_tile_sizes<T>(N, TG, add);
  return (size_t)ts.warp_bytes * kAggWarps <= 226 * 1024;   // 227 KB minus the static mbarriers
}

template <typename T>
static int agg_mask_dispatch(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N,
                             int F, cudaStream_t st) {
  const bool aligned = ((((uintptr_t)H) | ((uintptr_t)out) | ((uintptr_t)mask) | ((uintptr_t)addend)) & 15u) == 0;
  const bool add = addend != nullptr;
  if (F == 16 && N <= 32 && aligned && B > 0) {
    if (N <= 8 && fast_fits<T>(N, 8, add)) return launch_fast<T, 8, 1>(H, mask, addend, out, B, N, st);
    if (N <= 20 && fast_fits<T>(N, 4, add)) return launch_fast<T, 10, 2>(H, mask, addend, out, B, N, st);
    if (fast_fits<T>(N, 4, add)) return launch_fast<T, 16, 2>(H, mask, addend, out, B, N, st);
    if (fast_fits<T>(N, 2, add)) return launch_fast<T, 8, 4>(H, mask, addend, out, B, N, st);
  }
  const int W = ceil_div(N, 32);
  long total = (long)B * N * F;
  if (total == 0) return 0;
  int threads = 256;
  long blocks = (total + threads - 1) / threads;
  agg_mask_generic_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(H, mask, addend, out, total, N, F, W);
  return launch_status("agg_mask_generic_kernel");
}

}  // namespace v2v

using namespace v2v;

extern "C" int v2v_agg_mask(const void* H_dev, const uint32_t* mask_dev, const void* addend_dev,
                            void* out_dev, int B, int N, int F, int dtype, void* stream) {
  V2V_REQUIRE(B >= 0 && N > 0 && N <= 256 && F > 0, "v2v_agg_mask: bad shape B=%d N=%d F=%d", B, N, F);
  if (B == 0) return 0;
  V2V_REQUIRE(H_dev && mask_dev && out_dev, "v2v_agg_mask: null pointer");
  V2V_REQUIRE(H_dev != out_dev, "v2v_agg_mask: out must not alias H");
  V2V_REQUIRE(dtype == V2V_F32 || dtype == V2V_BF16, "v2v_agg_mask: unknown dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == V2V_F32)
    return agg_mask_dispatch<float>((const float*)H_dev, mask_dev, (const float*)addend_dev, (float*)out_dev, B, N, F, st);
  if (dtype == V2V_BF16)
    return agg_mask_dispatch<__nv_bfloat16>((const __nv_bfloat16*)H_dev, mask_dev, (const __nv_bfloat16*)addend_dev,
                                            (__nv_bfloat16*)out_dev, B, N, F, st);
  return fail("v2v_agg_mask: unknown dtype %d", dtype);
}

extern "C" int v2v_agg_dense(const float* H_dev, const float* adj_dev, const float* addend_dev,
                             float* out_dev, int B, int N, int F, int transpose, void* stream) {
  V2V_REQUIRE(B >= 0 && N > 0 && F > 0, "v2v_agg_dense: bad shape B=%d N=%d F=%d", B, N, F);
  if (B == 0) return 0;
  V2V_REQUIRE(H_dev && adj_dev && out_dev, "v2v_agg_dense: null pointer");
  V2V_REQUIRE(H_dev != out_dev, "v2v_agg_dense: out must not alias H");
  long total = (long)B * N * F;
  int threads = 256;
  long blocks = (total + threads - 1) / threads;
  agg_dense_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(H_dev, adj_dev, addend_dev, out_dev, total, N, F,
                                                                          transpose);
  return launch_status("agg_dense_kernel");
}

extern "C" int v2v_adj_pack_masks(const float* adj_dev, int B, int N, uint32_t* in_mask_dev,
                                  uint32_t* out_mask_dev, int* nonbinary_flag_dev, void* stream) {
  V2V_REQUIRE(B >= 0 && N > 0 && N <= 256, "v2v_adj_pack_masks: bad shape B=%d N=%d", B, N);
  if (B == 0) return 0;
  V2V_REQUIRE(adj_dev, "v2v_adj_pack_masks: null adjacency");
  const int W = ceil_div(N, 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (nonbinary_flag_dev) V2V_CHECK_CUDA(cudaMemsetAsync(nonbinary_flag_dev, 0, sizeof(int), st));
  long total = (long)B * N * W;
  int threads = 256;
  long blocks = (total + threads - 1) / threads;
  adj_pack_kernel<<<(unsigned)blocks, threads, 0, st>>>(adj_dev, B, N, W, in_mask_dev, out_mask_dev, nonbinary_flag_dev);
  return launch_status("adj_pack_kernel");
}
