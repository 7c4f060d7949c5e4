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
#include <type_traits>

#include "agg_kernels.cuh"

namespace v2v {

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
constexpr int kAggWarps = 8;        // warps per CTA (2 CTAs per SM when shared memory allows)

static int agg_block_min_n() {      // tunable for experiments: V2V_AGG_BLOCK_MIN_N
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("V2V_AGG_BLOCK_MIN_N");
    v = e ? atoi(e) : 96;          // measured: warp-per-graph wins up to N = 64, CTA-per-graph from N = 128
  }
  return v;
}

// 8 < N <= 20: dense predicated gather-reduce, or the set-bit / clear-bit walk against the graph's column total.
// Measured (profiles/agg_variants_r02.txt): with fp32 storage the kernel is bandwidth-bound and the walk buys nothing (it
// lengthens the per-tile latency and costs ~1e-6 of parity), with bf16 storage the bytes halve, the predicated adds bind
// (0.49 of the roofline) and the walk's 24 % fewer instructions pay; the stored result is rounded to bf16 either way.
// V2V_AGG_WALK=0 / 1 forces one form for both storage types (experiments).
template <typename T>
static bool agg_use_walk() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("V2V_AGG_WALK");
    v = e ? atoi(e) : -1;
  }
  if (v >= 0) return v != 0;
  return std::is_same<T, __nv_bfloat16>::value;
}

template <typename T, int MT, int MP, int NC = 0, bool WALK = false>
static int launch_fast(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N, bool independent,
                       cudaStream_t st) {
  constexpr int TG = 32 / (4 * MP);
  AggLaunchCfg cfg;
  cfg.dep_wait = !independent;
  // CTAs per SM (measured at N = 20, profiles/agg_variants_r02.txt).  A launch that waits for its predecessor wants every
  // byte in flight at once: as many CTAs as fit, up to 4 (bf16 storage: 0.229 -> 0.336 of the roofline at B = 8192, 0.425 ->
  // 0.519 at 32768; fp32 +2 %).  A stream of independent launches wants 2 -- and with bf16 storage 1 while a warp has at
  // most ~8 tiles, so that the next launch's CTAs do not queue behind a second wave (0.626 -> 0.693 at B = 8192).
  const bool add = addend != nullptr;
  int want = 2;
  if (!independent) want = 4;
  else if (sizeof(T) == 2 && ceil_div(B, TG) <= 8 * kAggWarps * sm_count()) want = 1;
  while (want > 1 && !agg_fast_fits<T>(N, TG, add, kAggWarps, want)) --want;
  cfg.ctas_per_sm = want;
  if (addend) return launch_agg_fast<T, MT, MP, true, kAggWarps, true, NC, WALK>(H, mask, addend, out, B, N, cfg, st);
  return launch_agg_fast<T, MT, MP, false, kAggWarps, true, NC, WALK>(H, mask, addend, out, B, N, cfg, st);
}

template <typename T>
static int agg_mask_dispatch(const T* H, const uint32_t* mask, const T* addend, T* out, int B, int N,
                             int F, bool independent, cudaStream_t st) {
  const bool aligned = ((((uintptr_t)H) | ((uintptr_t)out) | ((uintptr_t)mask) | ((uintptr_t)addend)) & 15u) == 0;
  const bool add = addend != nullptr;
  if (F == 16 && N <= 20 && aligned && B > 0) {
    // small graphs: dense predicated gather-reduce (profiles/agg_variants_r01.txt)
    if (N <= 8 && agg_fast_fits<T>(N, 8, add, kAggWarps, 1)) return launch_fast<T, 8, 1>(H, mask, addend, out, B, N, independent, st);
    const bool walk = agg_use_walk<T>();
    if (N == 20 && agg_fast_fits<T>(N, 2, add, kAggWarps, 1))       // the north-star shape: compile-time N
      return walk ? launch_fast<T, 5, 4, 20, true>(H, mask, addend, out, B, N, independent, st)
                  : launch_fast<T, 5, 4, 20>(H, mask, addend, out, B, N, independent, st);
    if (agg_fast_fits<T>(N, 2, add, kAggWarps, 1))
      return walk ? launch_fast<T, 5, 4, 0, true>(H, mask, addend, out, B, N, independent, st)
                  : launch_fast<T, 5, 4>(H, mask, addend, out, B, N, independent, st);
  }
  if (F == 16 && N <= 256 && aligned && B > 0) {
    // larger graphs: set-bit / clear-bit walk (work ~ N * min(deg, N - deg)); one warp per graph tile up to
    // kAggBlockMinN nodes, one CTA per graph beyond (addend aliasing out is fine: same thread reads then writes)
    int rc = (N >= agg_block_min_n()) ? launch_agg_block<T>(H, mask, addend, out, B, N, !independent, st)
                                      : launch_agg_sparse<T>(H, mask, addend, out, B, N, !independent, st);
    if (rc < 0 && N < agg_block_min_n()) rc = launch_agg_block<T>(H, mask, addend, out, B, N, !independent, st);
    if (rc >= 0) return rc;
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

extern "C" int v2v_agg_mask_ex(const void* H_dev, const uint32_t* mask_dev, const void* addend_dev,
                               void* out_dev, int B, int N, int F, int dtype, unsigned flags, void* stream) {
  V2V_REQUIRE(B >= 0 && N > 0 && N <= 256 && F > 0, "v2v_agg_mask: bad shape B=%d N=%d F=%d", B, N, F);
  const bool independent = (flags & V2V_AGG_INDEPENDENT) != 0;
  if (B == 0) return 0;
  V2V_REQUIRE(H_dev && mask_dev && out_dev, "v2v_agg_mask: null pointer");
  V2V_REQUIRE(H_dev != out_dev, "v2v_agg_mask: out must not alias H");
  V2V_REQUIRE(dtype == V2V_F32 || dtype == V2V_BF16, "v2v_agg_mask: unknown dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == V2V_F32)
    return agg_mask_dispatch<float>((const float*)H_dev, mask_dev, (const float*)addend_dev, (float*)out_dev, B, N, F, independent, st);
  if (dtype == V2V_BF16)
    return agg_mask_dispatch<__nv_bfloat16>((const __nv_bfloat16*)H_dev, mask_dev, (const __nv_bfloat16*)addend_dev,
                                            (__nv_bfloat16*)out_dev, B, N, F, independent, st);
  return fail("v2v_agg_mask: unknown dtype %d", dtype);
}

extern "C" int v2v_agg_mask(const void* H_dev, const uint32_t* mask_dev, const void* addend_dev,
                            void* out_dev, int B, int N, int F, int dtype, void* stream) {
  return v2v_agg_mask_ex(H_dev, mask_dev, addend_dev, out_dev, B, N, F, dtype, 0u, stream);
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
