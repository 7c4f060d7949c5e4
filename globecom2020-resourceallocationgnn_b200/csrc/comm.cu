// Gradient all-reduce fused with the optimiser, over NVLink peer memory (one process per GPU).
//
// The data-parallel brain exchanges ONE flat fp32 gradient per step (8,720 floats = 35 KB with shared weights):
// purely latency-bound, so instead of a library collective followed by an optimiser kernel, ONE kernel
//   1. reduces this rank's per-CTA gradient partials (fused_brain_kernel output), 64 columns per CTA, the partials
//      split over 8 slices so that every load is independent,
//   2. pushes each reduced element straight into its slot of every peer's communication buffer as ONE 8-byte store
//      {value, epoch} to a cudaIpc-mapped peer pointer (NVLink 5 / NVSwitch).  The epoch travels inside the same
//      atomic 8-byte transaction as the value, so no fence, no separate flag and no second NVLink round trip is
//      needed (the "low-latency" protocol of collective libraries),
//   3. polls its own slots until every rank's element carries the current epoch, sums them in rank order
//      (bit-identical result on every rank), scales by 1/world and applies the Keras-Adam update
//      (BS_brain.py:212) in place.
// Elements are independent: a CTA only waits for its own 64 columns, so the transfer of one chunk overlaps the
// math of another and there is no grid-wide barrier.  Slots are double-buffered by epoch parity: a rank can run
// at most one epoch ahead of its slowest peer because it needs that peer's values to finish its own epoch.
#include "v2v_common.cuh"

namespace v2v {

constexpr int kCommChunk = 64;           // floats per CTA
constexpr int kCommSlices = 8;           // warps-pairs splitting the sum over the per-CTA partials
constexpr int kCommThreads = kCommChunk * kCommSlices;
constexpr int kCommMaxWorld = 16;

struct CommPeers {
  uint2* slots[kCommMaxWorld];           // peer r's communication buffer: slots[2][world][n_pad] of {value bits, epoch}
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// one 8-byte transaction: the epoch can never be observed without its value
__device__ __forceinline__ void st_ll(uint2* p, float val, uint32_t epoch) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(val)), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kCommThreads)
allreduce_adam_kernel(const float* __restrict__ partial, int n_cta, long row_stride,  // local partials [n_cta][row_stride]
                      long n_adam, long n_src,        // payload columns 0..n_src-1 of the rows; the first n_adam get Adam
                      const float* __restrict__ extra, int n_extra,                   // appended payload
                      CommPeers peers, int world, int rank, uint32_t epoch, int n_pad,
                      float* __restrict__ grad, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ extra_out, float lr_t, float b1, float b2, float eps, int* __restrict__ error,
                      unsigned long long* __restrict__ trace) {
  __shared__ float red[kCommSlices][kCommChunk];
  const int chunk = blockIdx.x;
  const int col = threadIdx.x & (kCommChunk - 1), slice = threadIdx.x / kCommChunk;
  const int i = chunk * kCommChunk + col;                         // element of the padded payload
  const int buf = epoch & 1u;
  const bool tr = trace != nullptr && threadIdx.x == 0;          // optional phase trace: globaltimer stamps per chunk
  if (tr) trace[chunk * 6 + 0] = globaltimer_ns();
  asm volatile("griddepcontrol.wait;" ::: "memory");             // the producer of `partial` (programmatic dependent launch)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // every CTA is running: the next step may queue up
  if (tr) trace[chunk * 6 + 1] = globaltimer_ns();
  // 1. local reduction of this element: the partials are split over the slices, every load independent
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
  if (i < n_src) {
    const float* src = partial + i;
    int c = slice;
    for (; c + 3 * kCommSlices < n_cta; c += 4 * kCommSlices) {
      const float a0 = __ldcg(src + (long)c * row_stride), a1 = __ldcg(src + (long)(c + kCommSlices) * row_stride);
      const float a2 = __ldcg(src + (long)(c + 2 * kCommSlices) * row_stride), a3 = __ldcg(src + (long)(c + 3 * kCommSlices) * row_stride);
      g0 += a0; g1 += a1; g2 += a2; g3 += a3;
    }
    for (; c < n_cta; c += kCommSlices) g0 += __ldcg(src + (long)c * row_stride);
  }
  red[slice][col] = (g0 + g1) + (g2 + g3);
  __syncthreads();
  if (slice != 0) return;                                         // the first kCommChunk threads own one element each
  float g = 0.f;
  if (i < n_src) {
#pragma unroll
    for (int s = 0; s < kCommSlices; ++s) g += red[s][col];
  } else if (i < n_src + n_extra) {
    g = extra[i - n_src];
  }
  // 2. push {value, epoch} to every peer's slot for this rank
  const size_t slot_off = ((size_t)buf * world + rank) * n_pad + i;
  for (int r = 0; r < world; ++r)
    if (r != rank) st_ll(peers.slots[r] + slot_off, g, epoch);
  if (tr) trace[chunk * 6 + 2] = globaltimer_ns();
  // 3. gather every rank's element from our own slots.  All peers are polled in ONE pass per spin (the loads of a
  //    pass are independent, so a late rank does not serialise the others' round trips); the values are then summed in
  //    rank order (identical on every rank).
  const uint2* mine = peers.slots[rank] + (size_t)buf * world * n_pad + i;
  float vals[kCommMaxWorld];
  uint32_t pending = (world >= 32 ? 0xffffffffu : ((1u << world) - 1u)) & ~(1u << rank);
  long spins = 0;
  bool timed_out = false;
  while (pending) {
#pragma unroll
    for (int r = 0; r < kCommMaxWorld; ++r) {
      if (pending & (1u << r)) {
        const uint2 w = ld_ll(mine + (size_t)r * n_pad);
        if (w.y == epoch) { vals[r] = __uint_as_float(w.x); pending &= ~(1u << r); }
      }
    }
    if (pending && ++spins > (1L << 22)) { timed_out = true; break; }   // ~seconds: a peer is gone; never hang the GPU
  }
  if (timed_out) {               // never train on stale or missing peer data: flag it and leave p / m / v untouched
    *error = 1;
    return;
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < kCommMaxWorld; ++r)
    if (r < world) s += (r == rank) ? g : vals[r];
  if (tr) trace[chunk * 6 + 3] = globaltimer_ns();
  const float inv = 1.f / (float)world;
  if (i < n_adam) {
    const float gi = s * inv;
    grad[i] = gi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  } else if (i < n_src + n_extra && extra_out) {
    extra_out[i - n_adam] = s * inv;
  }
  if (tr) trace[chunk * 6 + 4] = globaltimer_ns();
}

}  // namespace v2v

using namespace v2v;

struct v2v_comm {
  int world = 1, rank = 0;
  long n = 0;                // payload floats (parameters + extra)
  int n_pad = 0, n_chunks = 0;
  uint32_t epoch = 0;
  uint2* slots = nullptr;    // local buffer: slots[2][world][n_pad] of {value bits, epoch}
  size_t bytes = 0;
  CommPeers peers{};
  void* opened[kCommMaxWorld] = {};
  int* error_dev = nullptr;
  int* error_host = nullptr;  // pinned mirror of error_dev, refreshed by v2v_comm_poll_error
  unsigned long long* trace_dev = nullptr;   // optional: [n_chunks][6] globaltimer stamps of the last exchange
  bool peers_ready = false;
};

extern "C" int v2v_comm_create(long n_floats, int world, int rank, v2v_comm** out) {
  V2V_REQUIRE(out && n_floats > 0 && world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world,
              "v2v_comm_create: bad arguments (world <= %d)", kCommMaxWorld);
  v2v_comm* c = new v2v_comm();
  c->world = world; c->rank = rank; c->n = n_floats;
  c->n_chunks = (int)((n_floats + kCommChunk - 1) / kCommChunk);
  c->n_pad = c->n_chunks * kCommChunk;
  c->bytes = (size_t)2 * world * c->n_pad * sizeof(uint2);
  if (cudaMalloc((void**)&c->slots, c->bytes) != cudaSuccess || cudaMalloc((void**)&c->error_dev, sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void**)&c->error_host, sizeof(int)) != cudaSuccess) {
    delete c;
    return fail("v2v_comm_create: cudaMalloc failed");
  }
  cudaMemset(c->slots, 0, c->bytes);
  cudaMemset(c->error_dev, 0, sizeof(int));
  *c->error_host = 0;
  c->peers.slots[rank] = c->slots;
  c->peers_ready = (world == 1);
  cudaDeviceSynchronize();
  *out = c;
  return 0;
}

extern "C" void v2v_comm_destroy(v2v_comm* c) {
  if (!c) return;
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
  cudaFree(c->slots);
  cudaFree(c->error_dev);
  if (c->error_host) cudaFreeHost(c->error_host);
  delete c;
}

extern "C" int v2v_comm_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int v2v_comm_get_ipc_handle(v2v_comm* c, void* handle_out) {
  V2V_REQUIRE(c && handle_out, "v2v_comm_get_ipc_handle: null argument");
  cudaIpcMemHandle_t h;
  V2V_CHECK_CUDA(cudaIpcGetMemHandle(&h, c->slots));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

// handles: world consecutive cudaIpcMemHandle_t blobs, in rank order (ours is ignored)
extern "C" int v2v_comm_open_peers(v2v_comm* c, const void* handles) {
  V2V_REQUIRE(c && handles, "v2v_comm_open_peers: null argument");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const uint8_t*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* base = nullptr;
    V2V_CHECK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c->opened[r] = base;
    c->peers.slots[r] = reinterpret_cast<uint2*>(base);
  }
  c->peers_ready = true;
  return 0;
}

// Reduce partial[n_cta][n_src] (+ extra[n_extra]) over ranks and apply Keras-Adam to p/m/v (n_src floats).
extern "C" int v2v_comm_allreduce_adam_ex(v2v_comm* c, const float* partial_dev, int n_cta, long row_stride, long n_adam,
                                          long n_src, const float* extra_dev, int n_extra, float* grad_dev, float* p_dev,
                                          float* m_dev, float* v_dev, float* extra_out_dev, int t, float lr, float beta1,
                                          float beta2, float eps, void* stream);
extern "C" int v2v_comm_allreduce_adam(v2v_comm* c, const float* partial_dev, int n_cta, long n_src,
                                       const float* extra_dev, int n_extra, float* grad_dev, float* p_dev,
                                       float* m_dev, float* v_dev, float* extra_out_dev, int t, float lr, float beta1,
                                       float beta2, float eps, void* stream) {
  return v2v_comm_allreduce_adam_ex(c, partial_dev, n_cta, n_src, n_src, n_src, extra_dev, n_extra, grad_dev, p_dev, m_dev,
                                    v_dev, extra_out_dev, t, lr, beta1, beta2, eps, stream);
}

// General form: the rows of partial_dev are row_stride floats apart and carry n_src payload columns; the first n_adam
// columns are parameters (gradient -> grad_dev, Keras-Adam on p/m/v), the remaining n_src - n_adam columns followed by
// the n_extra floats of extra_dev are only averaged and land in extra_out_dev.
extern "C" int v2v_comm_allreduce_adam_ex(v2v_comm* c, const float* partial_dev, int n_cta, long row_stride, long n_adam,
                                          long n_src, const float* extra_dev, int n_extra, float* grad_dev, float* p_dev,
                                          float* m_dev, float* v_dev, float* extra_out_dev, int t, float lr, float beta1,
                                          float beta2, float eps, void* stream) {
  V2V_REQUIRE(c && c->peers_ready, "v2v_comm_allreduce_adam: peers not opened");
  V2V_REQUIRE(partial_dev && grad_dev && p_dev && m_dev && v_dev && n_cta >= 1 && t >= 1, "v2v_comm_allreduce_adam: bad arguments");
  V2V_REQUIRE(n_adam >= 0 && n_adam <= n_src && n_src <= row_stride && (n_extra == 0 || extra_dev),
              "v2v_comm_allreduce_adam: bad layout (n_adam %ld, n_src %ld, row_stride %ld)", n_adam, n_src, row_stride);
  V2V_REQUIRE(n_src + n_extra <= c->n, "v2v_comm_allreduce_adam: payload %ld exceeds the communicator's %ld floats",
              n_src + n_extra, c->n);
  c->epoch += 1;
  const double lr_t = (double)lr * (sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t)));
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3(c->n_chunks);
  lc.blockDim = dim3(kCommThreads);
  lc.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  V2V_CHECK_CUDA(cudaLaunchKernelEx(&lc, allreduce_adam_kernel, partial_dev, n_cta, row_stride, n_adam, n_src, extra_dev, n_extra, c->peers,
                                    c->world, c->rank, c->epoch, c->n_pad, grad_dev, p_dev, m_dev, v_dev,
                                    extra_out_dev, (float)lr_t, beta1, beta2, eps, c->error_dev, c->trace_dev));
  return launch_status("allreduce_adam_kernel");
}

// trace_dev: device buffer of n_chunks * 6 uint64 (or null to disable); stamps per chunk: entry, producer done, pushed,
// all ranks arrived, Adam done (globaltimer ns)
extern "C" int v2v_comm_set_trace(v2v_comm* c, unsigned long long* trace_dev) {
  V2V_REQUIRE(c, "v2v_comm_set_trace: null argument");
  c->trace_dev = trace_dev;
  return 0;
}
extern "C" int v2v_comm_num_chunks(v2v_comm* c) { return c ? c->n_chunks : 0; }

// Stream-ordered, non-blocking error surfacing for the product path: enqueues a copy of the device error flag into a
// pinned mirror and fails if an EARLIER copy already reported a timeout (a rank that timed out skipped its update, so
// the replicas have diverged and training must stop).  Callers that synchronise the stream afterwards (the host
// entry points) call v2v_comm_poll_result to read the value that just landed.
extern "C" int v2v_comm_poll_error(v2v_comm* c, void* stream) {
  V2V_REQUIRE(c, "v2v_comm_poll_error: null argument");
  V2V_REQUIRE(*c->error_host == 0, "v2v_comm: a peer did not arrive at the gradient exchange (timeout); the update was skipped");
  V2V_CHECK_CUDA(cudaMemcpyAsync(c->error_host, c->error_dev, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
extern "C" int v2v_comm_poll_result(v2v_comm* c) {
  V2V_REQUIRE(c, "v2v_comm_poll_result: null argument");
  V2V_REQUIRE(*c->error_host == 0, "v2v_comm: a peer did not arrive at the gradient exchange (timeout); the update was skipped");
  return 0;
}

// non-zero if a wait timed out (a peer never arrived); synchronises the stream
extern "C" int v2v_comm_check(v2v_comm* c, void* stream) {
  V2V_REQUIRE(c, "v2v_comm_check: null argument");
  int e = 0;
  V2V_CHECK_CUDA(cudaMemcpyAsync(&e, c->error_dev, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  V2V_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  V2V_REQUIRE(e == 0, "v2v_comm: a peer did not arrive at the gradient exchange (timeout)");
  return 0;
}
