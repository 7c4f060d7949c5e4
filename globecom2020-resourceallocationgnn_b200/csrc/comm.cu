// Gradient all-reduce fused with the optimiser, over NVLink peer memory (one process per GPU).
//
// The data-parallel brain exchanges ONE flat fp32 gradient per step (8,720 floats = 35 KB with shared weights):
// purely latency-bound, so instead of a library collective followed by an optimiser kernel, ONE kernel
//   1. reduces this rank's per-CTA gradient partials (fused_brain_kernel output) chunk by chunk,
//   2. pushes each reduced chunk straight into a slot of every peer's communication buffer (plain stores to
//      cudaIpc-mapped peer pointers: they travel over NVLink 5 / NVSwitch), fences system-wide and raises a
//      per-(rank, chunk) epoch flag on every peer,
//   3. waits for the same chunk of all ranks to land locally, sums the slots in rank order (bit-identical
//      result on every rank), scales by 1/world and applies the Keras-Adam update (BS_brain.py:212) in place.
// Chunks are independent: CTA c only ever waits for chunk c, so the transfer of one chunk overlaps the math of
// another and there is no grid-wide barrier.  Slots are double-buffered by epoch parity: a rank can run at most
// one epoch ahead of its slowest peer because it needs that peer's flag to finish its own epoch.
#include "v2v_common.cuh"

namespace v2v {

constexpr int kCommChunk = 256;          // floats per CTA
constexpr int kCommMaxWorld = 16;

struct CommPeers {
  float* slots[kCommMaxWorld];           // peer r's communication buffer: slots[2][world][n_pad]
  uint32_t* flags[kCommMaxWorld];        // peer r's flags[world][n_chunks]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kCommChunk)
allreduce_adam_kernel(const float* __restrict__ partial, int n_cta, long n_src,      // local partials [n_cta][n_src]
                      const float* __restrict__ extra, int n_extra,                   // appended payload (head losses)
                      CommPeers peers, int world, int rank, uint32_t epoch, int n_pad, int n_chunks,
                      float* __restrict__ grad, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ extra_out, float lr_t, float b1, float b2, float eps, int* __restrict__ error) {
  const int chunk = blockIdx.x;
  const int i = chunk * kCommChunk + threadIdx.x;                 // element of the padded payload
  const int buf = epoch & 1u;
  // 1. local reduction of this element
  float g = 0.f;
  if (i < n_src) {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
    int c = 0;
    for (; c + 4 <= n_cta; c += 4) {
      g0 += partial[(long)c * n_src + i]; g1 += partial[(long)(c + 1) * n_src + i];
      g2 += partial[(long)(c + 2) * n_src + i]; g3 += partial[(long)(c + 3) * n_src + i];
    }
    for (; c < n_cta; ++c) g0 += partial[(long)c * n_src + i];
    g = (g0 + g1) + (g2 + g3);
  } else if (i < n_src + n_extra) {
    g = extra[i - n_src];
  }
  // 2. push to every rank's slot (including our own), fence, raise the flags
  const size_t slot_off = ((size_t)buf * world + rank) * n_pad + i;
  for (int r = 0; r < world; ++r) peers.slots[r][slot_off] = g;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + (size_t)rank * n_chunks + chunk, epoch);
  // 3. wait for this chunk of every rank
  if (threadIdx.x < world) {
    const uint32_t* f = peers.flags[rank] + (size_t)threadIdx.x * n_chunks + chunk;
    long spins = 0;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (++spins > (1L << 24)) { *error = 1; break; }           // ~seconds: a peer is gone; never hang the GPU
      __nanosleep(64);
    }
  }
  __syncthreads();
  const float* mine = peers.slots[rank] + (size_t)buf * world * n_pad + i;
  float s = 0.f;
  for (int r = 0; r < world; ++r) s += __ldcg(mine + (size_t)r * n_pad);   // rank order: identical on every rank
  const float inv = 1.f / (float)world;
  if (i < n_src) {
    const float gi = s * inv;
    grad[i] = gi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  } else if (i < n_src + n_extra && extra_out) {
    extra_out[i - n_src] = s * inv;
  }
}

}  // namespace v2v

using namespace v2v;

struct v2v_comm {
  int world = 1, rank = 0;
  long n = 0;                // payload floats (parameters + extra)
  int n_pad = 0, n_chunks = 0;
  uint32_t epoch = 0;
  float* slots = nullptr;    // local buffer: slots[2][world][n_pad] followed by flags[world][n_chunks]
  uint32_t* flags = nullptr;
  size_t bytes = 0;
  CommPeers peers{};
  void* opened[kCommMaxWorld] = {};
  int* error_dev = nullptr;
  bool peers_ready = false;
};

extern "C" int v2v_comm_create(long n_floats, int world, int rank, v2v_comm** out) {
  V2V_REQUIRE(out && n_floats > 0 && world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world,
              "v2v_comm_create: bad arguments (world <= %d)", kCommMaxWorld);
  v2v_comm* c = new v2v_comm();
  c->world = world; c->rank = rank; c->n = n_floats;
  c->n_chunks = (int)((n_floats + kCommChunk - 1) / kCommChunk);
  c->n_pad = c->n_chunks * kCommChunk;
  const size_t slot_bytes = (size_t)2 * world * c->n_pad * sizeof(float);
  const size_t flag_bytes = (size_t)world * c->n_chunks * sizeof(uint32_t);
  c->bytes = slot_bytes + flag_bytes;
  if (cudaMalloc((void**)&c->slots, c->bytes) != cudaSuccess || cudaMalloc((void**)&c->error_dev, sizeof(int)) != cudaSuccess) {
    delete c;
    return fail("v2v_comm_create: cudaMalloc failed");
  }
  cudaMemset(c->slots, 0, c->bytes);
  cudaMemset(c->error_dev, 0, sizeof(int));
  c->flags = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(c->slots) + slot_bytes);
  c->peers.slots[rank] = c->slots;
  c->peers.flags[rank] = c->flags;
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
  const size_t slot_bytes = (size_t)2 * c->world * c->n_pad * sizeof(float);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const uint8_t*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* base = nullptr;
    V2V_CHECK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c->opened[r] = base;
    c->peers.slots[r] = reinterpret_cast<float*>(base);
    c->peers.flags[r] = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(base) + slot_bytes);
  }
  c->peers_ready = true;
  return 0;
}

// Reduce partial[n_cta][n_src] (+ extra[n_extra]) over ranks and apply Keras-Adam to p/m/v (n_src floats).
extern "C" int v2v_comm_allreduce_adam(v2v_comm* c, const float* partial_dev, int n_cta, long n_src,
                                       const float* extra_dev, int n_extra, float* grad_dev, float* p_dev,
                                       float* m_dev, float* v_dev, float* extra_out_dev, int t, float lr, float beta1,
                                       float beta2, float eps, void* stream) {
  V2V_REQUIRE(c && c->peers_ready, "v2v_comm_allreduce_adam: peers not opened");
  V2V_REQUIRE(partial_dev && grad_dev && p_dev && m_dev && v_dev && n_cta >= 1 && t >= 1, "v2v_comm_allreduce_adam: bad arguments");
  V2V_REQUIRE(n_src + n_extra <= c->n, "v2v_comm_allreduce_adam: payload %ld exceeds the communicator's %ld floats",
              n_src + n_extra, c->n);
  c->epoch += 1;
  const double lr_t = (double)lr * (sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t)));
  allreduce_adam_kernel<<<c->n_chunks, kCommChunk, 0, (cudaStream_t)stream>>>(
      partial_dev, n_cta, n_src, extra_dev, n_extra, c->peers, c->world, c->rank, c->epoch, c->n_pad, c->n_chunks, grad_dev,
      p_dev, m_dev, v_dev, extra_out_dev, (float)lr_t, beta1, beta2, eps, c->error_dev);
  return launch_status("allreduce_adam_kernel");
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
