// Device-side pieces of the DQN loop that are pure launch overhead when written as a dozen tensor
// operations each -- one kernel per piece, so that a whole environment step of the batched agent is ~13 launches that a
// CUDA graph replays (dqn.BatchedAgent):
//   * epsilon-greedy action selection of Agent.select_action_while_training (BS_brain.py:308-352): epsilon from the
//     linear schedule (:315-324), per environment either a uniform random channel for every link (:330-333) or the FIRST
//     maximiser of every link's Q row (:336-344);
//   * Memory.add (BS_brain.py:252-256) for a device-resident ring: T transitions written at the ring's cursor, the cursor
//     (and the agent's step counter) advanced on the device.
#include "v2v_common.cuh"

namespace v2v {

// sched = {environment step, base, decrement per step, floor}: epsilon = max(floor, base - decrement * step)
__global__ void select_actions_kernel(const float* __restrict__ q, const float* __restrict__ u_explore,
                                      const int32_t* __restrict__ rnd, const float* __restrict__ sched,
                                      int32_t* __restrict__ actions, long EN, int N, int CH) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;       // (environment, link)
  if (idx >= EN) return;
  const float eps = fmaxf(sched[3], __fsub_rn(sched[1], __fmul_rn(sched[2], sched[0])));
  const long e = idx / N;
  int a;
  if (u_explore[e] < eps) {
    a = rnd[idx];
  } else {
    const float* row = q + idx * CH;
    a = 0;
    float best = row[0];
    for (int c = 1; c < CH; ++c)
      if (row[c] > best) { best = row[c]; a = c; }                     // strict: the first maximiser wins (:342-344)
  }
  actions[idx] = a;
}

struct RingPtrs {
  float *node, *edge, *node_, *edge_, *reward;
  int32_t *in_mask, *out_mask, *action;
};
struct RingSrc {
  const float *node, *edge, *node_, *edge_, *reward;
  const int32_t *in_mask, *out_mask, *action;
};

// slot(t) = (head + t) % capacity.  The block that finishes last advances the cursor (every block has read it by then).
__global__ void replay_write_kernel(RingPtrs dst, RingSrc src, long long* __restrict__ head, float* __restrict__ step,
                                    unsigned* __restrict__ done, int T, long capacity, int N, int Dn, int De, int W) {
  const long h = (long)*head;
  const int per = N * (2 * Dn + 2 * De + 2 * W + 1) + 1;               // 32-bit words of one transition
  const long total = (long)T * per;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / per);
    int w = (int)(i - (long)t * per);
    const long s = (h + t) % capacity;
    if (w < N * Dn) { dst.node[s * N * Dn + w] = src.node[(long)t * N * Dn + w]; continue; }
    w -= N * Dn;
    if (w < N * De) { dst.edge[s * N * De + w] = src.edge[(long)t * N * De + w]; continue; }
    w -= N * De;
    if (w < N * Dn) { dst.node_[s * N * Dn + w] = src.node_[(long)t * N * Dn + w]; continue; }
    w -= N * Dn;
    if (w < N * De) { dst.edge_[s * N * De + w] = src.edge_[(long)t * N * De + w]; continue; }
    w -= N * De;
    if (w < N * W) { dst.in_mask[s * N * W + w] = src.in_mask[(long)t * N * W + w]; continue; }
    w -= N * W;
    if (w < N * W) { dst.out_mask[s * N * W + w] = src.out_mask[(long)t * N * W + w]; continue; }
    w -= N * W;
    if (w < N) { dst.action[s * N + w] = src.action[(long)t * N + w]; continue; }
    dst.reward[s] = src.reward[t];
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *done = 0u;
    *head = (long long)((h + T) % capacity);
    if (step) *step += 1.f;
  }
}

}  // namespace v2v

using namespace v2v;

extern "C" int v2v_dqn_select_actions(const float* q_dev, const float* u_explore_dev, const int32_t* random_action_dev,
                                      const float* sched_dev, int32_t* action_dev, int E, int N, int CH, void* stream) {
  V2V_REQUIRE(E >= 0 && N > 0 && CH > 0, "v2v_dqn_select_actions: bad shape");
  if (E == 0) return 0;
  V2V_REQUIRE(q_dev && u_explore_dev && random_action_dev && sched_dev && action_dev, "v2v_dqn_select_actions: null pointer");
  const long EN = (long)E * N;
  select_actions_kernel<<<(unsigned)((EN + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q_dev, u_explore_dev, random_action_dev,
                                                                                         sched_dev, action_dev, EN, N, CH);
  return launch_status("select_actions_kernel");
}

extern "C" int v2v_dqn_replay_write(float* ring_node, float* ring_edge, float* ring_node_next, float* ring_edge_next,
                                    int32_t* ring_in_mask, int32_t* ring_out_mask, int32_t* ring_action, float* ring_reward,
                                    const float* node, const float* edge, const float* node_next, const float* edge_next,
                                    const int32_t* in_mask, const int32_t* out_mask, const int32_t* action, const float* reward,
                                    long long* head_dev, float* step_dev, unsigned* done_dev, int T, long capacity, int N, int Dn,
                                    int De, void* stream) {
  V2V_REQUIRE(T >= 0 && capacity > 0 && T <= capacity && N > 0 && Dn > 0 && De > 0, "v2v_dqn_replay_write: bad shape");
  if (T == 0) return 0;
  V2V_REQUIRE(ring_node && ring_edge && ring_node_next && ring_edge_next && ring_in_mask && ring_out_mask && ring_action &&
              ring_reward && node && edge && node_next && edge_next && in_mask && out_mask && action && reward && head_dev &&
              done_dev, "v2v_dqn_replay_write: null pointer");
  const int W = (N + 31) / 32;
  const long total = (long)T * (N * (2 * Dn + 2 * De + 2 * W + 1) + 1);
  const int blocks = (int)std::max<long>(1, std::min<long>((total + 255) / 256, 4L * sm_count()));
  RingPtrs d{ring_node, ring_edge, ring_node_next, ring_edge_next, ring_reward, ring_in_mask, ring_out_mask, ring_action};
  RingSrc s{node, edge, node_next, edge_next, reward, in_mask, out_mask, action};
  replay_write_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d, s, head_dev, step_dev, done_dev, T, capacity, N, Dn, De, W);
  return launch_status("replay_write_kernel");
}
