// Loss head, DQN target rule and optimiser of the brain.
//   huber_loss + compile   BS_brain.py:86-87, :212-214
//   TD target              BS_brain.py:668-692
//   Keras 2.2.4 Adam       BS_brain.py:212
#include "v2v_common.cuh"

namespace v2v {

// q, y, dq: [B][N][CH]; head_loss[N] += sum_b,a huber / (B*CH)
__global__ void huber_kernel(const float* __restrict__ q, const float* __restrict__ y,
                             float* __restrict__ dq, float* __restrict__ head_loss, long total,
                             int N, int CH, float inv_cnt, float grad_scale) {
  extern __shared__ float hl[];   // [N]
  for (int i = threadIdx.x; i < N; i += blockDim.x) hl[i] = 0.f;
  __syncthreads();
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const float e = q[idx] - y[idx];
    const float ae = fabsf(e);
    const float quad = fminf(ae, 1.f);
    const float lin = ae - quad;
    const float l = 0.5f * quad * quad + lin;
    dq[idx] = fminf(fmaxf(e, -1.f), 1.f) * inv_cnt * grad_scale;
    const int k = (int)((idx / CH) % N);
    atomicAdd(&hl[k], l);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (hl[i] != 0.f) atomicAdd(&head_loss[i], hl[i] * inv_cnt);
}

__global__ void td_target_kernel(const float* __restrict__ p, const float* __restrict__ pn,
                                 const int32_t* __restrict__ act, const float* __restrict__ rew,
                                 float gamma, float* __restrict__ y, long BN, int N, int CH) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, k)
  if (idx >= BN) return;
  const long b = idx / N;
  float mx = pn[idx * CH];
  for (int a = 1; a < CH; ++a) mx = fmaxf(mx, pn[idx * CH + a]);
  const int sel = act[idx];
  const float t = __fadd_rn(rew[b], __fmul_rn(gamma, mx));   // no FMA contraction: matches r + gamma*max exactly
  for (int a = 0; a < CH; ++a) y[idx * CH + a] = (a == sel) ? t : p[idx * CH + a];
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr_t, float b1, float b2, float eps,
                            float gscale) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace v2v

using namespace v2v;

extern "C" int v2v_huber_loss_grad(const float* q_dev, const float* y_dev, float* dq_dev,
                                   float* head_loss_dev, int B, int N, int CH, float grad_scale,
                                   void* stream) {
  V2V_REQUIRE(B >= 0 && N > 0 && CH > 0, "v2v_huber_loss_grad: bad shape B=%d N=%d CH=%d", B, N, CH);
  if (B == 0) return 0;
  V2V_REQUIRE(q_dev && y_dev && dq_dev && head_loss_dev, "v2v_huber_loss_grad: null pointer");
  long total = (long)B * N * CH;
  int threads = 256;
  int blocks = (int)std::min<long>((total + threads - 1) / threads, 4L * sm_count());
  huber_kernel<<<blocks, threads, N * sizeof(float), (cudaStream_t)stream>>>(
      q_dev, y_dev, dq_dev, head_loss_dev, total, N, CH, 1.f / ((float)B * (float)CH), grad_scale);
  return launch_status("huber_kernel");
}

extern "C" int v2v_td_target(const float* p_dev, const float* p_next_dev, const int32_t* action_dev,
                             const float* reward_dev, float gamma, float* y_dev, int B, int N, int CH,
                             void* stream) {
  V2V_REQUIRE(B >= 0 && N > 0 && CH > 0, "v2v_td_target: bad shape");
  if (B == 0) return 0;
  V2V_REQUIRE(p_dev && p_next_dev && action_dev && reward_dev && y_dev, "v2v_td_target: null pointer");
  long BN = (long)B * N;
  int threads = 256;
  td_target_kernel<<<(unsigned)((BN + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      p_dev, p_next_dev, action_dev, reward_dev, gamma, y_dev, BN, N, CH);
  return launch_status("td_target_kernel");
}

extern "C" int v2v_adam_step(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, long n,
                             int t, float lr, float beta1, float beta2, float eps, float grad_scale,
                             void* stream) {
  V2V_REQUIRE(n >= 0 && t >= 1, "v2v_adam_step: bad n=%ld t=%d", n, t);
  if (n == 0) return 0;
  V2V_REQUIRE(p_dev && g_dev && m_dev && v_dev, "v2v_adam_step: null pointer");
  // lr_t exactly as Keras computes it, in double then rounded once
  const double lr_t = (double)lr * (sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t)));
  int threads = 256;
  int blocks = (int)std::min<long>((n + threads - 1) / threads, 4L * sm_count());
  adam_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(p_dev, g_dev, m_dev, v_dev, n, (float)lr_t, beta1,
                                                            beta2, eps, grad_scale);
  return launch_status("adam_kernel");
}
