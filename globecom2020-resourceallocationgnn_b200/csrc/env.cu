// Batched V2V environment: E independent copies of the reference simulator stepped on the device.
//
// The reference simulates ONE scenario in Python loops (Environment.py: 0.8 ms per step at N = 4, 9.7 ms at N = 20,
// mostly element-wise random.gauss fills and O(N^2 RB) loops); the large-batch training configurations need thousands
// of graphs per step.  These kernels restate the per-step arithmetic for E environments at once, state resident in HBM:
//   v2v_env_renew_channels   renew_channel + renew_channels_fastfading   (Environment.py:63-120, :140-165, :378-404)
//   v2v_env_pack_state       Agent.get_state + packing + adjacency       (BS_brain.py:389-407, :441-469)
//   v2v_env_reward           compute_reward_with_channel_selection       (Environment.py:406-458)
//   v2v_env_renew_positions  renew_positions                             (Environment.py:236-345)
//   v2v_env_choose_destinations  renew_neighbor                          (Environment.py:360-376)
// Randomness is injected (arrays of normal / uniform draws), exactly as in oracle/env_oracle.py, so every kernel is
// checked against golden vectors recorded from the unmodified reference.  All kernels are element-wise / small-reduction
// HBM streamers: one thread per (environment, vehicle[, vehicle]) item, coalesced 128-bit accesses on the RB = 4 rows.
#include <math.h>

#include "v2v_common.cuh"

namespace v2v {

namespace {

__constant__ float c_up[6] = {1.75f, 5.25f, 251.75f, 255.25f, 501.75f, 505.25f};
__constant__ float c_down[6] = {244.75f, 248.25f, 494.75f, 498.25f, 744.75f, 748.25f};
__constant__ float c_left[6] = {1.75f, 5.25f, 434.75f, 438.25f, 867.75f, 871.25f};
__constant__ float c_right[6] = {427.75f, 431.25f, 860.75f, 864.25f, 1293.75f, 1297.25f};
constexpr float kWidth = 750.f, kHeight = 1299.f, kTimestep = 0.01f;
constexpr float kV2VPower = 10.f;            // V2V_power_dB_List[fixed_v2v_power_index] (Environment.py:194-195)
constexpr float kV2IPower = 23.f;            // :193
constexpr float kSig2 = 3.98107170553497e-12f;   // 10^(-114/10) (:196, :201)
constexpr float kBsAnt = 8.f, kBsNF = 5.f, kVehAnt = 3.f, kVehNF = 9.f;

__device__ __forceinline__ float db2lin(float x) { return exp10f(x * 0.1f); }
// log10 through the SFU: lg2.approx is accurate to ~2^-22 absolute here (arguments are distances and |h|^2, never
// denormal), i.e. < 1e-5 dB after the 10..40 x scaling -- far inside the 2e-4 dB the parity tests allow
__device__ __forceinline__ float flog10(float x) { return __log2f(x) * 0.30102999566398120f; }

__device__ __forceinline__ float pl_los(float d) {                  // Environment.py:99-107 with h_bs = h_ms = 1.5, fc = 2
  const float c0 = 41.f + 20.f * flog10(2.f / 5.f);
  if (d <= 3.f) return 22.7f * flog10(3.f) + c0;
  const float d_bp = 4.f * 0.5f * 0.5f * 2.f * 1e9f / 3e8f;
  if (d < d_bp) return 22.7f * flog10(d) + c0;
  return 40.f * flog10(d) + 9.45f - 2.f * 17.3f * flog10(1.5f) + 2.7f * flog10(2.f / 5.f);
}
__device__ __forceinline__ float pl_nlos(float da, float db) {     // :109-111
  const float nj = fmaxf(2.8f - 0.0024f * db, 1.84f);
  return pl_los(da) + 20.f - 12.5f * nj + 10.f * nj * flog10(db) + 3.f * flog10(2.f / 5.f);
}
__device__ __forceinline__ float v2v_pathloss(float ax, float ay, float bx, float by) {   // :93-120
  const float d1 = fabsf(ax - bx), d2 = fabsf(ay - by);
  if (fminf(d1, d2) < 7.f) return pl_los(hypotf(d1, d2) + 0.001f);
  return fminf(pl_nlos(d1, d2), pl_nlos(d2, d1));
}
// 20 log10 |(re + j im) / sqrt(2)| = 10 log10((re^2 + im^2) / 2)      (:85-91)
__device__ __forceinline__ float fading_db(float2 h) { return 10.f * flog10(0.5f * (h.x * h.x + h.y * h.y)); }

// one thread per (e, i, j); the j == 0 thread of a row also advances vehicle i's V2I link
template <int RB>
__global__ void env_channels_kernel(const float* __restrict__ pos, const float* __restrict__ vel, float* __restrict__ v2v_shadow,
                                    float* __restrict__ v2i_shadow, const float* __restrict__ z_v2v, const float* __restrict__ z_v2i,
                                    const float2* __restrict__ ff_v2v, const float2* __restrict__ ff_v2i, float* __restrict__ v2v_ff,
                                    float* __restrict__ v2i_ff, float* __restrict__ v2i_abs, long total, int N, int RBrt) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;     // (e * N + i) * N + j
  if (idx >= total) return;
  const int rb_n = RB ? RB : RBrt;
  const int j = (int)(idx % N);
  const long ei = idx / N;
  const int i = (int)(ei % N);
  const long e = ei / N;
  const float2 pi = reinterpret_cast<const float2*>(pos)[ei], pj = reinterpret_cast<const float2*>(pos)[e * N + j];
  const float di = 0.002f * vel[ei], dj = 0.002f * vel[e * N + j];  // Environment.py:386
  // V2V: shadow AR(1) over the distance both ends moved (:70-83), path loss, +50 dB on the diagonal (:389-390)
  const float dd = di + dj;
  const float sh = __expf(-dd / 10.f) * v2v_shadow[idx] + sqrtf(-expm1f(-2.f * dd / 10.f)) * z_v2v[idx];   // 1 - e^-x without cancellation
  v2v_shadow[idx] = sh;
  const float a = v2v_pathloss(pi.x, pi.y, pj.x, pj.y) + sh + (i == j ? 50.f : 0.f);
  if (RB == 4) {
    const float4* f = reinterpret_cast<const float4*>(ff_v2v + idx * 4);
    const float4 f0 = f[0], f1 = f[1];
    float4 o;
    o.x = a - fading_db(make_float2(f0.x, f0.y)); o.y = a - fading_db(make_float2(f0.z, f0.w));
    o.z = a - fading_db(make_float2(f1.x, f1.y)); o.w = a - fading_db(make_float2(f1.z, f1.w));
    reinterpret_cast<float4*>(v2v_ff)[idx] = o;
  } else {
    for (int r = 0; r < rb_n; ++r) v2v_ff[idx * rb_n + r] = a - fading_db(ff_v2v[idx * rb_n + r]);
  }
  if (j == 0) {                                                     // V2I link of vehicle i (:140-165, :391, :402-404)
    const float d1 = fabsf(pi.x - 375.f), d2 = fabsf(pi.y - 649.5f);
    const float dist = hypotf(d1, d2);
    const float pl = 128.1f + 37.6f * flog10(sqrtf(dist * dist + 23.5f * 23.5f) / 1000.f);
    const float s = expf(-di / 50.f) * v2i_shadow[ei] + sqrtf(-expm1f(-2.f * di / 50.f)) * z_v2i[ei];
    v2i_shadow[ei] = s;
    const float ab = pl + s;
    v2i_abs[ei] = ab;
    for (int r = 0; r < rb_n; ++r) v2i_ff[ei * rb_n + r] = ab - fading_db(ff_v2i[ei * rb_n + r]);
  }
}

// one thread per (e, i): features of node i and its mask words
__global__ void env_pack_state_kernel(const int* __restrict__ dest, const float* __restrict__ v2v_ff, const float* __restrict__ v2i_ff,
                                      float* __restrict__ node, float* __restrict__ edge, uint32_t* __restrict__ in_mask,
                                      uint32_t* __restrict__ out_mask, float* __restrict__ adj, long total, int N, int RB) {
  const long ei = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ei >= total) return;
  const int i = (int)(ei % N);
  const long e = ei / N;
  const int d = dest[ei];
  const float* V = v2v_ff + e * (long)N * N * RB;
  const int Dn = 2 * RB + 1;
  for (int r = 0; r < RB; ++r) {
    const float own = (V[((long)i * N + d) * RB + r] - 80.f) / 60.f;            // BS_brain.py:396-397
    float col = 0.f;
    for (int k = 0; k < N; ++k) col += V[((long)k * N + d) * RB + r];           // :401-402
    const float ed = (((col - V[((long)d * N + d) * RB + r]) - (N - 1) * 80.f) / 60.f - own) / (float)(N - 2);   // :401-406
    node[ei * Dn + r] = own;
    node[ei * Dn + RB + r] = (v2i_ff[ei * RB + r] - 80.f) / 60.f;               // :399
    edge[ei * RB + r] = ed;
  }
  node[ei * Dn + 2 * RB] = kV2VPower;                                           // :437-438
  // Adj[n][m] = 1 - I, Adj[dest[m]][m] = 0 (:441-445): in_mask[m] bit n, out_mask[n] bit m
  if (in_mask && N <= 32) {
    const uint32_t all = N == 32 ? 0xffffffffu : ((1u << N) - 1u);
    in_mask[ei] = all & ~(1u << i) & ~(1u << d);
    uint32_t om = 0;
    for (int m = 0; m < N; ++m)
      if (m != i && dest[e * N + m] != i) om |= 1u << m;
    out_mask[ei] = om;
  }
  if (adj) {
    for (int m = 0; m < N; ++m) adj[(e * N + i) * N + m] = (m != i && dest[e * N + m] != i) ? 1.f : 0.f;   // row n = i
  }
}

// one warp per environment, lane = V2V link (N <= 32)
__global__ void env_reward_kernel(const int* __restrict__ actions, const int* __restrict__ dest, const float* __restrict__ v2v_ff,
                                  const float* __restrict__ v2i_ff, const float* __restrict__ v2i_abs, float* __restrict__ v2v_rate,
                                  float* __restrict__ v2i_rate, float* __restrict__ interference, float* __restrict__ reward,
                                  float v2v_w, float v2i_w, int E, int N, int RB) {
  const int warp = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= E) return;
  const long e = warp;
  const float* V = v2v_ff + e * (long)N * N * RB;
  const bool on = lane < N;
  const int a = on ? actions[e * N + lane] : -1;
  const int rx = on ? dest[e * N + lane] : 0;
  // V2I interference per resource block (Environment.py:413-420): lane = transmitter
  const float mine = on ? db2lin(kV2VPower - v2i_ff[(e * N + lane) * RB + a] + kVehAnt + kBsAnt - kBsNF) : 0.f;
  float v2i_sum = 0.f;
  const int m = min(RB, N);
  for (int r = 0; r < RB; ++r) {
    float s = (a == r) ? mine : 0.f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0 && interference) interference[e * RB + r] = s;
    if (r < m) {                                                               // :454-456
      const float sig = kV2IPower - v2i_abs[e * N + r] + kVehAnt + kBsAnt - kBsNF;
      const float rate = log2f(1.f + db2lin(sig) / (s + kSig2));
      if (lane == 0 && v2i_rate) v2i_rate[e * m + r] = rate;
      v2i_sum += rate;
    }
  }
  // V2V link lane -> rx on channel a (:425-453)
  float rate = 0.f;
  if (on) {
    const float signal = db2lin(kV2VPower - V[((long)lane * N + rx) * RB + a] + 2.f * kVehAnt - kVehNF);
    float interf = kSig2;
    if (a < N) interf += db2lin(kV2IPower - V[((long)a * N + rx) * RB + a] + 2.f * kVehAnt - kVehNF);   // V2I link of block a (:436-440)
    for (int k = 0; k < N; ++k) {
      const int ak = actions[e * N + k];
      if (k != lane && ak == a) interf += db2lin(kV2VPower - V[((long)k * N + rx) * RB + a] + 2.f * kVehAnt - kVehNF);
    }
    rate = log2f(1.f + signal / interf);
    if (v2v_rate) v2v_rate[e * N + lane] = rate;
  }
  float tot = rate;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
  if (lane == 0 && reward) reward[e] = v2v_w * tot + v2i_w * v2i_sum;         // BS_brain.py:515-519
}

__device__ __forceinline__ bool cross_up(float c, float step, float lane) { return c <= lane && c + step >= lane; }
__device__ __forceinline__ bool cross_dn(float c, float step, float lane) { return c >= lane && c - step <= lane; }

// one thread per vehicle (Environment.py:236-345); u: the uniform draw used if the vehicle reaches a crossing
__global__ void env_move_kernel(float* __restrict__ pos, int* __restrict__ dir, const float* __restrict__ vel,
                                const float* __restrict__ u, long total) {
  const long ei = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ei >= total) return;
  float2 p = reinterpret_cast<float2*>(pos)[ei];
  int d = dir[ei];
  const float step = vel[ei] * kTimestep;
  const bool turn_ok = u[ei] < 0.4f;
  bool turned = false;
  if (d == 0) {                                   // up: left lanes first, then right lanes
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_up(p.y, step, c_left[j])) { if (turn_ok) { p.x -= step - (c_left[j] - p.y); p.y = c_left[j]; d = 2; turned = true; } break; }
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_up(p.y, step, c_right[j])) { if (turn_ok) { p.x += step + (c_right[j] - p.y); p.y = c_right[j]; d = 3; turned = true; } break; }
    if (!turned) p.y += step;
  } else if (d == 1) {                            // down
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_dn(p.y, step, c_left[j])) { if (turn_ok) { p.x -= step - (p.y - c_left[j]); p.y = c_left[j]; d = 2; turned = true; } break; }
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_dn(p.y, step, c_right[j])) { if (turn_ok) { p.x += step + (p.y - c_right[j]); p.y = c_right[j]; d = 3; turned = true; } break; }
    if (!turned) p.y -= step;
  } else if (d == 3) {                            // right: up lanes first, then down lanes
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_up(p.x, step, c_up[j])) { if (turn_ok) { p.y += step - (c_up[j] - p.x); p.x = c_up[j]; d = 0; turned = true; } break; }
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_up(p.x, step, c_down[j])) { if (turn_ok) { p.y -= step - (c_down[j] - p.x); p.x = c_down[j]; d = 1; turned = true; } break; }
    if (!turned) p.x += step;
  } else {                                        // left
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_dn(p.x, step, c_up[j])) { if (turn_ok) { p.y += step - (p.x - c_up[j]); p.x = c_up[j]; d = 0; turned = true; } break; }
    for (int j = 0; j < 6 && !turned; ++j)
      if (cross_dn(p.x, step, c_down[j])) { if (turn_ok) { p.y -= step - (p.x - c_down[j]); p.x = c_down[j]; d = 1; turned = true; } break; }
    if (!turned) p.x -= step;
  }
  if (p.x < 0.f || p.y < 0.f || p.x > kWidth || p.y > kHeight) {   // leaves the map: re-enter on the border lane (:323-343)
    if (d == 0) { d = 3; p.y = c_right[5]; }
    else if (d == 1) { d = 2; p.y = c_left[0]; }
    else if (d == 2) { d = 0; p.x = c_up[0]; }
    else { d = 1; p.x = c_down[5]; }
  }
  reinterpret_cast<float2*>(pos)[ei] = p;
  dir[ei] = d;
}

// one thread per vehicle: receiver = candidate floor(u (N-3)) among the other vehicles sorted by distance, the two
// farthest excluded (Environment.py:360-376; ties broken by index like a stable argsort)
__global__ void env_dest_kernel(const float* __restrict__ pos, const float* __restrict__ u, int* __restrict__ dest, long total, int N) {
  const long ei = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ei >= total) return;
  const int i = (int)(ei % N);
  const long e = ei / N;
  const float2* P = reinterpret_cast<const float2*>(pos) + e * N;
  const float2 me = P[i];
  const int ncand = N - 3;
  int want = 1 + min((int)(u[ei] * ncand), ncand - 1);            // rank in the sorted order, rank 0 = the vehicle itself
  int pick = i;
  for (int k = 0; k < N; ++k) {                                   // the vehicle itself always ranks first, also when another
    const float dk = k == i ? -1.f : hypotf(P[k].x - me.x, P[k].y - me.y);   // vehicle sits on the same spot
    int rank = 0;
    for (int m = 0; m < N; ++m) {
      const float dm = m == i ? -1.f : hypotf(P[m].x - me.x, P[m].y - me.y);
      rank += (dm < dk) || (dm == dk && m < k);
    }
    if (rank == want) pick = k;
  }
  dest[ei] = pick;
}

}  // namespace
}  // namespace v2v

using namespace v2v;

static inline unsigned blocks_for(long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

extern "C" int v2v_env_renew_channels(const float* pos, const float* vel, float* v2v_shadow, float* v2i_shadow, const float* z_v2v,
                                      const float* z_v2i, const float* ff_v2v, const float* ff_v2i, float* v2v_ff, float* v2i_ff,
                                      float* v2i_abs, int E, int N, int RB, void* stream) {
  V2V_REQUIRE(E >= 0 && N >= 1 && RB >= 1, "v2v_env_renew_channels: bad shape E=%d N=%d RB=%d", E, N, RB);
  if (E == 0) return 0;
  V2V_REQUIRE(pos && vel && v2v_shadow && v2i_shadow && z_v2v && z_v2i && ff_v2v && ff_v2i && v2v_ff && v2i_ff && v2i_abs,
              "v2v_env_renew_channels: null pointer");
  const long total = (long)E * N * N;
  cudaStream_t st = (cudaStream_t)stream;
  if (RB == 4)
    env_channels_kernel<4><<<blocks_for(total, 256), 256, 0, st>>>(pos, vel, v2v_shadow, v2i_shadow, z_v2v, z_v2i, (const float2*)ff_v2v,
                                                                    (const float2*)ff_v2i, v2v_ff, v2i_ff, v2i_abs, total, N, RB);
  else
    env_channels_kernel<0><<<blocks_for(total, 256), 256, 0, st>>>(pos, vel, v2v_shadow, v2i_shadow, z_v2v, z_v2i, (const float2*)ff_v2v,
                                                                    (const float2*)ff_v2i, v2v_ff, v2i_ff, v2i_abs, total, N, RB);
  return launch_status("env_channels_kernel");
}

extern "C" int v2v_env_pack_state(const int* dest, const float* v2v_ff, const float* v2i_ff, float* node, float* edge,
                                  uint32_t* in_mask, uint32_t* out_mask, float* adj, int E, int N, int RB, void* stream) {
  V2V_REQUIRE(E >= 0 && N >= 3 && RB >= 1, "v2v_env_pack_state: bad shape E=%d N=%d RB=%d (N >= 3)", E, N, RB);
  if (E == 0) return 0;
  V2V_REQUIRE(dest && v2v_ff && v2i_ff && node && edge, "v2v_env_pack_state: null pointer");
  V2V_REQUIRE((in_mask == nullptr) == (out_mask == nullptr), "v2v_env_pack_state: pass both masks or neither");
  V2V_REQUIRE(!in_mask || N <= 32, "v2v_env_pack_state: bit masks need N <= 32 (use adj)");
  const long total = (long)E * N;
  env_pack_state_kernel<<<blocks_for(total, 128), 128, 0, (cudaStream_t)stream>>>(dest, v2v_ff, v2i_ff, node, edge, in_mask, out_mask, adj,
                                                                                    total, N, RB);
  return launch_status("env_pack_state_kernel");
}

extern "C" int v2v_env_reward(const int* actions, const int* dest, const float* v2v_ff, const float* v2i_ff, const float* v2i_abs,
                              float* v2v_rate, float* v2i_rate, float* interference, float* reward, float v2v_weight, float v2i_weight,
                              int E, int N, int RB, void* stream) {
  V2V_REQUIRE(E >= 0 && N >= 1 && N <= 32 && RB >= 1, "v2v_env_reward: bad shape E=%d N=%d RB=%d (N <= 32)", E, N, RB);
  if (E == 0) return 0;
  V2V_REQUIRE(actions && dest && v2v_ff && v2i_ff && v2i_abs, "v2v_env_reward: null pointer");
  env_reward_kernel<<<blocks_for((long)E * 32, 128), 128, 0, (cudaStream_t)stream>>>(actions, dest, v2v_ff, v2i_ff, v2i_abs, v2v_rate, v2i_rate,
                                                                                       interference, reward, v2v_weight, v2i_weight, E, N, RB);
  return launch_status("env_reward_kernel");
}

extern "C" int v2v_env_renew_positions(float* pos, int* dir, const float* vel, const float* u, int E, int N, void* stream) {
  V2V_REQUIRE(E >= 0 && N >= 1, "v2v_env_renew_positions: bad shape");
  if (E == 0) return 0;
  V2V_REQUIRE(pos && dir && vel && u, "v2v_env_renew_positions: null pointer");
  const long total = (long)E * N;
  env_move_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(pos, dir, vel, u, total);
  return launch_status("env_move_kernel");
}

extern "C" int v2v_env_choose_destinations(const float* pos, const float* u, int* dest, int E, int N, void* stream) {
  V2V_REQUIRE(E >= 0 && N >= 4, "v2v_env_choose_destinations: N >= 4 (the reference excludes the vehicle and the two farthest)");
  if (E == 0) return 0;
  V2V_REQUIRE(pos && u && dest, "v2v_env_choose_destinations: null pointer");
  const long total = (long)E * N;
  env_dest_kernel<<<blocks_for(total, 128), 128, 0, (cudaStream_t)stream>>>(pos, u, dest, total, N);
  return launch_status("env_dest_kernel");
}
