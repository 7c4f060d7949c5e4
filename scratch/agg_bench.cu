// Exploration harness for the aggregation fast path (not product code).
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../globecom2020-resourceallocationgnn_b200/csrc/agg_kernels.cuh"
using namespace v2v;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}

struct Sets { std::vector<float*> H, O; std::vector<uint32_t*> M; int P; };

template <typename F>
float time_graph(F launch, int P, int reps, cudaStream_t st) {
  for (int i = 0; i < P; ++i) launch(i);
  CK(cudaStreamSynchronize(st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < P; ++i) launch(i);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return 1e3f * ms / (reps * P);
}

int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 8192, N = 20;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  size_t hbytes = (size_t)B * N * 16 * 4, mbytes = (size_t)B * N * 4;
  double alg = 2.0 * hbytes + mbytes;
  int P = std::max(8, (int)((2.0 * 126 * 1024 * 1024) / alg) + 1);
  Sets s; s.P = P;
  std::vector<uint32_t> hm((size_t)B * N);
  for (int b = 0; b < B; ++b) for (int m = 0; m < N; ++m) {
    uint32_t full = (1u << N) - 1; int dest = (m + 1 + (b * 7 + m * 3) % (N - 1)) % N;
    hm[(size_t)b * N + m] = full & ~(1u << m) & ~(1u << dest);
  }
  for (int i = 0; i < P; ++i) {
    float* h; float* o; uint32_t* m;
    CK(cudaMalloc(&h, hbytes)); CK(cudaMalloc(&o, hbytes)); CK(cudaMalloc(&m, mbytes));
    CK(cudaMemset(h, 0x3c, hbytes)); CK(cudaMemcpy(m, hm.data(), mbytes, cudaMemcpyHostToDevice));
    s.H.push_back(h); s.O.push_back(o); s.M.push_back(m);
  }
  printf("B=%d N=%d alg bytes/launch=%.0f P=%d  peak-time(6451.8GB/s)=%.2f us\n", B, N, alg, P, alg / 6451.8e3);
  int reps = 20;
  auto report = [&](const char* name, float us) { printf("%-58s %7.2f us  %7.1f GB/s  frac %.3f\n", name, us, alg / us / 1e3, alg / us / 1e3 / 6451.8); };

  report("cudaMemcpyAsync D2D (same H bytes)", time_graph([&](int i) { CK(cudaMemcpyAsync(s.O[i], s.H[i], hbytes, cudaMemcpyDeviceToDevice, st)); }, P, reps, st));
  report("float4 copy kernel 148x8 x 256thr", time_graph([&](int i) { copy_kernel<<<148 * 8, 256, 0, st>>>((const float4*)s.H[i], (float4*)s.O[i], hbytes / 16); }, P, reps, st));

  for (int pdl = 0; pdl <= 1; ++pdl) for (int dep = 1; dep >= 0; --dep) for (int cps = 1; cps <= 2; ++cps) {
    if (!pdl && !dep) continue;
    AggLaunchCfg cfg; cfg.pdl = pdl; cfg.dep_wait = dep; cfg.ctas_per_sm = cps;
    char nm[128];
    snprintf(nm, 128, "agg<10,2> W8  ctas/SM=%d pdl=%d dep_wait=%d", cps, pdl, dep);
    report(nm, time_graph([&](int i) { launch_agg_fast<float, 10, 2, false, 8>(s.H[i], s.M[i], nullptr, s.O[i], B, N, cfg, st); }, P, reps, st));
    snprintf(nm, 128, "agg<10,2> W8  ctas/SM=%d pdl=%d dep_wait=%d NOCOMPUTE", cps, pdl, dep);
    report(nm, time_graph([&](int i) { launch_agg_fast<float, 10, 2, false, 8, false>(s.H[i], s.M[i], nullptr, s.O[i], B, N, cfg, st); }, P, reps, st));
    snprintf(nm, 128, "agg<5,4>  W8  ctas/SM=%d pdl=%d dep_wait=%d", cps, pdl, dep);
    report(nm, time_graph([&](int i) { launch_agg_fast<float, 5, 4, false, 8>(s.H[i], s.M[i], nullptr, s.O[i], B, N, cfg, st); }, P, reps, st));
    if (cps == 1) {
      snprintf(nm, 128, "agg<10,2> W16 ctas/SM=%d pdl=%d dep_wait=%d", cps, pdl, dep);
      report(nm, time_graph([&](int i) { launch_agg_fast<float, 10, 2, false, 16>(s.H[i], s.M[i], nullptr, s.O[i], B, N, cfg, st); }, P, reps, st));
      snprintf(nm, 128, "agg<5,4>  W16 ctas/SM=%d pdl=%d dep_wait=%d", cps, pdl, dep);
      report(nm, time_graph([&](int i) { launch_agg_fast<float, 5, 4, false, 16>(s.H[i], s.M[i], nullptr, s.O[i], B, N, cfg, st); }, P, reps, st));
    }
  }
  return 0;
}
