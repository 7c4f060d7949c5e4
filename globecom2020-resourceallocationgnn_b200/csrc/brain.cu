// The brain: BS._create_model / train_dnn / predict / update_target_model
// (BS_brain.py:90-239) as a stream-ordered sequence of the engine's kernels.
//
// Wiring (BS_brain.py:147-200), S = number of GNN stages, last stage linear:
//   h0 = act([node|edge] . W0[:Dn+De])              (third input is zeros, :478/:589)
//   a0 = Agg(h0)
//   hs = act([h(s-1)|node|edge|a(s-1)] . Ws), as = Agg(hs)          s = 1..S-1
//   z  = [node | h(S-1) | a(S-1)] -> Dense 80 relu, 40 relu, 20 relu, CH linear
// Loss: per-head mean Huber, heads summed (:86-87, :214).  Optimiser: Keras Adam (:212).
#include <math.h>
#include <time.h>
#include <stdarg.h>
#include <vector>

#include <map>

#include <atomic>
#include "fused.cuh"
#include "host_stage.cuh"
#include "tc_forward.cuh"
#include "tc_train.cuh"

namespace v2v {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return 1;
}

long& launch_counter() {
  static long n = 0;
  return n;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    n = p.multiProcessorCount;
  }
  return n;
}

struct LayerDesc {
  int K, n_out;       // stacked weight rows / columns
  size_t w_off, b_off;  // offsets (floats) into a parameter set
};

}  // namespace v2v

using namespace v2v;

struct v2v_brain {
  v2v_brain_config cfg;
  int N, Dn, De, F, CH, S, G, n_layers;
  std::vector<LayerDesc> layers;
  size_t n_params = 0;
  int iterations = 0;
  // device memory
  float* params[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // online, target, grad, m, v
  std::vector<float*> h, agg;     // per stage [maxB][N][F]
  std::vector<float*> mlp;        // per MLP layer output [maxB][N][n_out]
  std::vector<float*> dmlp;       // gradient w.r.t. MLP layer *inputs* (gated), index j: input of layer j (j>=1)
  float* dq = nullptr;
  float* dh_a = nullptr;
  float* dh_b = nullptr;
  float* dagg = nullptr;
  float* head_loss = nullptr;     // [N]
  // staging for the host-buffer entry points
  float* st_node = nullptr; float* st_edge = nullptr; float* st_neigh = nullptr; float* st_adj = nullptr; float* st_y = nullptr;
  float* st_q = nullptr;
  uint32_t* st_in_mask = nullptr; uint32_t* st_out_mask = nullptr;
  int* st_flag = nullptr;
  int* st_flag_host = nullptr;    // pinned
  float* st_block = nullptr;      // ONE device block behind st_node .. st_out_mask, laid out like the pinned block below, so that
                                  // small batches travel as a single H2D copy (node | edge | y | in_mask | out_mask are adjacent)
  float* pin = nullptr;           // pinned staging of the *_views entry points, same offsets as st_block
  size_t pin_node = 0, pin_edge = 0, pin_neigh = 0, pin_adj = 0, pin_y = 0, pin_q = 0, pin_hl = 0, pin_im = 0, pin_om = 0;   // offsets (floats)
  // fused whole-network path (shared weights, N <= 32, binary adjacency)
  bool fused_capable = false;
  bool fused_enabled = true;
  std::vector<int> lK, lO;
  std::vector<size_t> lw, lb;
  struct FusedEntry { FusedProgram host; FusedProgram* dev = nullptr; };
  std::map<long, FusedEntry*> fused_cache;      // key: (TG << 1) | train
  std::map<long, int> tg_cache;                 // key: (B << 1) | train
  float* partial = nullptr;                     // [sm_count][n_params] per-CTA gradient partials
  int partial_ctas = 0;
  int last_grid = 0;                            // grid of the last fused train launch (0: layered path ran)
  bool defer_reduce = false;
  // tensor-core (tcgen05) forward of the shared-weight brain
  bool tc_capable = false;
  int tc_mode = 1;                              // 0: off, 1: auto (batches with >= 2 tiles per SM), 2: always
  TcPlan tc_plan;
  TcPlan* tc_plan_dev = nullptr;
  float* tc_wimg = nullptr;                     // staged weight image (hi/lo planes + biases) of the last call
  // bf16 configuration (cfg.dtype == V2V_BF16): every contraction of forward AND backward on tcgen05 (tc_train.cu)
  bool bf16 = false;
  TtPlan* tt_plan = nullptr;                    // host copy
  TtPlan* tt_plan_dev = nullptr;
};

static FusedShape fused_shape(const v2v_brain* b) {
  FusedShape s;
  s.N = b->N; s.Dn = b->Dn; s.De = b->De; s.F = b->F; s.CH = b->CH; s.S = b->S; s.G = b->G;
  s.H1 = b->cfg.hidden[0]; s.H2 = b->cfg.hidden[1]; s.H3 = b->cfg.hidden[2];
  s.layer_K = b->lK.data(); s.layer_O = b->lO.data(); s.w_off = b->lw.data(); s.b_off = b->lb.data();
  s.n_layers = b->n_layers; s.n_params = b->n_params;
  return s;
}

// program for (B, train), built and uploaded on first use (not during stream capture)
static int fused_get(v2v_brain* b, int B, int train, v2v_brain::FusedEntry** out) {
  const long bkey = ((long)B << 1) | train;
  auto it = b->tg_cache.find(bkey);
  int tg;
  if (it == b->tg_cache.end()) {
    tg = fused_pick_tg(fused_shape(b), B, train);
    b->tg_cache[bkey] = tg;
  } else {
    tg = it->second;
  }
  V2V_REQUIRE(tg > 0, "fused path: no tile size fits shared memory");
  const long key = ((long)tg << 1) | train;
  auto ie = b->fused_cache.find(key);
  if (ie == b->fused_cache.end()) {
    auto* e = new v2v_brain::FusedEntry();
    if (int rc = fused_build_program(fused_shape(b), tg, train, &e->host)) { delete e; return rc; }
    if (cudaMalloc((void**)&e->dev, sizeof(FusedProgram)) != cudaSuccess) { delete e; return fail("fused path: cudaMalloc failed"); }
    if (cudaMemcpy(e->dev, &e->host, sizeof(FusedProgram), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {      // the kernel reads the program before its dependency wait
      cudaFree(e->dev); delete e; return fail("fused path: program upload failed");
    }
    b->fused_cache[key] = e;
    *out = e;
  } else {
    *out = ie->second;
  }
  return 0;
}

static bool use_fused(const v2v_brain* b, const uint32_t* in_mask, const float* neigh) {
  return b->fused_capable && b->fused_enabled && in_mask != nullptr && neigh == nullptr;
}

static int dmalloc(float** p, size_t n_floats) {
  V2V_CHECK_CUDA(cudaMalloc((void**)p, std::max<size_t>(n_floats, 1) * sizeof(float)));
  return 0;
}

extern "C" const char* v2v_last_error(void) { return last_error().c_str(); }
extern "C" int v2v_version(void) { return 100; }
extern "C" int v2v_device_sm_count(void) { return sm_count(); }
extern "C" long v2v_launch_count(void) { return launch_counter(); }

extern "C" void v2v_brain_destroy(v2v_brain* b) {
  if (!b) return;
  for (auto& p : b->params) cudaFree(p);
  for (auto p : b->h) cudaFree(p);
  for (auto p : b->agg) cudaFree(p);
  for (auto p : b->mlp) cudaFree(p);
  for (auto p : b->dmlp) cudaFree(p);
  cudaFree(b->dq); cudaFree(b->dh_a); cudaFree(b->dh_b); cudaFree(b->dagg); cudaFree(b->head_loss);
  cudaFree(b->st_block); cudaFree(b->st_flag);
  if (b->st_flag_host) cudaFreeHost(b->st_flag_host);
  if (b->pin) cudaFreeHost(b->pin);
  cudaFree(b->partial);
  cudaFree(b->tc_plan_dev);
  cudaFree(b->tc_wimg);
  cudaFree(b->tt_plan_dev);
  delete b->tt_plan;
  for (auto& kv : b->fused_cache) { cudaFree(kv.second->dev); delete kv.second; }
  delete b;
}

extern "C" int v2v_brain_create(const v2v_brain_config* cfg, v2v_brain** out) {
  V2V_REQUIRE(cfg && out, "v2v_brain_create: null argument");
  V2V_REQUIRE(cfg->num_d2d >= 1 && cfg->num_d2d <= 256, "v2v_brain_create: num_d2d=%d out of range [1,256]", cfg->num_d2d);
  V2V_REQUIRE(cfg->node_dim > 0 && cfg->edge_dim > 0 && cfg->feedback > 0 && cfg->num_ch > 0, "v2v_brain_create: non-positive dimension");
  V2V_REQUIRE(cfg->stages >= 1 && cfg->stages <= 8, "v2v_brain_create: stages=%d out of range [1,8]", cfg->stages);
  V2V_REQUIRE(cfg->max_batch >= 1, "v2v_brain_create: max_batch must be >= 1");
  V2V_REQUIRE(cfg->dtype == V2V_F32 || cfg->dtype == V2V_BF16, "v2v_brain_create: unknown dtype %d", cfg->dtype);
  V2V_REQUIRE(cfg->dtype == V2V_F32 || !cfg->per_slot, "v2v_brain_create: the bf16 brain is the shared-weight one (per_slot = 0)");
  for (int i = 0; i < 3; ++i) V2V_REQUIRE(cfg->hidden[i] > 0, "v2v_brain_create: hidden[%d] must be > 0", i);
  v2v_brain* b = new v2v_brain();
  b->cfg = *cfg;
  b->N = cfg->num_d2d; b->Dn = cfg->node_dim; b->De = cfg->edge_dim; b->F = cfg->feedback;
  b->CH = cfg->num_ch; b->S = cfg->stages; b->G = cfg->per_slot ? cfg->num_d2d : 1;
  size_t off = 0;
  auto add_layer = [&](int K, int n_out) {
    LayerDesc L;
    L.K = K; L.n_out = n_out;
    L.w_off = off; off += (size_t)b->G * K * n_out;
    L.b_off = off; off += (size_t)b->G * n_out;
    b->layers.push_back(L);
  };
  for (int s = 0; s < b->S; ++s) add_layer((s == 0 ? b->Dn : b->F + b->Dn) + b->De + b->F, b->F);
  int k = b->Dn + 2 * b->F;
  for (int i = 0; i < 3; ++i) { add_layer(k, cfg->hidden[i]); k = cfg->hidden[i]; }
  add_layer(k, b->CH);
  b->n_layers = (int)b->layers.size();
  b->n_params = off;

  const size_t mb = (size_t)cfg->max_batch, rows = mb * b->N;
  int rc = 0;
  for (auto& p : b->params) rc |= dmalloc(&p, b->n_params);
  b->h.resize(b->S); b->agg.resize(b->S);
  for (int s = 0; s < b->S; ++s) { rc |= dmalloc(&b->h[s], rows * b->F); rc |= dmalloc(&b->agg[s], rows * b->F); }
  b->mlp.resize(4); b->dmlp.resize(4, nullptr);
  for (int j = 0; j < 4; ++j) {
    rc |= dmalloc(&b->mlp[j], rows * b->layers[b->S + j].n_out);
    if (j >= 1) rc |= dmalloc(&b->dmlp[j], rows * b->layers[b->S + j].K);
  }
  rc |= dmalloc(&b->dq, rows * b->CH);
  rc |= dmalloc(&b->dh_a, rows * b->F); rc |= dmalloc(&b->dh_b, rows * b->F); rc |= dmalloc(&b->dagg, rows * b->F);
  rc |= dmalloc(&b->head_loss, b->N);
  const int W = ceil_div(b->N, 32);
  if (cudaMalloc((void**)&b->st_flag, sizeof(int)) != cudaSuccess) rc = 1;
  if (cudaMallocHost((void**)&b->st_flag_host, sizeof(int)) != cudaSuccess) rc = 1;
  {
    // staging layout (floats), identical in the pinned block and in the device block: the tensors every call ships are
    // adjacent (node | edge | y | in_mask | out_mask), the rarely shipped ones (neighbour input, dense adjacency) follow
    size_t o = 0;
    auto take = [&](size_t n) { const size_t at = o; o += (n + 63) & ~(size_t)63; return at; };
    b->pin_node = take(rows * b->Dn); b->pin_edge = take(rows * b->De); b->pin_y = take(rows * b->CH);
    b->pin_im = take(rows * W); b->pin_om = take(rows * W);
    b->pin_neigh = take(rows * b->F); b->pin_adj = take(rows * b->N); b->pin_q = take(rows * b->CH); b->pin_hl = take(b->N);
    if (cudaMallocHost((void**)&b->pin, o * sizeof(float)) != cudaSuccess) rc = 1;
    if (cudaMalloc((void**)&b->st_block, o * sizeof(float)) != cudaSuccess) rc = 1;
    if (!rc) {
      b->st_node = b->st_block + b->pin_node; b->st_edge = b->st_block + b->pin_edge; b->st_y = b->st_block + b->pin_y;
      b->st_in_mask = reinterpret_cast<uint32_t*>(b->st_block + b->pin_im);
      b->st_out_mask = reinterpret_cast<uint32_t*>(b->st_block + b->pin_om);
      b->st_neigh = b->st_block + b->pin_neigh; b->st_adj = b->st_block + b->pin_adj; b->st_q = b->st_block + b->pin_q;
    }
  }
  if (rc) {
    std::string e = last_error();
    v2v_brain_destroy(b);
    return fail("v2v_brain_create: device allocation failed (%s)", e.c_str());
  }
  for (auto& p : b->params) cudaMemset(p, 0, b->n_params * sizeof(float));
  for (auto& L : b->layers) { b->lK.push_back(L.K); b->lO.push_back(L.n_out); b->lw.push_back(L.w_off); b->lb.push_back(L.b_off); }
  // fused whole-network kernel: shared weights up to N = 32, per-slot weights (the reference's model) up to N = 8
  if ((b->G == 1 && b->N <= 32) || (b->G == b->N && b->N <= 8)) {
    FusedProgram* probe = new FusedProgram();
    if (fused_build_program(fused_shape(b), 1, 1, probe) == 0 && fused_smem_bytes(*probe) <= 226 * 1024) {
      b->partial_ctas = sm_count();
      const size_t pbytes = (size_t)b->partial_ctas * fused_partial_stride((long)b->n_params) * sizeof(float);
      if (cudaMalloc((void**)&b->partial, pbytes) == cudaSuccess) {
        cudaMemset(b->partial, 0, pbytes);
        b->fused_capable = true;
      }
    }
    delete probe;
    last_error().clear();
  }
  if (b->G == 1 && b->N <= 32) {
    // tensor-core forward plan (predict at large batch)
    TcShape ts;
    ts.N = b->N; ts.Dn = b->Dn; ts.De = b->De; ts.F = b->F; ts.CH = b->CH; ts.S = b->S;
    ts.H1 = cfg->hidden[0]; ts.H2 = cfg->hidden[1]; ts.H3 = cfg->hidden[2];
    ts.w_off = b->lw.data(); ts.b_off = b->lb.data();
    if (tc_build_plan(ts, &b->tc_plan) == 0 && cudaMalloc((void**)&b->tc_plan_dev, sizeof(TcPlan)) == cudaSuccess &&
        cudaMemcpy(b->tc_plan_dev, &b->tc_plan, sizeof(TcPlan), cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMalloc((void**)&b->tc_wimg, (size_t)(b->tc_plan.w_floats + b->tc_plan.bias_floats) * sizeof(float)) == cudaSuccess) {
      b->tc_capable = true;
    }
    last_error().clear();
    if (const char* e = getenv("V2V_TENSOR_CORE")) b->tc_mode = atoi(e);
  }
  if (cfg->dtype == V2V_BF16) {
    TtShape ts;
    ts.N = b->N; ts.Dn = b->Dn; ts.De = b->De; ts.F = b->F; ts.CH = b->CH; ts.S = b->S;
    ts.H1 = cfg->hidden[0]; ts.H2 = cfg->hidden[1]; ts.H3 = cfg->hidden[2];
    ts.w_off = b->lw.data(); ts.b_off = b->lb.data(); ts.n_params = b->n_params;
    b->tt_plan = new TtPlan();
    int rc2 = tt_build_plan(ts, b->tt_plan);
    if (rc2 == 0) {
      if (cudaMalloc((void**)&b->tt_plan_dev, sizeof(TtPlan)) != cudaSuccess ||
          cudaMemcpy(b->tt_plan_dev, b->tt_plan, sizeof(TtPlan), cudaMemcpyHostToDevice) != cudaSuccess)
        rc2 = fail("v2v_brain_create: device allocation failed (bf16 plan)");
    }
    if (rc2 == 0 && !b->partial) {
      b->partial_ctas = sm_count();
      const size_t pbytes = (size_t)b->partial_ctas * tt_partial_stride((long)b->n_params) * sizeof(float);
      if (cudaMalloc((void**)&b->partial, pbytes) != cudaSuccess) rc2 = fail("v2v_brain_create: device allocation failed (partials)");
      else cudaMemset(b->partial, 0, pbytes);
    }
    if (rc2) {
      std::string e = last_error();
      v2v_brain_destroy(b);
      return fail("%s", e.c_str());
    }
    b->bf16 = true;
  }
  cudaDeviceSynchronize();
  *out = b;
  return 0;
}

extern "C" long v2v_brain_param_count(const v2v_brain* b) { return b ? (long)b->n_params : -1; }
extern "C" float* v2v_brain_param_ptr(v2v_brain* b, int which) {
  return (b && which >= 0 && which < 5) ? b->params[which] : nullptr;
}
extern "C" int v2v_brain_get_iterations(const v2v_brain* b) { return b ? b->iterations : -1; }
extern "C" int v2v_brain_set_iterations(v2v_brain* b, int t) {
  V2V_REQUIRE(b && t >= 0, "v2v_brain_set_iterations: bad argument");
  b->iterations = t;
  return 0;
}

extern "C" int v2v_brain_get_params(v2v_brain* b, int which, float* host_out, void* stream) {
  V2V_REQUIRE(b && host_out && which >= 0 && which < 5, "v2v_brain_get_params: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  V2V_CHECK_CUDA(cudaMemcpyAsync(host_out, b->params[which], b->n_params * sizeof(float), cudaMemcpyDeviceToHost, st));
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int v2v_brain_set_params(v2v_brain* b, int which, const float* host_in, void* stream) {
  V2V_REQUIRE(b && host_in && which >= 0 && which < 5, "v2v_brain_set_params: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->params[which], host_in, b->n_params * sizeof(float), cudaMemcpyHostToDevice, st));
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int v2v_fused_set_trace(long long* dev_buf) { return fused_set_trace(dev_buf); }
extern "C" int v2v_fused_set_mma(int mode) { return fused_set_mma(mode); }
extern "C" int v2v_fused_get_mma(void) { return fused_get_mma(); }
extern "C" int v2v_tt_set_trace(long long* dev_buf) { return tt_set_trace(dev_buf); }

extern "C" int v2v_brain_set_fused(v2v_brain* b, int enable) {
  V2V_REQUIRE(b, "v2v_brain_set_fused: null brain");
  b->fused_enabled = enable != 0;
  return 0;
}

extern "C" int v2v_brain_fused_info(v2v_brain* b, int B, int train, int* info8) {
  V2V_REQUIRE(b && info8 && B > 0, "v2v_brain_fused_info: bad argument");
  for (int i = 0; i < 8; ++i) info8[i] = 0;
  if (!b->fused_capable) return 0;
  FusedProgram* p = new FusedProgram();
  const int tg = fused_pick_tg(fused_shape(b), B, train);
  int rc = tg > 0 ? fused_build_program(fused_shape(b), tg, train, p) : 1;
  if (rc == 0) {
    info8[0] = 1; info8[1] = tg; info8[2] = p->n_rows; info8[3] = (int)fused_smem_bytes(*p); info8[4] = p->n_ops;
    info8[5] = p->n_blocks; info8[6] = p->n_bias; info8[7] = p->n_tab;
  }
  delete p;
  return rc;
}

// Host-only planning query of the tensor-core forward (no device needed).  info8 = {capable, graphs per tile, layers,
// shared-memory bytes, operand planes per tile slot, floats of the staged weight image, tensor-memory columns used per
// slot, tcgen05.mma instructions per tile}.
extern "C" int v2v_tc_plan(const v2v_brain_config* cfg, int* info8) {
  V2V_REQUIRE(cfg && info8, "v2v_tc_plan: bad argument");
  for (int i = 0; i < 8; ++i) info8[i] = 0;
  if (cfg->per_slot || cfg->num_d2d > 32 || cfg->num_d2d < 1 || cfg->stages < 1 || cfg->stages > 8) return 0;
  std::vector<size_t> lw, lb;
  size_t off = 0;
  auto add_layer = [&](int K, int O) { lw.push_back(off); off += (size_t)K * O; lb.push_back(off); off += O; };
  for (int s = 0; s < cfg->stages; ++s) add_layer((s == 0 ? cfg->node_dim : cfg->feedback + cfg->node_dim) + cfg->edge_dim + cfg->feedback, cfg->feedback);
  int k = cfg->node_dim + 2 * cfg->feedback;
  for (int i = 0; i < 3; ++i) { add_layer(k, cfg->hidden[i]); k = cfg->hidden[i]; }
  add_layer(k, cfg->num_ch);
  TcShape ts;
  ts.N = cfg->num_d2d; ts.Dn = cfg->node_dim; ts.De = cfg->edge_dim; ts.F = cfg->feedback; ts.CH = cfg->num_ch; ts.S = cfg->stages;
  ts.H1 = cfg->hidden[0]; ts.H2 = cfg->hidden[1]; ts.H3 = cfg->hidden[2];
  ts.w_off = lw.data(); ts.b_off = lb.data();
  TcPlan* p = new TcPlan();
  if (tc_build_plan(ts, p) == 0) {
    int cols = 0, mmas = 0;
    for (int l = 0; l < p->n_layers; ++l) {
      cols = std::max(cols, p->layers[l].dcol + 2 * p->layers[l].Npad);
      mmas += 2 * (p->layers[l].Kpad / 8);
    }
    info8[0] = 1; info8[1] = p->TG; info8[2] = p->n_layers; info8[3] = p->smem_bytes; info8[4] = p->stage_planes;
    info8[5] = p->w_floats + p->bias_floats; info8[6] = cols; info8[7] = mmas;
  }
  last_error().clear();
  delete p;
  return 0;
}

// Host-only planning query of the bf16 tensor-core training kernel (csrc/tc_train.cu).  info8 = {capable, graphs per
// tile, steps per training tile, tcgen05.mma instructions per training tile, shared-memory bytes, operand planes
// (activations + gradients), bf16 weight-image elements, weight-gradient column blocks}.
extern "C" int v2v_tt_plan(const v2v_brain_config* cfg, int* info8) {
  V2V_REQUIRE(cfg && info8, "v2v_tt_plan: bad argument");
  for (int i = 0; i < 8; ++i) info8[i] = 0;
  if (cfg->per_slot || cfg->num_d2d < 1) return 0;
  std::vector<size_t> lw, lb;
  size_t off = 0;
  auto add_layer = [&](int K, int O) { lw.push_back(off); off += (size_t)K * O; lb.push_back(off); off += O; };
  for (int s = 0; s < cfg->stages; ++s) add_layer((s == 0 ? cfg->node_dim : cfg->feedback + cfg->node_dim) + cfg->edge_dim + cfg->feedback, cfg->feedback);
  int k = cfg->node_dim + 2 * cfg->feedback;
  for (int i = 0; i < 3; ++i) { add_layer(k, cfg->hidden[i]); k = cfg->hidden[i]; }
  add_layer(k, cfg->num_ch);
  TtShape ts;
  ts.N = cfg->num_d2d; ts.Dn = cfg->node_dim; ts.De = cfg->edge_dim; ts.F = cfg->feedback; ts.CH = cfg->num_ch; ts.S = cfg->stages;
  ts.H1 = cfg->hidden[0]; ts.H2 = cfg->hidden[1]; ts.H3 = cfg->hidden[2];
  ts.w_off = lw.data(); ts.b_off = lb.data(); ts.n_params = off;
  TtPlan* p = new TtPlan();
  if (tt_build_plan(ts, p) == 0) {
    info8[0] = 1; info8[1] = p->TG; info8[2] = p->n_steps_train; info8[3] = p->n_mma; info8[4] = p->smem_bytes;
    info8[5] = p->x_planes + p->dz_planes; info8[6] = p->w_elems; info8[7] = p->n_blocks;
  }
  last_error().clear();
  delete p;
  return 0;
}

// Host-only planning query (no device needed): the fused program a brain of this configuration would
// run for a batch of B graphs.  info8 as in v2v_brain_fused_info.
extern "C" int v2v_fused_plan(const v2v_brain_config* cfg, int B, int train, int* info8) {
  V2V_REQUIRE(cfg && info8 && B > 0, "v2v_fused_plan: bad argument");
  for (int i = 0; i < 8; ++i) info8[i] = 0;
  if (cfg->num_d2d > 32 || (cfg->per_slot && cfg->num_d2d > 8)) return 0;
  const int G = cfg->per_slot ? cfg->num_d2d : 1;
  std::vector<int> lK, lO;
  std::vector<size_t> lw, lb;
  size_t off = 0;
  auto add_layer = [&](int K, int O) { lK.push_back(K); lO.push_back(O); lw.push_back(off); off += (size_t)G * K * O; lb.push_back(off); off += (size_t)G * O; };
  for (int s = 0; s < cfg->stages; ++s) add_layer((s == 0 ? cfg->node_dim : cfg->feedback + cfg->node_dim) + cfg->edge_dim + cfg->feedback, cfg->feedback);
  int k = cfg->node_dim + 2 * cfg->feedback;
  for (int i = 0; i < 3; ++i) { add_layer(k, cfg->hidden[i]); k = cfg->hidden[i]; }
  add_layer(k, cfg->num_ch);
  FusedShape s;
  s.N = cfg->num_d2d; s.Dn = cfg->node_dim; s.De = cfg->edge_dim; s.F = cfg->feedback; s.CH = cfg->num_ch; s.S = cfg->stages; s.G = G;
  s.H1 = cfg->hidden[0]; s.H2 = cfg->hidden[1]; s.H3 = cfg->hidden[2];
  s.layer_K = lK.data(); s.layer_O = lO.data(); s.w_off = lw.data(); s.b_off = lb.data();
  s.n_layers = (int)lK.size(); s.n_params = off;
  FusedProgram* p = new FusedProgram();
  const int tg = fused_pick_tg(s, B, train);
  int rc = tg > 0 ? fused_build_program(s, tg, train, p) : 0;
  if (rc == 0 && tg > 0) {
    info8[0] = 1; info8[1] = tg; info8[2] = p->n_rows; info8[3] = (int)fused_smem_bytes(*p); info8[4] = p->n_ops;
    info8[5] = p->n_blocks; info8[6] = p->n_bias; info8[7] = p->n_tab;
  }
  delete p;
  return rc;
}

extern "C" int v2v_brain_update_target(v2v_brain* b, void* stream) {
  V2V_REQUIRE(b, "v2v_brain_update_target: null brain");
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->params[1], b->params[0], b->n_params * sizeof(float), cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
  return 0;
}

// ---------------------------------------------------------------------------
// forward / backward sequencing
// ---------------------------------------------------------------------------
static int agg_any(v2v_brain* b, const float* H, const uint32_t* mask, const float* adj, int transpose,
                   const float* addend, float* out, int B, void* stream) {
  if (mask) return v2v_agg_mask(H, mask, addend, out, B, b->N, b->F, V2V_F32, stream);
  return v2v_agg_dense(H, adj, addend, out, B, b->N, b->F, transpose, stream);
}

static int check_batch(const v2v_brain* b, int B, const char* who) {
  V2V_REQUIRE(b, "%s: null brain", who);
  V2V_REQUIRE(B >= 0 && B <= b->cfg.max_batch, "%s: batch %d exceeds max_batch %d", who, B, b->cfg.max_batch);
  return 0;
}

static int forward_impl(v2v_brain* b, const float* P, const float* node, const float* edge, const float* neigh,
                        const uint32_t* in_mask, const float* adj, int B, float* q_out, void* stream) {
  const int N = b->N, S = b->S, G = b->G, F = b->F;
  {
    const float* seg[3] = {node, edge, neigh};
    const int sw[3] = {b->Dn, b->De, F};
    const LayerDesc& L = b->layers[0];
    if (int rc = v2v_dense_fwd(neigh ? 3 : 2, seg, sw, P + L.w_off, L.K, P + L.b_off, b->h[0], B, N, G, F, S > 1, stream)) return rc;
    if (int rc = agg_any(b, b->h[0], in_mask, adj, 0, nullptr, b->agg[0], B, stream)) return rc;
  }
  for (int s = 1; s < S; ++s) {
    const float* seg[4] = {b->h[s - 1], node, edge, b->agg[s - 1]};
    const int sw[4] = {F, b->Dn, b->De, F};
    const LayerDesc& L = b->layers[s];
    if (int rc = v2v_dense_fwd(4, seg, sw, P + L.w_off, L.K, P + L.b_off, b->h[s], B, N, G, F, s < S - 1, stream)) return rc;
    if (int rc = agg_any(b, b->h[s], in_mask, adj, 0, nullptr, b->agg[s], B, stream)) return rc;
  }
  {
    const float* seg[3] = {node, b->h[S - 1], b->agg[S - 1]};
    const int sw[3] = {b->Dn, F, F};
    const LayerDesc& L = b->layers[S];
    if (int rc = v2v_dense_fwd(3, seg, sw, P + L.w_off, L.K, P + L.b_off, b->mlp[0], B, N, G, L.n_out, 1, stream)) return rc;
  }
  for (int j = 1; j < 4; ++j) {
    const LayerDesc& L = b->layers[S + j];
    const float* seg[1] = {b->mlp[j - 1]};
    const int sw[1] = {L.K};
    float* dst = (j == 3 && q_out) ? q_out : b->mlp[j];
    if (int rc = v2v_dense_fwd(1, seg, sw, P + L.w_off, L.K, P + L.b_off, dst, B, N, G, L.n_out, j < 3, stream)) return rc;
  }
  return 0;
}

extern "C" int v2v_brain_forward(v2v_brain* b, const float* node_dev, const float* edge_dev,
                                 const float* neighbor_dev, const uint32_t* in_mask_dev, const float* adj_dev, int B, int target,
                                 float* q_dev, void* stream) {
  if (int rc = check_batch(b, B, "v2v_brain_forward")) return rc;
  if (B == 0) return 0;
  V2V_REQUIRE(node_dev && edge_dev && q_dev, "v2v_brain_forward: null pointer");
  V2V_REQUIRE(in_mask_dev || adj_dev, "v2v_brain_forward: need in_mask or adj");
  if (b->bf16) {
    V2V_REQUIRE(in_mask_dev && !neighbor_dev, "v2v_brain_forward: the bf16 brain needs the binary adjacency masks and the "
                "reference's all-zero neighbour input");
    return tt_launch(*b->tt_plan, b->tt_plan_dev, b->params[target ? 1 : 0], node_dev, edge_dev, in_mask_dev, nullptr, q_dev,
                     nullptr, B, 0, (cudaStream_t)stream);
  }
  if (b->tc_capable && b->tc_mode > 0 && b->fused_enabled && in_mask_dev && !neighbor_dev &&
      (b->tc_mode >= 2 || ceil_div(B, b->tc_plan.TG) >= 2 * sm_count()))
    return tc_forward_launch(b->tc_plan, b->tc_plan_dev, b->params[target ? 1 : 0], b->tc_wimg, node_dev, edge_dev, in_mask_dev,
                             q_dev, B, (cudaStream_t)stream);
  if (use_fused(b, in_mask_dev, neighbor_dev)) {
    v2v_brain::FusedEntry* e = nullptr;
    if (int rc = fused_get(b, B, 0, &e)) return rc;
    return fused_launch(e->host, e->dev, b->params[target ? 1 : 0], node_dev, edge_dev, in_mask_dev, nullptr, nullptr, q_dev,
                        nullptr, nullptr, B, fused_grid(e->host, B), (cudaStream_t)stream);
  }
  return forward_impl(b, b->params[target ? 1 : 0], node_dev, edge_dev, neighbor_dev, in_mask_dev, adj_dev, B, q_dev, stream);
}

extern "C" int v2v_brain_forward_backward(v2v_brain* b, const float* node, const float* edge,
                                          const float* neigh, const uint32_t* in_mask, const uint32_t* out_mask,
                                          const float* adj, const float* y, int B, float* head_loss_dev,
                                          void* stream) {
  if (int rc = check_batch(b, B, "v2v_brain_forward_backward")) return rc;
  V2V_REQUIRE(B > 0, "v2v_brain_forward_backward: empty batch");
  V2V_REQUIRE(node && edge && y, "v2v_brain_forward_backward: null pointer");
  V2V_REQUIRE((in_mask && out_mask) || adj, "v2v_brain_forward_backward: need both masks, or adj");
  const bool use_mask = in_mask && out_mask;
  const uint32_t* im = use_mask ? in_mask : nullptr;
  const uint32_t* om = use_mask ? out_mask : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const int N = b->N, S = b->S, G = b->G, F = b->F, Dn = b->Dn, De = b->De;
  const float* P = b->params[0];
  float* Gd = b->params[2];
  float* hl = head_loss_dev ? head_loss_dev : b->head_loss;

  if (b->bf16) {
    V2V_REQUIRE(im && om && !neigh, "v2v_brain_forward_backward: the bf16 brain needs the binary adjacency masks and the "
                "reference's all-zero neighbour input");
    const int grid = tt_grid(*b->tt_plan, B);
    b->last_grid = grid;
    if (int rc = tt_launch(*b->tt_plan, b->tt_plan_dev, P, node, edge, im, y, nullptr, b->partial, B, 1, st)) return rc;
    if (b->defer_reduce) return 0;
    return fused_reduce_adam(b->partial, grid, tt_partial_stride((long)b->n_params), Gd, nullptr, nullptr, nullptr,
                             (long)b->n_params, N, hl, 0, 0.f, 0.f, 0.f, 0.f, 1.f, st);
  }
  if (use_fused(b, im, neigh)) {
    v2v_brain::FusedEntry* e = nullptr;
    if (int rc = fused_get(b, B, 1, &e)) return rc;
    const int grid = fused_grid(e->host, B);
    b->last_grid = grid;
    if (int rc = fused_launch(e->host, e->dev, P, node, edge, im, om, y, nullptr, b->partial, hl, B, grid, st)) return rc;
    if (b->defer_reduce) return 0;      // train_step fuses the reduction (gradient + per-head losses) with Adam
    return fused_reduce_adam(b->partial, grid, fused_partial_stride((long)b->n_params), Gd, nullptr, nullptr, nullptr,
                             (long)b->n_params, N, hl, 0, 0.f, 0.f, 0.f, 0.f, 1.f, st);
  }
  b->last_grid = 0;
  if (int rc = forward_impl(b, P, node, edge, neigh, im, adj, B, nullptr, stream)) return rc;
  V2V_CHECK_CUDA(cudaMemsetAsync(Gd, 0, b->n_params * sizeof(float), st));
  V2V_CHECK_CUDA(cudaMemsetAsync(hl, 0, N * sizeof(float), st));
  if (int rc = v2v_huber_loss_grad(b->mlp[3], y, b->dq, hl, B, N, b->CH, 1.f, stream)) return rc;

  // decision MLP, last layer first
  const float* dY = b->dq;
  for (int j = 3; j >= 1; --j) {
    const LayerDesc& L = b->layers[S + j];
    const float* seg[1] = {b->mlp[j - 1]};
    const int sw[1] = {L.K};
    if (int rc = v2v_dense_bwd_weight(1, seg, sw, dY, nullptr, Gd + L.w_off, L.K, Gd + L.b_off, B, N, G, L.n_out, stream)) return rc;
    if (int rc = v2v_dense_bwd_data(dY, nullptr, P + L.w_off, L.K, 0, L.K, b->dmlp[j], 0, 0, nullptr, b->mlp[j - 1], B, N, G,
                                    L.n_out, stream)) return rc;
    dY = b->dmlp[j];
  }
  float* dh = b->dh_a;
  float* dh_alt = b->dh_b;
  {
    const LayerDesc& L = b->layers[S];
    const float* seg[3] = {node, b->h[S - 1], b->agg[S - 1]};
    const int sw[3] = {Dn, F, F};
    if (int rc = v2v_dense_bwd_weight(3, seg, sw, dY, nullptr, Gd + L.w_off, L.K, Gd + L.b_off, B, N, G, L.n_out, stream)) return rc;
    if (int rc = v2v_dense_bwd_data(dY, nullptr, P + L.w_off, L.K, Dn, F, dh, Dn + F, F, b->dagg, nullptr, B, N, G, L.n_out,
                                    stream)) return rc;
    if (int rc = agg_any(b, b->dagg, om, adj, 1, dh, dh, B, stream)) return rc;
  }
  for (int s = S - 1; s >= 1; --s) {
    const LayerDesc& L = b->layers[s];
    const float* gate = (s < S - 1) ? b->h[s] : nullptr;
    const float* seg[4] = {b->h[s - 1], node, edge, b->agg[s - 1]};
    const int sw[4] = {F, Dn, De, F};
    if (int rc = v2v_dense_bwd_weight(4, seg, sw, dh, gate, Gd + L.w_off, L.K, Gd + L.b_off, B, N, G, F, stream)) return rc;
    if (int rc = v2v_dense_bwd_data(dh, gate, P + L.w_off, L.K, 0, F, dh_alt, F + Dn + De, F, b->dagg, nullptr, B, N, G, F,
                                    stream)) return rc;
    if (int rc = agg_any(b, b->dagg, om, adj, 1, dh_alt, dh_alt, B, stream)) return rc;
    std::swap(dh, dh_alt);
  }
  {
    const LayerDesc& L = b->layers[0];
    const float* gate = (S > 1) ? b->h[0] : nullptr;
    const float* seg[3] = {node, edge, neigh};
    const int sw[3] = {Dn, De, F};
    if (int rc = v2v_dense_bwd_weight(neigh ? 3 : 2, seg, sw, dh, gate, Gd + L.w_off, L.K, Gd + L.b_off, B, N, G, F, stream)) return rc;
  }
  return 0;
}

extern "C" int v2v_brain_apply_adam(v2v_brain* b, float grad_scale, void* stream) {
  V2V_REQUIRE(b, "v2v_brain_apply_adam: null brain");
  b->iterations += 1;
  return v2v_adam_step(b->params[0], b->params[2], b->params[3], b->params[4], (long)b->n_params, b->iterations,
                       b->cfg.lr, b->cfg.beta1, b->cfg.beta2, b->cfg.eps, grad_scale, stream);
}

extern "C" int v2v_brain_train_step(v2v_brain* b, const float* node, const float* edge,
                                    const float* neigh, const uint32_t* in_mask, const uint32_t* out_mask, const float* adj,
                                    const float* y, int B, float* head_loss_dev, void* stream) {
  b->defer_reduce = true;
  int rc = v2v_brain_forward_backward(b, node, edge, neigh, in_mask, out_mask, adj, y, B, head_loss_dev, stream);
  b->defer_reduce = false;
  if (rc) return rc;
  if (b->last_grid > 0) {             // fused path: partial reduction + Keras-Adam in one kernel
    b->iterations += 1;
    return fused_reduce_adam(b->partial, b->last_grid, fused_partial_stride((long)b->n_params), b->params[2], b->params[0],
                             b->params[3], b->params[4], (long)b->n_params, b->N, head_loss_dev ? head_loss_dev : b->head_loss,
                             b->iterations, b->cfg.lr, b->cfg.beta1, b->cfg.beta2, b->cfg.eps, 1.f, (cudaStream_t)stream);
  }
  return v2v_brain_apply_adam(b, 1.f, stream);
}

// Data-parallel train step: local fwd + Huber + bwd, then ONE kernel that reduces the per-CTA partials,
// exchanges the gradient (and the per-head losses) with all peers over NVLink and applies Adam (comm.cu).
extern "C" int v2v_comm_allreduce_adam_ex(struct v2v_comm* c, const float* partial_dev, int n_cta, long row_stride, long n_adam,
                                          long n_src, const float* extra_dev, int n_extra, float* grad_dev, float* p_dev,
                                          float* m_dev, float* v_dev, float* extra_out_dev, int t, float lr, float beta1,
                                          float beta2, float eps, void* stream);

extern "C" int v2v_brain_train_step_dp(v2v_brain* b, struct v2v_comm* comm, const float* node, const float* edge,
                                       const float* neigh, const uint32_t* in_mask, const uint32_t* out_mask,
                                       const float* adj, const float* y, int B, float* head_loss_dev, void* stream) {
  V2V_REQUIRE(b && comm, "v2v_brain_train_step_dp: null argument");
  float* hl = head_loss_dev ? head_loss_dev : b->head_loss;
  b->defer_reduce = true;
  int rc = v2v_brain_forward_backward(b, node, edge, neigh, in_mask, out_mask, adj, y, B, hl, stream);
  b->defer_reduce = false;
  if (rc) return rc;
  b->iterations += 1;
  if ((b->iterations & 31) == 0)        // a timed-out exchange skipped its update: surface it, never train on diverged replicas
    if (int rc2 = v2v_comm_poll_error(comm, stream)) return rc2;
  const long np = (long)b->n_params;
  if (b->last_grid > 0)     // fused path: the per-head losses are the tail columns of the per-CTA partial rows
    return v2v_comm_allreduce_adam_ex(comm, b->partial, b->last_grid, fused_partial_stride(np), np, np + b->N, nullptr, 0,
                                      b->params[2], b->params[0], b->params[3], b->params[4], hl, b->iterations, b->cfg.lr,
                                      b->cfg.beta1, b->cfg.beta2, b->cfg.eps, stream);
  return v2v_comm_allreduce_adam_ex(comm, b->params[2], 1, np, np, np, hl, b->N, b->params[2], b->params[0], b->params[3],
                                    b->params[4], hl, b->iterations, b->cfg.lr, b->cfg.beta1, b->cfg.beta2, b->cfg.eps, stream);
}

// ---------------------------------------------------------------------------
// host-buffer entry points (the reference-facing plugin path)
// ---------------------------------------------------------------------------
static int stage_inputs(v2v_brain* b, const float* node_host, const float* edge_host, const float* neigh_host,
                        const float* adj_host,
                        int B, bool need_out_mask, bool* weighted, cudaStream_t st) {
  const size_t rows = (size_t)B * b->N;
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_node, node_host, rows * b->Dn * sizeof(float), cudaMemcpyHostToDevice, st));
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_edge, edge_host, rows * b->De * sizeof(float), cudaMemcpyHostToDevice, st));
  if (neigh_host)
    V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_neigh, neigh_host, rows * b->F * sizeof(float), cudaMemcpyHostToDevice, st));
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_adj, adj_host, rows * b->N * sizeof(float), cudaMemcpyHostToDevice, st));
  if (int rc = v2v_adj_pack_masks(b->st_adj, B, b->N, b->st_in_mask, need_out_mask ? b->st_out_mask : nullptr, b->st_flag, st))
    return rc;
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_flag_host, b->st_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  *weighted = (*b->st_flag_host != 0);
  return 0;
}

extern "C" int v2v_brain_predict_host(v2v_brain* b, const float* node_host, const float* edge_host,
                                      const float* neigh_host, const float* adj_host, int B, int target, float* q_host, void* stream) {
  if (int rc = check_batch(b, B, "v2v_brain_predict_host")) return rc;
  if (B == 0) return 0;
  V2V_REQUIRE(node_host && edge_host && adj_host && q_host, "v2v_brain_predict_host: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool weighted = false;
  if (int rc = stage_inputs(b, node_host, edge_host, neigh_host, adj_host, B, false, &weighted, st)) return rc;
  if (int rc = v2v_brain_forward(b, b->st_node, b->st_edge, neigh_host ? b->st_neigh : nullptr,
                                 weighted ? nullptr : b->st_in_mask, b->st_adj, B, target, b->st_q, stream)) return rc;
  V2V_CHECK_CUDA(cudaMemcpyAsync(q_host, b->st_q, (size_t)B * b->N * b->CH * sizeof(float), cudaMemcpyDeviceToHost, st));
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int v2v_brain_train_host(v2v_brain* b, const float* node_host, const float* edge_host,
                                    const float* neigh_host, const float* adj_host, const float* y_host, int B, float* head_loss_host,
                                    void* stream) {
  if (int rc = check_batch(b, B, "v2v_brain_train_host")) return rc;
  V2V_REQUIRE(B > 0, "v2v_brain_train_host: empty batch");
  V2V_REQUIRE(node_host && edge_host && adj_host && y_host, "v2v_brain_train_host: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_y, y_host, (size_t)B * b->N * b->CH * sizeof(float), cudaMemcpyHostToDevice, st));
  bool weighted = false;
  if (int rc = stage_inputs(b, node_host, edge_host, neigh_host, adj_host, B, true, &weighted, st)) return rc;
  if (int rc = v2v_brain_train_step(b, b->st_node, b->st_edge, neigh_host ? b->st_neigh : nullptr,
                                    weighted ? nullptr : b->st_in_mask,
                                    weighted ? nullptr : b->st_out_mask, b->st_adj, b->st_y, B, b->head_loss, stream))
    return rc;
  if (head_loss_host)
    V2V_CHECK_CUDA(cudaMemcpyAsync(head_loss_host, b->head_loss, b->N * sizeof(float), cudaMemcpyDeviceToHost, st));
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------
// strided-view entry points: gather + convert on the host worker pool, pipelined H2D, one synchronisation
// ---------------------------------------------------------------------------
static int stage_views(v2v_brain* b, const v2v_host_view* node, int n_node, const v2v_host_view* edge, int n_edge,
                       const v2v_host_view* neigh, int n_neigh, const v2v_host_view* adj, int n_adj,
                       const v2v_host_view* y, int n_y, int B, bool need_out_mask, bool* weighted, bool* has_neigh,
                       cudaStream_t st, const char* what) {
  const long rows = (long)B * b->N;
  V2V_REQUIRE(n_node > 0 && n_edge > 0 && n_adj > 0, "%s: node, edge and adjacency views are required", what);
  if (int rc = host_stage_validate(node, n_node, rows * b->Dn, what)) return rc;
  if (int rc = host_stage_validate(edge, n_edge, rows * b->De, what)) return rc;
  if (int rc = host_stage_validate(neigh, n_neigh, rows * b->F, what)) return rc;
  if (int rc = host_stage_validate(adj, n_adj, rows * b->N, what)) return rc;
  if (int rc = host_stage_validate(y, n_y, rows * b->CH, what)) return rc;
  // the workers build the bit masks themselves when the adjacency is one whole [B*N][N] view and N <= 32
  const bool host_pack = b->N <= 32 && n_adj == 1 && adj[0].rows == rows && adj[0].cols == b->N && adj[0].dst_off == 0 &&
                         adj[0].dst_row_stride == b->N;
  const size_t mask_bytes = (size_t)rows * sizeof(uint32_t);
  HostStageTensor t[5];
  int n = 0, i_neigh = -1, i_adj = -1;
  auto plain = [&](const v2v_host_view* v, int nv, size_t pin_off, float* dev, size_t elems) {
    HostStageTensor T{};
    T.views = v; T.n_views = nv; T.pinned = b->pin + pin_off; T.device = dev; T.bytes = elems * sizeof(float);
    return T;
  };
  t[n++] = plain(node, n_node, b->pin_node, b->st_node, (size_t)rows * b->Dn);
  t[n++] = plain(edge, n_edge, b->pin_edge, b->st_edge, (size_t)rows * b->De);
  if (n_y > 0) t[n++] = plain(y, n_y, b->pin_y, b->st_y, (size_t)rows * b->CH);
  if (n_neigh > 0) {                 // the reference always feeds zeros here (:478, :589): only non-zero data travels
    i_neigh = n;
    t[n] = plain(neigh, n_neigh, b->pin_neigh, b->st_neigh, (size_t)rows * b->F);
    t[n].check = kCheckNonzero; t[n].copy_if = kFlagNonzero;
    ++n;
  }
  i_adj = n;
  t[n] = plain(adj, n_adj, b->pin_adj, b->st_adj, (size_t)rows * b->N);
  t[n].check = kCheckBinary;
  if (host_pack) {
    t[n].copy_if = kFlagNonbinary;   // the dense matrix is only needed by the weighted-adjacency kernels
    t[n].pack_N = b->N;
    t[n].pin_in_mask = reinterpret_cast<uint32_t*>(b->pin + b->pin_im);
    t[n].pin_out_mask = reinterpret_cast<uint32_t*>(b->pin + b->pin_om);
    t[n].dev_in_mask = b->st_in_mask;
    t[n].dev_out_mask = need_out_mask ? b->st_out_mask : nullptr;
    t[n].mask_bytes = mask_bytes;
  }
  ++n;
  // Small batches: every H2D copy costs ~5 us of driver time whatever its size, so the adjacent block node | edge | y |
  // masks travels as ONE copy once all of it is staged; large batches keep one copy per tensor, issued the moment the
  // tensor is complete (the copy of the features overlaps the bit-packing of the adjacency).
  const size_t merged_floats = b->pin_om + (size_t)rows * ceil_div(b->N, 32) - b->pin_node;
  const bool merge = host_pack && merged_floats * sizeof(float) <= 512 * 1024;
  if (merge) {
    for (int i = 0; i < n; ++i)
      if (i != i_neigh) { t[i].device = nullptr; t[i].dev_in_mask = nullptr; t[i].dev_out_mask = nullptr; }
  } else if (host_pack) {
    // large batches: two copies instead of five (every cudaMemcpyAsync costs ~7 us of driver time).  node | edge | y are
    // adjacent in both blocks and staged first: the last of them ships all three, while the workers are still packing
    // the adjacency; the two mask orientations are adjacent too and ship together once packed.
    const int last_plain = (n_y > 0) ? 2 : 1;                 // t[0] node, t[1] edge, t[2] y (training)
    const size_t end = (n_y > 0) ? b->pin_y + (size_t)rows * b->CH : b->pin_edge + (size_t)rows * b->De;
    for (int i = 0; i < last_plain; ++i) t[i].device = nullptr;
    t[last_plain].device = b->st_block + b->pin_node;
    t[last_plain].h2d_from = b->pin + b->pin_node;
    t[last_plain].bytes = (end - b->pin_node) * sizeof(float);
    if (need_out_mask) {
      t[i_adj].dev_out_mask = nullptr;
      t[i_adj].mask_bytes = (b->pin_om + (size_t)rows * ceil_div(b->N, 32) - b->pin_im) * sizeof(float);
    }
  }
  int flags[kHostStageMaxTensors] = {0};
  if (int rc = host_stage_run(t, n, st, flags)) return rc;
  if (merge) {
    V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_block + b->pin_node, b->pin + b->pin_node, merged_floats * sizeof(float), cudaMemcpyHostToDevice, st));
    if (flags[i_adj] & kFlagNonbinary)           // weighted adjacency: the dense matrix was staged into the pinned block by the fallback
      V2V_CHECK_CUDA(cudaMemcpyAsync(b->st_adj, b->pin + b->pin_adj, (size_t)rows * b->N * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  *weighted = (flags[i_adj] & kFlagNonbinary) != 0;
  *has_neigh = i_neigh >= 0 && (flags[i_neigh] & kFlagNonzero) != 0;
  if (!*weighted && !host_pack)
    if (int rc = v2v_adj_pack_masks(b->st_adj, B, b->N, b->st_in_mask, need_out_mask ? b->st_out_mask : nullptr, nullptr, st))
      return rc;
  return 0;
}

extern "C" int v2v_host_stage_threads(void) { return host_stage_threads(); }

// mode 0: never, 1: automatic (batches that give every SM at least two 128-row tiles (one per tile slot)), 2: whenever the brain is capable
extern "C" int v2v_brain_set_tensor_core(v2v_brain* b, int mode) {
  V2V_REQUIRE(b, "v2v_brain_set_tensor_core: null brain");
  V2V_REQUIRE(mode >= 0 && mode <= 2, "v2v_brain_set_tensor_core: mode %d outside [0,2]", mode);
  V2V_REQUIRE(mode == 0 || b->tc_capable, "v2v_brain_set_tensor_core: this configuration has no tensor-core forward "
              "(needs shared weights, N <= 32, feedback width a multiple of 8)");
  b->tc_mode = mode;
  return 0;
}
// debugging aid: runs the tensor-core forward and dumps the raw accumulator [128][Npad] of one layer of the first tile
extern "C" int v2v_brain_tc_debug(v2v_brain* b, const float* node, const float* edge, const uint32_t* in_mask, int B,
                                  int layer, float* q_dev, float* dbg_dev, int* npad_out, void* stream) {
  V2V_REQUIRE(b && b->tc_capable && layer >= -2 && layer < b->tc_plan.n_layers, "v2v_brain_tc_debug: bad arguments");
  if (npad_out) *npad_out = layer >= 0 ? b->tc_plan.layers[layer].Npad : b->tc_plan.n_layers;
  return tc_forward_launch(b->tc_plan, b->tc_plan_dev, b->params[0], b->tc_wimg, node, edge, in_mask, q_dev, B,
                           (cudaStream_t)stream, dbg_dev, layer);
}
extern "C" int v2v_brain_tensor_core_info(const v2v_brain* b, int* info4) {
  V2V_REQUIRE(b && info4, "v2v_brain_tensor_core_info: null argument");
  info4[0] = b->tc_capable ? 1 : 0; info4[1] = b->tc_mode; info4[2] = b->tc_capable ? b->tc_plan.TG : 0;
  info4[3] = b->tc_capable ? b->tc_plan.smem_bytes : 0;
  return 0;
}

// host-only pieces of the staging path, exported so that they can be checked without a device
extern "C" int v2v_host_gather(const v2v_host_view* views, int n_views, float* dst, long dst_elems, int check, int* flags_out) {
  V2V_REQUIRE(dst || dst_elems == 0, "v2v_host_gather: null destination");
  if (int rc = host_stage_validate(views, n_views, dst_elems, "v2v_host_gather")) return rc;
  HostStageTensor T{};
  T.views = views; T.n_views = n_views; T.pinned = dst; T.check = check;
  return host_stage_run(&T, 1, nullptr, flags_out);
}

extern "C" int v2v_host_pack_adjacency(const v2v_host_view* adj_view, int B, int N, uint32_t* in_mask, uint32_t* out_mask,
                                       int* flags_out) {
  V2V_REQUIRE(adj_view && in_mask && out_mask && B >= 0 && N >= 1 && N <= 32, "v2v_host_pack_adjacency: bad arguments (N <= 32)");
  V2V_REQUIRE(adj_view->rows == (long)B * N && adj_view->cols == N, "v2v_host_pack_adjacency: the view must be [B*N][N]");
  V2V_REQUIRE(adj_view->dtype == V2V_F32 || adj_view->dtype == V2V_F64, "v2v_host_pack_adjacency: dtype");
  HostStageTensor T{};
  T.views = adj_view; T.n_views = 1; T.check = kCheckBinary; T.copy_if = kFlagNonbinary;
  T.pack_N = N; T.pin_in_mask = in_mask; T.pin_out_mask = out_mask;
  int f = 0;
  if (int rc = host_stage_run(&T, 1, nullptr, &f)) return rc;
  if (flags_out) *flags_out = f;
  return 0;
}

extern "C" int v2v_brain_predict_views(v2v_brain* b, const v2v_host_view* node, int n_node, const v2v_host_view* edge,
                                       int n_edge, const v2v_host_view* neigh, int n_neigh, const v2v_host_view* adj,
                                       int n_adj, int B, int target, float* q_host, void* stream) {
  if (int rc = check_batch(b, B, "v2v_brain_predict_views")) return rc;
  if (B == 0) return 0;
  V2V_REQUIRE(q_host, "v2v_brain_predict_views: null output");
  cudaStream_t st = (cudaStream_t)stream;
  bool weighted = false, has_neigh = false;
  if (int rc = stage_views(b, node, n_node, edge, n_edge, neigh, n_neigh, adj, n_adj, nullptr, 0, B, false, &weighted,
                           &has_neigh, st, "v2v_brain_predict_views")) return rc;
  if (int rc = v2v_brain_forward(b, b->st_node, b->st_edge, has_neigh ? b->st_neigh : nullptr,
                                 weighted ? nullptr : b->st_in_mask, b->st_adj, B, target, b->st_q, stream)) return rc;
  const size_t qb = (size_t)B * b->N * b->CH * sizeof(float);
  V2V_CHECK_CUDA(cudaMemcpyAsync(b->pin + b->pin_q, b->st_q, qb, cudaMemcpyDeviceToHost, st));
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  memcpy(q_host, b->pin + b->pin_q, qb);
  return 0;
}

static int train_views_impl(v2v_brain* b, struct v2v_comm* comm, const v2v_host_view* node, int n_node,
                            const v2v_host_view* edge, int n_edge, const v2v_host_view* neigh, int n_neigh,
                            const v2v_host_view* adj, int n_adj, const v2v_host_view* y, int n_y, int B,
                            float* head_loss_host, void* stream) {
  if (int rc = check_batch(b, B, "v2v_brain_train_views")) return rc;
  V2V_REQUIRE(B > 0, "v2v_brain_train_views: empty batch");
  V2V_REQUIRE(n_y > 0, "v2v_brain_train_views: target views are required");
  cudaStream_t st = (cudaStream_t)stream;
  bool weighted = false;
  static const bool trace = getenv("V2V_HOST_TRACE") != nullptr;   // host-side phase timing, printed every 100 calls
  static double acc[3] = {0, 0, 0};
  static int calls = 0;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
  const double t0 = trace ? now() : 0;
  bool has_neigh = false;
  if (int rc = stage_views(b, node, n_node, edge, n_edge, neigh, n_neigh, adj, n_adj, y, n_y, B, true, &weighted, &has_neigh,
                           st, "v2v_brain_train_views")) return rc;
  const double t1 = trace ? now() : 0;
  const float* ng = has_neigh ? b->st_neigh : nullptr;
  const uint32_t* im = weighted ? nullptr : b->st_in_mask;
  const uint32_t* om = weighted ? nullptr : b->st_out_mask;
  // The per-head losses are N floats written once by the reduction kernel: on the fused paths it stores them straight into
  // the pinned (device-addressable) host block, which saves the device-to-host copy and its latency at the end of the
  // call.  The layered path accumulates them with atomics and keeps the device buffer + copy.
  const bool direct = b->bf16 || use_fused(b, im, ng);
  float* hl_out = direct ? b->pin + b->pin_hl : b->head_loss;
  if (int rc = comm ? v2v_brain_train_step_dp(b, comm, b->st_node, b->st_edge, ng, im, om, b->st_adj, b->st_y, B, hl_out, stream)
                    : v2v_brain_train_step(b, b->st_node, b->st_edge, ng, im, om, b->st_adj, b->st_y, B, hl_out, stream))
    return rc;
  if (!direct)
    V2V_CHECK_CUDA(cudaMemcpyAsync(b->pin + b->pin_hl, b->head_loss, b->N * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (comm) if (int rc = v2v_comm_poll_error(comm, stream)) return rc;
  const double t2 = trace ? now() : 0;
  V2V_CHECK_CUDA(cudaStreamSynchronize(st));
  if (comm) if (int rc = v2v_comm_poll_result(comm)) return rc;
  if (trace) {
    const double t3 = now();
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2;
    if (++calls % 100 == 0) {
      fprintf(stderr, "[v2v host trace] train_views B=%d: stage+enqueue H2D %.1f us, enqueue step %.1f us, wait %.1f us (avg of 100)\n",
              B, acc[0] / 100, acc[1] / 100, acc[2] / 100);
      acc[0] = acc[1] = acc[2] = 0;
    }
  }
  if (head_loss_host) memcpy(head_loss_host, b->pin + b->pin_hl, b->N * sizeof(float));
  return 0;
}

extern "C" int v2v_brain_train_views(v2v_brain* b, const v2v_host_view* node, int n_node, const v2v_host_view* edge,
                                     int n_edge, const v2v_host_view* neigh, int n_neigh, const v2v_host_view* adj,
                                     int n_adj, const v2v_host_view* y, int n_y, int B, float* head_loss_host,
                                     void* stream) {
  return train_views_impl(b, nullptr, node, n_node, edge, n_edge, neigh, n_neigh, adj, n_adj, y, n_y, B, head_loss_host, stream);
}

// data-parallel variant: this rank's rows, gradients (and the per-head losses) exchanged by v2v_comm_allreduce_adam
extern "C" int v2v_brain_train_views_dp(v2v_brain* b, struct v2v_comm* comm, const v2v_host_view* node, int n_node,
                                        const v2v_host_view* edge, int n_edge, const v2v_host_view* neigh, int n_neigh,
                                        const v2v_host_view* adj, int n_adj, const v2v_host_view* y, int n_y, int B,
                                        float* head_loss_host, void* stream) {
  V2V_REQUIRE(comm, "v2v_brain_train_views_dp: null communicator");
  return train_views_impl(b, comm, node, n_node, edge, n_edge, neigh, n_neigh, adj, n_adj, y, n_y, B, head_loss_host, stream);
}
