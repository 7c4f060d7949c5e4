// Tensor-core (tcgen05 / TMEM) forward of the shared-weight brain: declarations (see tc_forward.cu).
#pragma once
#include "v2v_common.cuh"

namespace v2v {

constexpr int kTcMaxLayers = 12;       // up to 8 combine stages + the 4-layer decision MLP
constexpr int kTcMaxK = 96;            // padded contraction length of one layer
constexpr int kTcEpiThreads = 512;      // 16 epilogue warps: 4 TMEM lane quarters x 4 column quarters
constexpr int kTcThreads = kTcEpiThreads + 32;   // + one MMA-issuing warp
constexpr int kTcRows = 128;           // rows of a tile = TMEM lanes = UMMA M

struct TcLayer {
  int Kpad, Npad;        // contraction length (multiple of 8) and output columns (multiple of 16) as issued to the MMA
  int N;                 // true output columns (row length of W in the parameter buffer)
  int relu;
  int pw_off, pb_off;    // parameter offsets (floats) of W[K_true][N] and bias[N]
  int w_off;             // shared-memory float offset of the layer's hi planes (lo planes follow at + Kpad * Npad)
  int bias_off;          // shared-memory float offset of the zero-padded bias
  int a_src;             // 0: x0 planes, 1: [h | agg | x0] planes (split on the fly), 2: the previous layer's epilogue output
  int out_kind;          // 0: h planes (fp32 + hi/lo) then aggregation, 1: next layer's operand (hi/lo, in TMEM), 2: Q to global
  int dcol;              // accumulator base column inside the tile slot's 256 TMEM columns
  int acol;              // a_src == 2: base column of the operand's hi half (lo half at acol + Kpad)
  short kmap[kTcMaxK];   // contraction index -> row of W in the parameter buffer (-1: zero row)
};

struct TcPlan {
  int n_layers, S;
  int N, TG, Dn, De, F, CH;
  int x_planes;          // planes of x0 = [node (padded to 4) | edge (padded to 4)]
  int dn_pad;
  int w_floats;          // floats of all hi+lo weight planes
  int bias_floats;
  int stage_planes;      // planes of the largest MMA operand
  int smem_bytes;
  TcLayer layers[kTcMaxLayers];
};

struct TcShape {
  int N, Dn, De, F, CH, S, H1, H2, H3;
  const size_t* w_off;   // per layer parameter offsets (stages then MLP), floats
  const size_t* b_off;
};

// Builds the plan; returns non-zero (with last_error) when the configuration is outside the tensor-core path.
int tc_build_plan(const TcShape& s, TcPlan* out);
int tc_grid(const TcPlan& p, int B);
// wimg: device scratch of (w_floats + bias_floats) floats for the staged weight image (rebuilt by every call)
int tc_forward_launch(const TcPlan& plan_host, const TcPlan* plan_dev, const float* params, float* wimg, const float* node,
                      const float* edge, const uint32_t* in_mask, float* q_out, int B, cudaStream_t st, float* dbg = nullptr,
                      int dbg_layer = -1);

}  // namespace v2v
