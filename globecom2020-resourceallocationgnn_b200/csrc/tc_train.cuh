// bf16 tensor-core (tcgen05 / TMEM) forward + Huber + backward of the shared-weight brain: declarations (see tc_train.cu).
#pragma once
#include "v2v_common.cuh"

namespace v2v {

constexpr int kTtRows = 128;                        // rows of a tile = TMEM lanes = UMMA M
constexpr int kTtPlaneBytes = kTtRows * 16;         // one operand plane: 128 rows x 8 bf16
constexpr int kTtEpiThreads = 512;                  // 16 epilogue warps: 4 TMEM lane quarters x 4 column quarters
constexpr int kTtThreads = kTtEpiThreads + 32;      // + one MMA-issuing warp
constexpr int kTtMaxLayers = 8;                     // <= 3 combine stages + the 4-layer decision MLP (+1 spare)
constexpr int kTtMaxSteps = 16;
constexpr int kTtMaxMma = 96;
constexpr int kTtMaxK = 128;
constexpr int kTtMaxBlocks = 12;                    // column blocks of the weight-gradient accumulators
constexpr int kTtWorkCols = 128;                    // TMEM columns [0, 128): accumulator of the current contraction
constexpr int kTtPartialTail = 32;                  // per-head Huber sums behind the gradient in a CTA's partial row

enum TtStepKind : int {
  TT_COMBINE = 0,      // GNNLayer.call forward: bias, (relu), fp32 copy -> neighbour aggregation -> bf16 h and agg planes
  TT_MLP = 1,          // Dense forward: bias, relu -> bf16 planes
  TT_Q = 2,            // output layer: bias -> Q (global) or Huber loss + dq plane
  TT_DGRAD_MLP = 3,    // data gradient of a Dense layer: relu gate of the producer -> bf16 dz planes
  TT_DGRAD_AGG = 4,    // data gradient [dh | dagg]: dz = gate * (dh + Agg^T dagg) -> bf16 dz planes
  TT_WGRAD = 5,        // weight-gradient contractions of the whole tile (no epilogue; accumulators persist in TMEM)
};

struct TtMma {          // one tcgen05.mma.kind::f16: byte offsets are relative to the dynamic shared-memory base
  uint32_t a_off, a_lbo, a_sbo;
  uint32_t b_off, b_lbo, b_sbo;
  uint32_t idesc;
  uint32_t dcol;        // TMEM column of the accumulator
  uint32_t acc;         // 0: overwrite, 1: accumulate, 2: accumulate except on the CTA's first tile
  uint32_t pad[3];
};

struct TtStep {
  int kind;
  int mma0, n_mma;
  int npad;             // accumulator columns
  int n_chunks;         // 8-column chunks that are written out (padding chunks are skipped)
  int out_plane;        // first destination plane (bf16 operand planes; unified index: activations, then gradients)
  int out2_plane;       // TT_COMBINE: aggregated planes
  int gate_plane;       // relu gate source planes (-1: none)
  int bias_off;         // float offset into the bias image
  int relu;
  int pad[2];
};

struct TtLayerImg {
  int Kpad, Npad, No;   // image K (multiple of 16), rows per k-plane (multiple of 16), true output columns
  int w_off;            // element offset of the layer inside the bf16 weight image
  int bias_off;         // float offset inside the bias image
  int pw_off, pb_off;   // parameter offsets (floats)
  int pad;
  short kmap[kTtMaxK];  // image k -> row of W in the parameter buffer (-1: zero)
};

struct TtBlock {        // one column block of a weight-gradient accumulator: lane l's columns [col, col + n) go to dst[l]
  int tcol, n;
  int dst[kTtRows];     // float offset into the parameter vector (-1: nothing)
};

struct TtPlan {
  int N, TG, Dn, De, F, CH, S;
  int n_layers, n_steps_fwd, n_steps_train, n_mma, n_blocks;
  int XP, FP;                        // planes of x0 and of one h / agg tensor
  int x_planes, dz_planes;           // activation planes, gradient planes (the latter follow the former in memory)
  int plane_dq;                      // plane of dq (unified index)
  int w_elems, bias_floats;          // bf16 weight image elements, fp32 bias image floats
  int off_w, off_bias, off_planes, off_scr, off_mask, off_misc, off_tab, smem_bytes;   // byte offsets from the smem base
  int ones_feature;                  // x0 feature that is 1 for valid rows (bias gradients come out of the same contraction)
  long n_params;
  TtLayerImg layers[kTtMaxLayers];
  TtStep steps[kTtMaxSteps];
  TtMma mma[kTtMaxMma];
  TtBlock blocks[kTtMaxBlocks];
};

struct TtShape {
  int N, Dn, De, F, CH, S, H1, H2, H3;
  const size_t* w_off;
  const size_t* b_off;
  size_t n_params;
};

// Builds the plan; non-zero (with last_error) when the configuration is outside this path.
int tt_build_plan(const TtShape& s, TtPlan* out);
int tt_grid(const TtPlan& p, int B);
inline long tt_partial_stride(long n_params) { return n_params + kTtPartialTail; }
// One call = weight staging (fp32 master weights -> bf16 image) + the tile kernel.
//   train = 0: q_out [B][N][CH] fp32.   train = 1: partial_dev [grid][n_params + 32] (gradient + per-head Huber sums).
int tt_launch(const TtPlan& plan_host, const TtPlan* plan_dev, const float* params, void* wimg, const float* node,
              const float* edge, const uint32_t* in_mask, const uint32_t* out_mask, const float* y, float* q_out,
              float* partial_dev, int B, int train, cudaStream_t st);

}  // namespace v2v
