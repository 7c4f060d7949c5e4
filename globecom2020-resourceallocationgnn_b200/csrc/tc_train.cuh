// bf16 tensor-core (tcgen05 / TMEM) forward + Huber + backward of the shared-weight brain: declarations (see tc_train.cu).
#pragma once
#include "v2v_common.cuh"

namespace v2v {

constexpr int kTtRows = 128;                        // rows of a tile = TMEM lanes = UMMA M
constexpr int kTtPlaneBytes = kTtRows * 16;         // one operand plane: 128 rows x 8 bf16
constexpr int kTtEpiThreads = 384;                  // 12 epilogue warps: 4 TMEM lane quarters x 3 column groups (13 warps in
                                                    // all: at most 4 per SM sub-partition, so 128 registers per thread)
constexpr int kTtCQ = kTtEpiThreads / kTtRows;      // column groups: chunk c of a step belongs to group c % kTtCQ
constexpr int kTtThreads = kTtEpiThreads + 32;      // + one MMA-issuing warp
constexpr int kTtMaxLayers = 8;                     // <= 3 combine stages + the 4-layer decision MLP (+1 spare)
constexpr int kTtMaxSteps = 24;
constexpr int kTtMaxMma = 128;
constexpr int kTtMaxK = 128;
constexpr int kTtMaxBlocks = 12;                    // column blocks of the weight-gradient accumulators
constexpr int kTtWorkCols = 128;                    // TMEM columns [0, 128): accumulator of the current contraction
constexpr int kTtPartialTail = 32;                  // per-head Huber sums behind the gradient in a CTA's partial row

enum TtStepKind : int {
  TT_PLANES = 0,       // accumulator columns -> (+ bias) (relu) (relu gate of a saved activation) -> bf16 operand planes
  TT_Q = 1,            // output layer: bias -> Q (global) or Huber loss + dq plane
  TT_WGRAD = 2,        // weight-gradient contractions of the whole tile (no epilogue; accumulators persist in TMEM)
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
  int chunk0;           // first 8-column chunk of the accumulator that is read
  int n_chunks;         // 8-column chunks that are written out (padding chunks are skipped)
  int out_plane;        // first destination plane (bf16 operand planes; unified index: activations, then gradients)
  int gate_plane;       // relu gate source planes (-1: none)
  int bias_off;         // float offset into the bias image (-1: no bias)
  int relu;
  int post0, n_post;    // MMAs issued right AFTER this step's commit (weight-gradient chains whose operands are complete):
                        // the tensor pipe works on them while the epilogue warps process the step
  int pad;
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

struct TtMmaRt {        // the same, ready to issue: descriptors with the CTA's shared-memory base folded in (32 bytes)
  unsigned long long a_desc, b_desc;
  uint32_t idesc, dcol, acc, pad;
};

struct TtPlan {
  int N, TG, Dn, De, F, CH, S;
  int n_layers, n_steps_fwd, n_steps_train, n_mma, n_blocks;
  int XP, FP;                        // planes of x0 and of one h / agg tensor
  int x_planes, dz_planes;           // activation planes, gradient planes (the latter follow the former in memory)
  int plane_dq;                      // plane of dq (unified index)
  int plane_adj;                     // first plane of the tile's block-diagonal adjacency operand (16 planes)
  int w_elems, bias_floats;          // bf16 weight image elements, fp32 bias image floats
  int off_w, off_bias, off_planes, off_misc, off_tab, smem_bytes;   // byte offsets from the smem base
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
int tt_set_trace(long long* dev_buf);      // dev_buf: 2 * kTtMaxSteps * 8 clock stamps of CTA 0's first two tiles, or nullptr
inline long tt_partial_stride(long n_params) { return n_params + kTtPartialTail; }
// One launch.  train = 0: q_out [B][N][CH] fp32.   train = 1: partial_dev [grid][n_params + 32] (gradient + per-head
// Huber sums).  Only the in_mask orientation of the adjacency is needed (the transposed aggregation reads the same
// block-diagonal operand MN-major).
int tt_launch(const TtPlan& plan_host, const TtPlan* plan_dev, const float* params, const float* node, const float* edge,
              const uint32_t* in_mask, const float* y, float* q_out, float* partial_dev, int B, int train, cudaStream_t st);

}  // namespace v2v
