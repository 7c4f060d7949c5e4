// Whole-network fused kernel: shared declarations (see fused.cu).
#pragma once
#include "v2v_common.cuh"

namespace v2v {

enum FusedOpType : int {
  FOP_GEMM = 0,    // out[o][r] = act(bias[o] + sum_k W[k][o] * in[k][r])
  FOP_AGG = 1,     // out[f][g,m] = (add) + sum_n A[g][n][m] in[f][g,n]       (transposed = 1: roles of n, m swapped)
  FOP_LOSS = 2,    // dq in place over q, per-head Huber sums
  FOP_BWD = 3,     // weight-gradient blocks of one layer + its data gradient, one phase
};

// One step of the per-tile program.  All tensors live in a feature-major shared-memory arena
// arena[feature_row][RP]; a tensor is a list of feature rows kept in the `tab` array, so that
// concatenations (BS_brain.py:154-175) and in-place reuse of dead rows cost nothing.
struct FusedOp {
  int type;
  int in_tab, K;          // GEMM/BWD: input rows (K of them); AGG: input rows (F)
  int out_tab, O;         // GEMM: output rows (O); AGG: output rows (F)
  int w_off, b_off;       // parameter offsets (floats) of W[K_ld][O] and bias[O]
  int relu;               // GEMM: relu epilogue
  int add_tab;            // AGG: rows added to the result (-1: none)
  int gate_tab;           // AGG: relu gate rows for the result (-1); BWD: gate rows of the dx outputs (-1)
  int transposed;         // AGG
  int dz_tab;             // BWD: dz rows (O)
  int dxk_tab, dx_tab, n_dx;   // BWD: weight-row index of every requested dx column, its output rows, count (multiple of 4)
  int blk0, nblk;         // BWD: range of global weight-gradient block ids of this layer
  int bias0;              // BWD: first global bias-gradient slot of this layer
  int wt0, nwt;           // BWD: nwt = RS > 0 marks a small layer (blocks split over RS lanes by rows), wt0 = its offset in
                          // the shared accumulator dWs
  int slot_w, slot_b;     // per-slot weights (G == N): floats between the W / bias of consecutive node slots
  int w_span;             // per-slot weights: floats of the layer's contiguous [W[G][K][O] | bias[G][O]] span (staged by TMA)
  int pad_;               // (sizeof(FusedOp) stays a multiple of 16 bytes: shared-memory carve-up alignment)
};
static_assert(sizeof(FusedOp) % 16 == 0, "FusedOp must stay 16-byte sized");

constexpr int kFusedMaxOps = 48;
constexpr int kFusedMaxTab = 3072;
constexpr int kFusedThreads = 384;
constexpr int kFusedBlkPerThread = 1;
constexpr int kFusedWarps = kFusedThreads / 32;

struct FusedProgram {
  int n_ops;
  int n_tab;
  int n_rows;             // arena feature rows
  int zero_row;           // an all-zero feature row
  int x0_row0, Dn, De;    // input rows: node features then edge features, contiguous from x0_row0
  int q_tab, y_tab;       // Q rows / target rows (CH each)
  int N, TG, R, RP, CH, F;
  int n_params;
  int n_blocks;           // weight-gradient 4x4 blocks over all layers
  int n_bias;             // bias-gradient slots over all layers
  int train;              // 1: loss + backward, 0: forward only
  int n_small;            // floats of the shared weight-gradient accumulator (small layers)
  // Row layout of a tile.  Shared weights (G == 1): graph-major, node (g, n) is arena row g * N + n.  Per-slot weights
  // (G == N, the reference's model, BS_brain.py:121-200): slot-major, row n * TGp + g with TGp = TG rounded up to 4, so
  // that the 4 rows of a register tile always share one weight set.
  int G, TGp, row_g, row_n;
  int wstage_floats;      // per-slot weights: floats of ONE of the two shared-memory weight buffers the layers' spans are
                          // streamed through (bulk-async copies, one layer ahead); 0: read the weights from global memory
  int stage_row0;         // shared weights: first of the arena rows (dead at tile start) that receive the tile's contiguous
                          // node | edge | target spans by bulk-async copy before they are transposed into feature rows; -1: none
  FusedOp ops[kFusedMaxOps];
  int tab[kFusedMaxTab];
  // block b -> op index, k0, o0 (packed: op<<16 | k0<<8 | o0), bias slot -> op<<16 | o
  int blk_info[kFusedThreads * kFusedBlkPerThread];
  int bias_info[kFusedThreads];
};

// Build the program for a brain configuration.  Returns 0 on success, non-zero (with last_error) if
// the configuration is outside the fused path (then the layered kernels are used).
struct FusedShape {
  int N, Dn, De, F, CH, S, H1, H2, H3;
  int G;                   // 1: shared weights, N: one weight set per node slot
  const int* layer_K;      // stacked weight rows per layer
  const int* layer_O;
  const size_t* w_off;
  const size_t* b_off;
  int n_layers;
  size_t n_params;
};
int fused_build_program(const FusedShape& s, int TG, int train, FusedProgram* out);
size_t fused_smem_bytes(const FusedProgram& p);
int fused_pick_tg(const FusedShape& s, int B, int train);

// Launch.  prog_dev: the program in device memory.  partial_dev: [grid][n_params] per-CTA gradient partials.
int fused_launch(const FusedProgram& prog_host, const FusedProgram* prog_dev, const float* params, const float* node,
                 const float* edge, const uint32_t* in_mask, const uint32_t* out_mask, const float* y, float* q_out,
                 float* partial_dev, float* head_loss, int B, int grid, cudaStream_t st);
int fused_grid(const FusedProgram& p, int B);
int fused_set_trace(long long* dev_buf);      // dev_buf: >= kFusedMaxOps + 1 entries, or nullptr to disable
// backward contractions of the shared-weight kernel: 0 = FP32 pipe, 1 = tensor cores (mma.sync TF32, 3 passes per product)
constexpr int kFusedMmaDefault = 1;
int fused_get_mma();
int fused_set_mma(int mode);

// Per-CTA partial rows are kFusedPartialTail floats longer than the parameter vector: the tail carries the CTA's
// per-head Huber sums, so that loss and gradient leave through the same reduction (no atomics, no memset).
constexpr int kFusedPartialTail = 32;
inline long fused_partial_stride(long n_params) { return n_params + kFusedPartialTail; }

// grad[i] = sum_c partial[c][i] for i < n; optional Keras-Adam in the same pass (t >= 1); columns n .. n + n_tail - 1
// are summed into tail_out (the per-head losses).  partial rows are `stride` floats apart.
int fused_reduce_adam(const float* partial, int n_cta, long stride, float* grad, float* p, float* m, float* v, long n,
                      int n_tail, float* tail_out, int t, float lr, float b1, float b2, float eps, float gscale,
                      cudaStream_t st);

}  // namespace v2v
