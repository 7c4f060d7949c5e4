/*
 * v2v_gnn.h -- C-ABI of the B200-native V2V graph-convolution engine.
 *
 * Drop-in boundary for the hot path of Coolzyh/Globecom2020-ResourceAllocationGNN:
 * the custom Keras layers and the "brain" of BS_brain.py (reference file:line
 * cited per entry point).  Plain `extern "C"`, raw pointers and sizes, no torch
 * or C++ types.  The Python mirror of the reference classes
 * (globecom2020-resourceallocationgnn_b200/{layers,brain}.py) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is
 *     available from v2v_last_error() (thread-local).  The Python layer turns a
 *     non-zero status into ValueError/RuntimeError like Keras would raise.
 *   - `*_dev` pointers are device pointers, `*_host` pointers are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     All device entry points are stream-ordered, never synchronise and never
 *     allocate, so a step is CUDA-graph capturable.
 *   - tensors are row-major and dense: H[B][N][F] etc.
 *   - dtype: V2V_F32 = 0 (fp32 storage and math), V2V_BF16 = 1 (bf16 storage,
 *     fp32 accumulate).
 *   - adjacency is carried in compact form: one bit per directed edge,
 *     W = ceil(N/32) 32-bit words per node, in both orientations
 *        in_mask [b][m][w] bit n = Adj[b][n][m]   (whom m aggregates from)
 *        out_mask[b][n][w] bit m = Adj[b][n][m]   (whom n contributes to)
 *     Adj is the matrix of BS_brain.py:441-445 (NOT the Kronecker expansion of
 *     :492-493; the host adapter recovers Adj from it by strided sampling).
 */
#ifndef V2V_GNN_H
#define V2V_GNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2V_F32 0
#define V2V_BF16 1
#define V2V_F64 2   /* host views only: the reference feeds numpy fp64 (cast to fp32 on feed, as Keras does) */

#define V2V_MAX_SEG 4

const char* v2v_last_error(void);
int v2v_version(void);
/* number of SMs of the current device (grid sizing), <0 on error */
int v2v_device_sm_count(void);
/* kernels this library has launched in this process so far (bench.py's gpu_launches) */
long v2v_launch_count(void);

/* ------------------------------------------------------------------------
 * Adjacency packing.  Replaces the host-side np.kron of BS_brain.py:492-493,
 * :603, :621 (dense (B,NF,NF) operand) by 2*N*W words per graph.
 * adj_dev: fp32 [B][N][N].  nonbinary_flag_dev (int, may be NULL) is set to 1
 * if any entry is neither 0 nor 1 (then the weighted path must be used).
 * ---------------------------------------------------------------------- */
int v2v_adj_pack_masks(const float* adj_dev, int B, int N,
                       uint32_t* in_mask_dev, uint32_t* out_mask_dev,
                       int* nonbinary_flag_dev, void* stream);

/* ------------------------------------------------------------------------
 * AggLayer.call  (BS_brain.py:69-76):
 *     out[b][m][:] = sum_n Adj[b][n][m] * H[b][n][:]      (+ addend[b][m][:])
 * mask_dev = in_mask for the forward; pass out_mask to get the transposed
 * aggregation, which is the backward w.r.t. H.  addend_dev may be NULL, or may
 * alias out_dev (accumulate).  H/out/addend are `dtype` [B][N][F].
 * ---------------------------------------------------------------------- */
int v2v_agg_mask(const void* H_dev, const uint32_t* mask_dev, const void* addend_dev,
                 void* out_dev, int B, int N, int F, int dtype, void* stream);

/* Same operator with launch flags.  V2V_AGG_INDEPENDENT: the caller asserts that none of the
 * operands is produced or still read by the kernel launched immediately before on `stream`
 * (e.g. a stream of aggregations over distinct buffers); the kernel then skips the
 * programmatic-dependent-launch wait and may overlap the tail of its predecessor. */
#define V2V_AGG_INDEPENDENT 1u
int v2v_agg_mask_ex(const void* H_dev, const uint32_t* mask_dev, const void* addend_dev,
                    void* out_dev, int B, int N, int F, int dtype, unsigned flags, void* stream);

/* Weighted (non 0/1) adjacency, fp32 only: adj_dev fp32 [B][N][N];
 * transpose = 0: out[b][m] = sum_n adj[b][n][m] H[b][n]   (forward)
 * transpose = 1: out[b][n] = sum_m adj[b][n][m] H[b][m]   (backward) */
int v2v_agg_dense(const float* H_dev, const float* adj_dev, const float* addend_dev,
                  float* out_dev, int B, int N, int F, int transpose, void* stream);

/* ------------------------------------------------------------------------
 * GNNLayer.call (BS_brain.py:44-51) and Dense (:176-200) as one row-GEMM:
 *     out[r][:] = act( [seg0[r] | seg1[r] | ...] . W[g(r)] + bias[g(r)] )
 * rows r = b*N + n; group g(r) = n if G == N (one weight set per node slot,
 * as the reference instantiates them, :121-142) or 0 if G == 1 (shared).
 * W: fp32 [G][ldw][n_out], the first K = sum(seg_width) rows are used.
 * act: 0 linear, 1 relu.   All fp32.
 * ---------------------------------------------------------------------- */
int v2v_dense_fwd(int n_seg, const float* const* seg_dev, const int* seg_width,
                  const float* W_dev, int ldw, const float* bias_dev,
                  float* out_dev, int B, int N, int G, int n_out, int act, void* stream);

/* Backward w.r.t. selected input columns:
 *     dX[r][k] = sum_o dZ[r][o] * W[g][k][o],   dZ = gate_in ? dY*(gate_in>0) : dY
 * for k in [k0_a, k0_a+w_a) -> dxa_dev [rows][w_a] and (optional, w_b > 0)
 * k in [k0_b, k0_b+w_b) -> dxb_dev.  gate_out_dev (optional, [rows][w_a], only
 * when w_b == 0): dxa is zeroed where gate_out <= 0 (relu of the producer). */
int v2v_dense_bwd_data(const float* dY_dev, const float* gate_in_dev,
                       const float* W_dev, int ldw,
                       int k0_a, int w_a, float* dxa_dev,
                       int k0_b, int w_b, float* dxb_dev,
                       const float* gate_out_dev,
                       int B, int N, int G, int n_out, void* stream);

/* Weight/bias gradient, ACCUMULATED into dW/db (zero them first):
 *     dW[g][k][o] += sum_r x[r][k] dZ[r][o],  db[g][o] += sum_r dZ[r][o] */
int v2v_dense_bwd_weight(int n_seg, const float* const* seg_dev, const int* seg_width,
                         const float* dY_dev, const float* gate_in_dev,
                         float* dW_dev, int ldw, float* db_dev,
                         int B, int N, int G, int n_out, void* stream);

/* ------------------------------------------------------------------------
 * huber_loss + compile(loss=huber_loss) (BS_brain.py:86-87, :214):
 * per head k: mean over (B, CH) of huber_1(q - y); dq = clip(q-y,-1,1)/(B*CH)
 * scaled by grad_scale.  head_loss_dev fp32 [N] is ACCUMULATED (zero it first).
 * ---------------------------------------------------------------------- */
int v2v_huber_loss_grad(const float* q_dev, const float* y_dev, float* dq_dev,
                        float* head_loss_dev, int B, int N, int CH,
                        float grad_scale, void* stream);

/* DQN target rule (BS_brain.py:668-692):
 * y = p, then y[b][k][a[b][k]] = r[b] + gamma * max_a p_next[b][k][a]. */
int v2v_td_target(const float* p_dev, const float* p_next_dev, const int32_t* action_dev,
                  const float* reward_dev, float gamma, float* y_dev,
                  int B, int N, int CH, void* stream);

/* Epsilon-greedy action selection of Agent.select_action_while_training (BS_brain.py:308-352) for E environments:
 * sched_dev fp32 [4] = {environment step, base, decrement per step, floor}; epsilon = max(floor, base - decrement * step)
 * (the linear anneal of :315-324).  Environment e explores when u_explore_dev[e] < epsilon: every link takes its entry of
 * random_action_dev [E][N] (:330-333); otherwise link (e, n) takes the FIRST maximiser of q_dev[e][n][:] (:336-344). */
int v2v_dqn_select_actions(const float* q_dev, const float* u_explore_dev, const int32_t* random_action_dev,
                           const float* sched_dev, int32_t* action_dev, int E, int N, int CH, void* stream);

/* Memory.add (BS_brain.py:252-256) for a device-resident ring of `capacity` slots: transition t of T goes to slot
 * (*head_dev + t) % capacity of every ring tensor (node / node_next [.][N][Dn], edge / edge_next [.][N][De], the two
 * adjacency mask orientations [.][N][ceil(N/32)], action [.][N], reward [.]); then *head_dev advances by T (mod capacity)
 * and, if step_dev is not NULL, *step_dev += 1 -- all on the device, so the launch can be replayed from a CUDA graph.
 * done_dev: one zero-initialised unsigned int of scratch. */
int v2v_dqn_replay_write(float* ring_node, float* ring_edge, float* ring_node_next, float* ring_edge_next,
                         int32_t* ring_in_mask, int32_t* ring_out_mask, int32_t* ring_action, float* ring_reward,
                         const float* node, const float* edge, const float* node_next, const float* edge_next,
                         const int32_t* in_mask, const int32_t* out_mask, const int32_t* action, const float* reward,
                         long long* head_dev, float* step_dev, unsigned* done_dev, int T, long capacity, int N, int Dn,
                         int De, void* stream);

/* keras.optimizers.Adam(lr, beta_1, beta_2) update rule of Keras 2.2.4
 * (BS_brain.py:212), epsilon outside the sqrt; t = 1-based iteration,
 * grad_scale multiplies g first (1/world_size after a sum all-reduce). */
int v2v_adam_step(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, long n,
                  int t, float lr, float beta1, float beta2, float eps, float grad_scale,
                  void* stream);

/* ------------------------------------------------------------------------
 * The brain: BS (BS_brain.py:90-239).  Owns online + target parameters, Adam
 * state, gradients and the activation workspace for up to max_batch graphs.
 * ---------------------------------------------------------------------- */
typedef struct v2v_brain v2v_brain;

typedef struct v2v_brain_config {
  int num_d2d;        /* N: nodes (V2V pairs) per graph            (:95)  */
  int node_dim;       /* Dn = ((node_info-1)*CH+1)*neighbor         (:101) */
  int edge_dim;       /* De = edge_info*CH                          (:102) */
  int feedback;       /* F: GNN feature width                       (:98)  */
  int num_ch;         /* CH: Q-values per node                      (:97)  */
  int stages;         /* GNN stages (reference: 3, :147-166)               */
  int per_slot;       /* 1: one weight set per node slot (reference, :121-200); 0: shared */
  int hidden[3];      /* decision MLP widths (reference 80,40,20, :176-178) */
  int max_batch;      /* workspace capacity in graphs                      */
  int dtype;          /* V2V_F32, or V2V_BF16: bf16 operands / fp32 accumulate on the tensor cores (see v2v_tt_plan) */
  float lr, beta1, beta2, eps;   /* Adam (:212): 1e-3, 0.5, 0.999, 1e-7    */
} v2v_brain_config;

int v2v_brain_create(const v2v_brain_config* cfg, v2v_brain** out);
void v2v_brain_destroy(v2v_brain* b);
/* floats in one parameter set (online == target == grads == Adam m == v) */
long v2v_brain_param_count(const v2v_brain* b);
/* get_weights / set_weights (BS_brain.py:239): flat fp32, per layer W[G][K][n_out] then bias[G][n_out];
 * which = 0 online, 1 target, 2 grads, 3 adam m, 4 adam v */
int v2v_brain_get_params(v2v_brain* b, int which, float* host_out, void* stream);
int v2v_brain_set_params(v2v_brain* b, int which, const float* host_in, void* stream);
float* v2v_brain_param_ptr(v2v_brain* b, int which);     /* device pointer */
/* The shared-weight brain (per_slot = 0, N <= 32, binary adjacency, zero neighbour input) runs
 * forward/backward as ONE fused kernel per call; enable = 0 forces the layer-by-layer kernels
 * (what per-slot weights, weighted adjacency or N > 32 always use).  Default: enabled. */
int v2v_brain_set_fused(v2v_brain* b, int enable);
/* Describes the fused program for a batch of B graphs: info8 = {capable, graphs per tile, arena
 * feature rows, shared-memory bytes, phases, weight-gradient blocks, bias slots, table entries}. */
int v2v_brain_fused_info(v2v_brain* b, int B, int train, int* info8);
/* predict on the tensor cores (csrc/tc_forward.cu: tcgen05.mma kind::tf32, 3 passes per contraction = fp32-grade
 * products, accumulators in TMEM): mode 0 never, 1 automatic (default: batches that give every SM >= 2 tiles of 128 node
 * rows), 2 whenever the brain is capable (shared weights, N <= 32, binary adjacency, zero neighbour input).
 * info4 = {capable, mode, graphs per tile, shared-memory bytes}.  Environment override at creation: V2V_TENSOR_CORE. */
int v2v_brain_set_tensor_core(v2v_brain* b, int mode);
int v2v_brain_tensor_core_info(const v2v_brain* b, int* info4);
/* Host-only query (no device): info8 = {capable, graphs per tile, layers, shared-memory bytes, operand planes per tile
 * slot, floats of the staged weight image, tensor-memory columns used per slot, tcgen05.mma instructions per tile}. */
int v2v_tc_plan(const v2v_brain_config* cfg, int* info8);
/* The bf16 configuration (dtype = V2V_BF16; BASELINE configs[2]): forward, Huber head and the whole backward run as ONE
 * tcgen05 kernel (csrc/tc_train.cu) -- bf16 contraction operands, fp32 accumulation in tensor memory, fp32 bias / ReLU /
 * aggregation / loss, fp32 master weights and Adam.  Shared weights, N <= 32, <= 3 stages, feedback width 16.
 * Host-only query: info8 = {capable, graphs per tile, steps per training tile, tcgen05.mma per training tile,
 * shared-memory bytes, operand planes, bf16 weight-image elements, weight-gradient column blocks}. */
int v2v_tt_plan(const v2v_brain_config* cfg, int* info8);
/* Profiling aid of that kernel: CTA 0 writes clock64() stamps of its first two tiles into dev_buf[(tile * 24 + step) * 8 + i]
 * (i = 0 MMA thread saw the operands, 1 MMAs issued and committed, 2 epilogue saw the accumulator, 3 tensor-memory loads
 * back, 4 planes written, 5 fenced and arrived; step 23 / i = 6: tile start).  dev_buf >= 2 * 24 * 8 entries; NULL disables. */
int v2v_tt_set_trace(long long* dev_buf);
/* debugging aid: tensor-core forward that also dumps the raw fp32 accumulator [128][Npad] of `layer` for the first tile */
int v2v_brain_tc_debug(v2v_brain* b, const float* node_dev, const float* edge_dev, const uint32_t* in_mask_dev, int B,
                       int layer, float* q_dev, float* dbg_dev, int* npad_out, void* stream);
/* Profiling aid: lane 0 of every warp of CTA 0 writes clock64() into dev_buf[(phase*12 + warp)*2 + {0: work
 * done, 1: barrier released}] for its first tile (dev_buf >= 49*12*2 entries, device memory); NULL disables. */
int v2v_fused_set_trace(long long* dev_buf);
/* Which pipe runs the backward contractions (weight and data gradients of a.W1 + b.W2 + c.W3 and of the decision MLP:
 * BS_brain.py:44-51, :176-200 under :218-223) of the fused shared-weight fp32 kernel: 0 = FP32 pipe (FFMA), 1 = tensor
 * cores (mma.sync m16n8k8 TF32, three passes per product = fp32-grade products, fp32 accumulation).  The forward always
 * runs in plain fp32.  Default 1 (configs[1]: 65.8 -> 56.5 us per step; gradients within 3e-6 of the FFMA form).
 * Process-wide; the environment variable V2V_FUSED_MMA sets the initial value.  Per-slot weights always use the FP32 pipe. */
int v2v_fused_set_mma(int mode);
int v2v_fused_get_mma(void);
/* Same query from a configuration alone (host-only, no device needed). */
int v2v_fused_plan(const v2v_brain_config* cfg, int B, int train, int* info8);
/* update_target_model (BS_brain.py:237-239) */
int v2v_brain_update_target(v2v_brain* b, void* stream);
int v2v_brain_get_iterations(const v2v_brain* b);
int v2v_brain_set_iterations(v2v_brain* b, int t);

/* predict (BS_brain.py:225-235): node [B][N][Dn], edge [B][N][De] fp32 device,
 * in_mask device (binary adjacency) or adj_dev fp32 [B][N][N] (weighted; pass
 * in_mask = NULL).  q_dev fp32 [B][N][CH].  target = 1 uses the target net.
 * neighbor_dev: the D{k}_Neighbor_Input of the first GNN stage, fp32 [B][N][F];
 * NULL means all zeros, which is what the reference always feeds (:478, :589) and
 * lets the engine skip that contraction. */
int v2v_brain_forward(v2v_brain* b, const float* node_dev, const float* edge_dev,
                      const float* neighbor_dev,
                      const uint32_t* in_mask_dev, const float* adj_dev,
                      int B, int target, float* q_dev, void* stream);

/* forward (online net, activations kept) + Huber + backward into the gradient
 * buffer.  head_loss_dev fp32 [N] receives the per-head mean losses.  Does not
 * touch the parameters: the caller may all-reduce v2v_brain_param_ptr(b, 2)
 * across ranks before v2v_brain_apply_adam. */
int v2v_brain_forward_backward(v2v_brain* b, const float* node_dev, const float* edge_dev,
                               const float* neighbor_dev,
                               const uint32_t* in_mask_dev, const uint32_t* out_mask_dev,
                               const float* adj_dev, const float* y_dev, int B,
                               float* head_loss_dev, void* stream);
/* iterations += 1; Keras-Adam on the online parameters with g * grad_scale */
int v2v_brain_apply_adam(v2v_brain* b, float grad_scale, void* stream);

/* ------------------------------------------------------------------------
 * Host-buffer entry points over STRIDED VIEWS of the caller's memory: what BS.predict / BS.train_dnn receive
 * from the reference's Agent (BS_brain.py:495-504, :664-665, :724-728) -- one array per node slot, numpy fp64,
 * the adjacency as kron(Adj, I_F) -- without any repacking in Python.  Element (r, c) of a view is read at
 * ptr + (r * row_stride + c * col_stride) * sizeof(dtype) and written, converted to fp32, to element
 * dst_off + r * dst_row_stride + c of the packed staging tensor it belongs to:
 *   node [B][N][Dn], edge [B][N][De], neighbor [B][N][F] (n_neigh = 0: all zeros, the reference's case),
 *   adj [B][N][N] (values outside {0,1} select the weighted-adjacency kernels), y [B][N][CH].
 * A per-slot array D{k}_Node_Input (B, Dn) is the view {rows B, cols Dn, dst_off k*Dn, dst_row_stride N*Dn};
 * the Kronecker adjacency (B, N*F, N*F) is ONE view {rows B*N, cols N, row_stride F*N*F, col_stride F}.
 * A persistent worker pool (V2V_HOST_THREADS, default min(8, cores/2)) gathers the views into pinned staging and
 * each tensor's H2D copy is enqueued as soon as it is complete; the call synchronises `stream` once, at the end.
 * q_host fp32 [B][N][CH]; head_loss_host fp32 [N]. */
typedef struct v2v_host_view {
  const void* ptr;
  int dtype;                 /* V2V_F32 or V2V_F64 */
  long rows, cols;
  long row_stride, col_stride;   /* in elements of dtype */
  long dst_off, dst_row_stride;  /* in floats of the staging tensor */
} v2v_host_view;
int v2v_brain_predict_views(v2v_brain* b, const v2v_host_view* node, int n_node, const v2v_host_view* edge, int n_edge,
                            const v2v_host_view* neigh, int n_neigh, const v2v_host_view* adj, int n_adj, int B,
                            int target, float* q_host, void* stream);
int v2v_brain_train_views(v2v_brain* b, const v2v_host_view* node, int n_node, const v2v_host_view* edge, int n_edge,
                          const v2v_host_view* neigh, int n_neigh, const v2v_host_view* adj, int n_adj,
                          const v2v_host_view* y, int n_y, int B, float* head_loss_host, void* stream);
int v2v_host_stage_threads(void);
/* The host-only pieces of that path (no device involved): gather + convert views into dst (dst_elems floats) on the
 * worker pool; check = 1 reports bit 0 of *flags_out if any element is outside {0,1}, check = 2 reports bit 1 if any
 * element is non-zero.  v2v_host_pack_adjacency is adj_pack_kernel on the host (N <= 32), reading ONE [B*N][N] view of
 * the caller's adjacency (dense or Kronecker): in_mask[b][m] bit n = out_mask[b][n] bit m = (adj[b][n][m] != 0);
 * bit 0 of *flags_out reports values outside {0,1}. */
int v2v_host_gather(const v2v_host_view* views, int n_views, float* dst, long dst_elems, int check, int* flags_out);
int v2v_host_pack_adjacency(const v2v_host_view* adj_view, int B, int N, uint32_t* in_mask, uint32_t* out_mask,
                            int* flags_out);

/* ------------------------------------------------------------------------
 * Batched environment: E independent copies of the reference simulator stepped on the device, state in
 * HBM, fp32.  Randomness is INJECTED (arrays of draws) so that every kernel is checked against vectors recorded from the
 * unmodified Environment.py.  Layouts: pos [E][N][2] (x, y), dir [E][N] (0 up, 1 down, 2 left, 3 right), vel [E][N],
 * v2v_shadow [E][N][N], v2i_shadow [E][N], dest [E][N], v2v_ff [E][N][N][RB] (= V2V_channels_with_fastfading),
 * v2i_ff [E][N][RB], v2i_abs [E][N].
 *  - renew_channels (Environment.py:378-404 with :63-120, :140-165): z_v2v [E][N][N] ~ N(0, 3), z_v2i [E][N] ~ N(0, 8) are
 *    the shadowing draws, ff_v2v [E][N][N][RB][2] / ff_v2i [E][N][RB][2] the standard-normal (re, im) fast-fading draws.
 *  - pack_state (BS_brain.py:389-407, :441-469): node [E][N][2RB+1], edge [E][N][RB], the adjacency as bit masks
 *    (N <= 32; both or neither) and/or dense adj [E][N][N] (may be NULL).
 *  - reward (Environment.py:406-458; BS_brain.py:515-519): per-link V2V rates [E][N], V2I rates [E][min(RB,N)], the V2I
 *    interference [E][RB] and reward[e] = v2v_weight * sum(V2V) + v2i_weight * sum(V2I); outputs may be NULL.  N <= 32.
 *  - renew_positions (Environment.py:236-345): u [E][N] = the uniform draw a vehicle uses if it reaches a crossing.
 *  - choose_destinations (Environment.py:360-376): receiver = candidate floor(u (N-3)) of the other vehicles by distance,
 *    the two farthest excluded. */
int v2v_env_renew_channels(const float* pos, const float* vel, float* v2v_shadow, float* v2i_shadow, const float* z_v2v,
                           const float* z_v2i, const float* ff_v2v, const float* ff_v2i, float* v2v_ff, float* v2i_ff,
                           float* v2i_abs, int E, int N, int RB, void* stream);
int v2v_env_pack_state(const int* dest, const float* v2v_ff, const float* v2i_ff, float* node, float* edge,
                       uint32_t* in_mask, uint32_t* out_mask, float* adj, int E, int N, int RB, void* stream);
int v2v_env_reward(const int* actions, const int* dest, const float* v2v_ff, const float* v2i_ff, const float* v2i_abs,
                   float* v2v_rate, float* v2i_rate, float* interference, float* reward, float v2v_weight, float v2i_weight,
                   int E, int N, int RB, void* stream);
int v2v_env_renew_positions(float* pos, int* dir, const float* vel, const float* u, int E, int N, void* stream);
int v2v_env_choose_destinations(const float* pos, const float* u, int* dest, int E, int N, void* stream);

/* ------------------------------------------------------------------------
 * Data parallelism (one process per GPU).  The reference has no distributed path; every batch row is an
 * independent graph, so ranks own contiguous batch shards and exchange ONE flat gradient per step.
 * v2v_comm is that exchange over NVLink peer memory (cudaIpc): create one per rank with the payload size
 * (parameters + num_d2d per-head losses), all-gather the handles by any means (torch.distributed here), open.
 * v2v_comm_allreduce_adam = reduce the local per-CTA partials + push to every peer + wait + sum in rank order +
 * Keras-Adam (BS_brain.py:212), in ONE kernel.  All ranks must call it the same number of times.
 * ---------------------------------------------------------------------- */
typedef struct v2v_comm v2v_comm;
int v2v_comm_create(long n_floats, int world, int rank, v2v_comm** out);
void v2v_comm_destroy(v2v_comm* c);
int v2v_comm_ipc_handle_bytes(void);
int v2v_comm_get_ipc_handle(v2v_comm* c, void* handle_out);
int v2v_comm_open_peers(v2v_comm* c, const void* handles /* world blobs in rank order */);
int v2v_comm_allreduce_adam(v2v_comm* c, const float* partial_dev, int n_cta, long n_src,
                            const float* extra_dev, int n_extra, float* grad_dev, float* p_dev,
                            float* m_dev, float* v_dev, float* extra_out_dev, int t, float lr, float beta1,
                            float beta2, float eps, void* stream);
/* general row layout: partial rows are row_stride floats apart with n_src payload columns, of which the first n_adam are
 * parameters (gradient + Keras-Adam); the other n_src - n_adam columns and the n_extra floats of extra_dev are only
 * averaged over ranks into extra_out_dev (the fused brain kernel appends its per-head Huber sums to every partial row) */
int v2v_comm_allreduce_adam_ex(v2v_comm* c, const float* partial_dev, int n_cta, long row_stride, long n_adam, long n_src,
                               const float* extra_dev, int n_extra, float* grad_dev, float* p_dev, float* m_dev,
                               float* v_dev, float* extra_out_dev, int t, float lr, float beta1, float beta2, float eps,
                               void* stream);
int v2v_comm_check(v2v_comm* c, void* stream);
/* A rank whose wait for a peer times out sets an error flag and SKIPS the parameter update (it never trains on stale
 * data).  v2v_comm_poll_error enqueues a stream-ordered copy of that flag into a pinned mirror and fails if an earlier
 * copy already reported a timeout (non-blocking; v2v_brain_train_step_dp calls it every 32 steps);
 * v2v_comm_poll_result reads the mirror after the caller has synchronised the stream (the host entry points do). */
int v2v_comm_poll_error(v2v_comm* c, void* stream);
int v2v_comm_poll_result(v2v_comm* c);
/* optional phase trace of the exchange kernel: trace_dev = device buffer of v2v_comm_num_chunks() * 6 uint64 (null
 * disables); per chunk: kernel entry, producer complete, pushed + fenced, all ranks arrived, Adam done (globaltimer ns) */
int v2v_comm_set_trace(v2v_comm* c, unsigned long long* trace_dev);
int v2v_comm_num_chunks(v2v_comm* c);
/* data-parallel train_dnn: local fwd + Huber + bwd, then v2v_comm_allreduce_adam; head_loss_dev receives the
 * per-head losses averaged over ranks */
/* the strided-view host entry point (see v2v_brain_train_views) for one rank of a data-parallel job */
int v2v_brain_train_views_dp(v2v_brain* b, v2v_comm* comm, const v2v_host_view* node, int n_node,
                             const v2v_host_view* edge, int n_edge, const v2v_host_view* neigh, int n_neigh,
                             const v2v_host_view* adj, int n_adj, const v2v_host_view* y, int n_y, int B,
                             float* head_loss_host, void* stream);
int v2v_brain_train_step_dp(v2v_brain* b, v2v_comm* comm, const float* node_dev, const float* edge_dev,
                            const float* neighbor_dev, const uint32_t* in_mask_dev, const uint32_t* out_mask_dev,
                            const float* adj_dev, const float* y_dev, int B, float* head_loss_dev, void* stream);

/* train_dnn (BS_brain.py:218-223) == one fwd+bwd+Adam step on exactly B rows. */
int v2v_brain_train_step(v2v_brain* b, const float* node_dev, const float* edge_dev,
                         const float* neighbor_dev,
                         const uint32_t* in_mask_dev, const uint32_t* out_mask_dev,
                         const float* adj_dev, const float* y_dev, int B,
                         float* head_loss_dev, void* stream);

/* Host-buffer variants: what the reference-facing plugin calls.  Inputs are
 * host fp32 arrays (pinned for async copies), adjacency as fp32 [B][N][N];
 * copies H2D, packs masks on device, runs, copies results D2H and waits. */
int v2v_brain_predict_host(v2v_brain* b, const float* node_host, const float* edge_host,
                           const float* neighbor_host /* may be NULL */,
                           const float* adj_host, int B, int target, float* q_host, void* stream);
int v2v_brain_train_host(v2v_brain* b, const float* node_host, const float* edge_host,
                         const float* neighbor_host /* may be NULL */,
                         const float* adj_host, const float* y_host, int B,
                         float* head_loss_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2V_GNN_H */
