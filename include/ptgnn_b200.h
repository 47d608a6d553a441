/*
 * ptgnn_b200 -- C ABI of the B200-native message-passing hot path of microsoft/ptgnn.
 *
 * The reference is pure Python; the native boundary its hot path crosses is the third-party
 * torch_scatter operator (`torch_scatter.scatter`, called at
 * ptgnn/neuralmodels/gnn/messagepassing/abstractmessagepassing.py:44-50) plus ATen ops.  This library
 * replaces that boundary and the layer bodies above it.  Every entry point:
 *   - is extern "C", takes plain pointers/sizes (no torch types),
 *   - takes DEVICE pointers unless the parameter is marked [host],
 *   - enqueues its work on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     and returns without synchronising (except the *_host entry points, which copy from/to host
 *     buffers and synchronise before returning),
 *   - returns PTGNN_OK (0) or a negative PTGNN_E_* code; ptgnn_b200_last_error() gives the message.
 *
 * Index dtype: the reference delivers int64 index tensors (graphneuralnetwork.py:461-467); the edge
 * plan down-converts them once per minibatch to int32 (E, N < 2^31 is checked).
 * Floating-point dtype: fp32 state/weights (entry points suffixed _f32); bf16 state variants are
 * suffixed _bf16 (fp32 accumulation, like the reference's AMP path abstractmessagepassing.py:43-50).
 */
#ifndef PTGNN_B200_H_
#define PTGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTGNN_B200_ABI_VERSION 2
#define PTGNN_MAX_EDGE_TYPES 128 /* etype is stored as uint8 in the plan; 128 keeps launch params < 4 KB */

enum {
    PTGNN_OK = 0,
    PTGNN_E_INVALID = -1,     /* bad argument (null pointer, unsupported dimension, ...) */
    PTGNN_E_UNSUPPORTED = -2, /* valid reference configuration without a native kernel yet */
    PTGNN_E_CUDA = -3,        /* CUDA runtime error (message in ptgnn_b200_last_error) */
    PTGNN_E_WORKSPACE = -4,   /* workspace too small */
    PTGNN_E_INDEX = -5        /* edge index out of [0, num_nodes) (only reported by *_host / validate calls) */
};

/* torch_scatter `reduce=` strings (torch_scatter.scatter). */
enum { PTGNN_REDUCE_SUM = 0, PTGNN_REDUCE_MEAN = 1, PTGNN_REDUCE_MAX = 2, PTGNN_REDUCE_MIN = 3 };
/* activations used by MlpMessagePassingLayer (mlpmessagepassing.py:20,26). */
enum { PTGNN_ACT_NONE = 0, PTGNN_ACT_GELU = 1, PTGNN_ACT_TANH = 2, PTGNN_ACT_RELU = 3 };

int ptgnn_b200_abi_version(void);
const char *ptgnn_b200_last_error(void);
/* Number of kernel launches issued by this library in this process (bench.py's "gpu_launches"). */
int64_t ptgnn_b200_launch_count(void);

/* Optional per-kernel timing (bench.py's roofline leg): while enabled, every launch is bracketed by CUDA events
 * on its own stream.  ptgnn_b200_kernel_timing_read synchronises, adds each category's elapsed milliseconds and
 * launch count into ms[cat] / launches[cat] (cat < ncat) and clears the record. */
enum {
    PTGNN_KERNEL_PLAN = 0,    /* all edge-plan kernels */
    PTGNN_KERNEL_MESSAGE = 1, /* edge_message_kernel */
    PTGNN_KERNEL_REDUCE = 2,  /* segment_reduce_kernel */
    PTGNN_KERNEL_GRU = 3,     /* gru_update_kernel */
    PTGNN_KERNEL_DENSE = 4,   /* dense_update_kernel */
    PTGNN_KERNEL_PACK = 5,    /* weight packing / conversion */
    PTGNN_KERNEL_CATEGORIES = 6
};
int ptgnn_b200_kernel_timing_enable(int32_t enable);
int ptgnn_b200_kernel_timing_read(double *ms /*[host]*/, int64_t *launches /*[host]*/, int32_t ncat);

/* ------------------------------------------------------------------------------------------------
 * Edge plan -- replaces `torch.cat([adj[1] for adj in adjacency_lists])` (gatedmessagepassing.py:46,
 * mlpmessagepassing.py:102-109) and the index->row grouping inside torch_scatter.scatter: a canonical,
 * STABLE target-sorted CSR over the concatenated per-type edge lists, built once per minibatch and
 * reused by every layer.  Edge id e = position in cat(types) order.
 *   row_ptr[N+1]     CSR offsets over targets
 *   perm[E]          sorted position j -> edge id (stable: per target, edge-id ascending)
 *   pos[E]           edge id -> sorted position
 *   src_sorted[E]    source node of the edge at sorted position j
 *   etype_sorted[E]  edge type of the edge at sorted position j
 *   src32/tgt32[E]   the int64 inputs down-converted, edge-id order
 *   status[1]        number of out-of-range indices seen (they are clamped to 0); 0 = valid.  Any device-accessible
 *                    int32: device memory, or pinned host memory the caller can poll without synchronising
 * ---------------------------------------------------------------------------------------------- */
size_t ptgnn_b200_plan_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int ptgnn_b200_plan_build(int64_t num_nodes, int64_t num_source_nodes /* bound for src ids; <= 0: num_nodes */,
                          int32_t num_types,
                          const int64_t *const *src_ptrs /*[host] T device pointers*/,
                          const int64_t *const *tgt_ptrs /*[host] T device pointers*/,
                          const int64_t *counts /*[host] T edge counts*/, int32_t *row_ptr, int32_t *perm,
                          int32_t *pos, int32_t *src_sorted, uint8_t *etype_sorted, int32_t *src32, int32_t *tgt32,
                          int32_t *status, void *workspace, size_t workspace_bytes, void *stream);

/* The two phases of ptgnn_b200_plan_build as separate calls.  `plan_convert` = down-conversion, range check, in-degree
 * histogram, row_ptr (everything the fused layer kernels and the block plan need); `plan_sort` = the stable sort by target and
 * the sorted arrays (needed by the unfused layer kernels and by ptgnn_b200_segment_reduce_f32 callers).  Same workspace size. */
int ptgnn_b200_plan_convert(int64_t num_nodes, int64_t num_source_nodes, int32_t num_types,
                            const int64_t *const *src_ptrs /*[host]*/, const int64_t *const *tgt_ptrs /*[host]*/,
                            const int64_t *counts /*[host]*/, int32_t *row_ptr, int32_t *src32, int32_t *tgt32, int32_t *status,
                            void *workspace, size_t workspace_bytes, void *stream);
int ptgnn_b200_plan_sort(int64_t num_nodes, int32_t num_types, const int64_t *counts /*[host]*/, int32_t *perm, int32_t *pos,
                         int32_t *src_sorted, uint8_t *etype_sorted, const int32_t *src32, const int32_t *tgt32,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Segmented reduce -- replaces torch_scatter.scatter(src, index, dim=0, dim_size=N, reduce) as called
 * at abstractmessagepassing.py:44-50.  `messages` is [E, D] fp32.  If `perm` is NULL the rows are
 * already in plan order (row j belongs to the target whose CSR range contains j); otherwise row
 * perm[j] is read for sorted position j (rows in edge-id order, as torch_scatter receives them).
 * out [N, D] fp32: empty targets -> 0.  arg_out (optional, max/min only) [N, D] int64 = edge id of the
 * winning message (first occurrence wins ties), E for empty targets (torch_scatter's sentinel).
 * ---------------------------------------------------------------------------------------------- */
int ptgnn_b200_segment_reduce_f32(const float *messages, const int32_t *row_ptr, const int32_t *perm,
                                  int64_t num_nodes, int64_t num_edges, int32_t dim, int32_t reduce, float *out,
                                  int64_t *arg_out, void *stream);

/* One-shot torch_scatter.scatter drop-in: builds a single-type plan from the int64 `index` and reduces.
 * workspace >= ptgnn_b200_scatter_workspace_bytes(N, E).  status (optional, device-accessible int32): receives the number of
 * indices outside [0, num_nodes) (they are routed to row 0; the reference raises / asserts in that case). */
size_t ptgnn_b200_scatter_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int ptgnn_b200_scatter_f32(const float *src, const int64_t *index, int64_t num_edges, int32_t dim, int64_t num_nodes,
                           int32_t reduce, float *out, int64_t *arg_out, int32_t *status, void *workspace,
                           size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * GatedMessagePassingLayer.forward (gatedmessagepassing.py:37-69), eval mode, no edge features:
 *   m_e = W_{t(e)} h_{src(e)} ; a_v = reduce_{e: tgt(e)=v} m_e ; h'_v = GRUCell(a_v, h_v)
 * edge_weights: [host] array of T device pointers, each nn.Linear.weight [D, H] row-major.
 * gru_*: nn.GRUCell parameters weight_ih [3H, D], weight_hh [3H, H], bias_ih/bias_hh [3H] (gate order r,z,n).
 * type_off: [host] T+1 prefix offsets of the per-type edge counts (edge-id space).
 * gather_states: rows that the plan's source ids index.  NULL = node_states (the normal, single-GPU case).  For a
 * node-range shard (multi-GPU split of one connected graph) node_states holds the num_nodes OWNED rows (targets,
 * local ids) and gather_states the all-gathered [num_source_nodes, H] states (sources, global ids).
 * workspace >= ptgnn_b200_gated_workspace_bytes(...): message buffer [E, D] + aggregate [N, D] + packed / TF32-split
 * weights.  Dimensions that fit the tensor-core tiles (H % 32 == 0, D % 16 == 0) run on tcgen05 (3xTF32, fp32-exact);
 * other multiples of 4 run on the FFMA kernels.  PTGNN_B200_DISABLE_TC=1 forces the FFMA kernels.
 * ---------------------------------------------------------------------------------------------- */
size_t ptgnn_b200_gated_workspace_bytes(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t state_dim,
                                        int32_t message_dim);
int ptgnn_b200_gated_forward_f32(const float *node_states, const float *gather_states /* NULL: node_states */,
                                 int64_t num_nodes, int32_t state_dim, int32_t message_dim,
                                 int32_t num_types, const int64_t *type_off /*[host]*/, const int32_t *row_ptr,
                                 const int32_t *pos, const int32_t *src32,
                                 const float *const *edge_weights /*[host] T device pointers*/, const float *gru_w_ih,
                                 const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh, int32_t reduce,
                                 float *out_states, void *workspace, size_t workspace_bytes, void *stream);

/* Weight cache (optional).  Every forward call first derives working copies of the parameters (TF32 hi/lo splits,
 * gate-blocked GRU packing; bf16 conversions in the bf16 variant).  A caller whose parameters do not change between calls
 * (inference, or between optimiser steps) can own that buffer: pass `weight_cache` (device memory of at least
 * `*_weight_cache_bytes`, 256-byte aligned) and `cache_valid` = 0 on the first call with a given set of parameter VALUES
 * (the copies are derived into the cache), 1 afterwards (they are reused; the parameter pointers are then not read by the
 * derivation).  `*_weight_cache_bytes` == 0 means these dimensions have nothing to cache: pass NULL.  Results are
 * bit-identical to the uncached entry points. */
size_t ptgnn_b200_gated_weight_cache_bytes(int32_t num_types, int32_t state_dim, int32_t message_dim);
int ptgnn_b200_gated_forward_cached_f32(const float *node_states, const float *gather_states, int64_t num_nodes,
                                        int32_t state_dim, int32_t message_dim, int32_t num_types,
                                        const int64_t *type_off /*[host]*/, const int32_t *row_ptr, const int32_t *pos,
                                        const int32_t *src32, const float *const *edge_weights /*[host]*/,
                                        const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih,
                                        const float *gru_b_hh, int32_t reduce, float *out_states, void *workspace,
                                        size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                        int32_t cache_valid, void *stream);
size_t ptgnn_b200_gated_weight_cache_bytes_bf16(int32_t num_types, int32_t state_dim, int32_t message_dim);
int ptgnn_b200_gated_forward_cached_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes,
                                         int32_t state_dim, int32_t message_dim, int32_t num_types,
                                         const int64_t *type_off /*[host]*/, const int32_t *row_ptr, const int32_t *pos,
                                         const int32_t *src32, const float *const *edge_weights /*[host]*/,
                                         const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih,
                                         const float *gru_b_hh, int32_t reduce, uint16_t *out_states, void *workspace,
                                         size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                         int32_t cache_valid, void *stream);

/* bf16 variant (BASELINE.json configs[3]): node_states / gather_states / out_states are bf16 [*, H] (raw uint16 bits),
 * module parameters stay fp32 and are converted per call; messages and aggregates are bf16 in HBM, every accumulation
 * (tensor-core accumulators, segmented reduce, gate math) is fp32 -- the arithmetic of the reference under
 * torch.autocast(bfloat16) (fp32 scatter: abstractmessagepassing.py:43-50).  Needs H % 32 == 0, D % 16 == 0, 64 <= D <= 256. */
size_t ptgnn_b200_gated_workspace_bytes_bf16(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t state_dim,
                                             int32_t message_dim);
int ptgnn_b200_gated_forward_bf16(const uint16_t *node_states, const uint16_t *gather_states /* NULL: node_states */,
                                  int64_t num_nodes, int32_t state_dim, int32_t message_dim, int32_t num_types,
                                  const int64_t *type_off /*[host]*/, const int32_t *row_ptr, const int32_t *pos,
                                  const int32_t *src32, const float *const *edge_weights /*[host] T device pointers, fp32*/,
                                  const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                                  int32_t reduce, uint16_t *out_states, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * MlpMessagePassingLayer.forward (mlpmessagepassing.py:68-117), eval mode, default message MLP
 * (mlp_hidden_layers = 0: one bias-free Linear, mlp.py:65-74), string aggregator, no edge features:
 *   m_e = W_{t(e)} [h_src ; h_tgt] ; a_v = reduce m_e ; h'_v = act2(W_d LN(act1(a_v)) + b_d)
 * edge_weights[t]: [D, 2H] (or [D, H] when use_target_state == 0).  ln_weight/ln_bias NULL => no LayerNorm;
 * dense_weight NULL => no dense layer (output dim = D).  dense_weight [Hout, D], dense_bias [Hout].
 * ---------------------------------------------------------------------------------------------- */
size_t ptgnn_b200_mlp_workspace_bytes(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t in_dim,
                                      int32_t message_dim, int32_t out_dim, int32_t use_target_state);
int ptgnn_b200_mlp_forward_f32(const float *node_states, const float *gather_states /* NULL: node_states */,
                               int64_t num_nodes, int32_t in_dim, int32_t message_dim,
                               int32_t out_dim, int32_t num_types, const int64_t *type_off /*[host]*/,
                               const int32_t *row_ptr, const int32_t *pos, const int32_t *src32, const int32_t *tgt32,
                               const float *const *edge_weights /*[host] T device pointers*/,
                               int32_t use_target_state, int32_t reduce, int32_t message_activation,
                               const float *ln_weight, const float *ln_bias, float ln_eps, const float *dense_weight,
                               const float *dense_bias, int32_t dense_activation, float *out_states, void *workspace,
                               size_t workspace_bytes, void *stream);

/* bf16 variant: node_states / gather_states / out_states are bf16 (raw uint16 bits), every parameter stays fp32 and is
 * converted per call; messages bf16, aggregation + activation + LayerNorm in fp32, dense update bf16 with fp32 accumulation
 * (the reference under torch.autocast(bfloat16)).  Needs in_dim % 32 == 0 (>= 64), message_dim % 16 == 0 in [64, 256],
 * out_dim % 16 == 0 (>= 64) when a dense layer is present. */
size_t ptgnn_b200_mlp_workspace_bytes_bf16(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t in_dim,
                                           int32_t message_dim, int32_t out_dim, int32_t use_target_state);
int ptgnn_b200_mlp_forward_bf16(const uint16_t *node_states, const uint16_t *gather_states /* NULL: node_states */,
                                int64_t num_nodes, int32_t in_dim, int32_t message_dim, int32_t out_dim, int32_t num_types,
                                const int64_t *type_off /*[host]*/, const int32_t *row_ptr, const int32_t *pos,
                                const int32_t *src32, const int32_t *tgt32,
                                const float *const *edge_weights /*[host] T device pointers, fp32*/,
                                int32_t use_target_state, int32_t reduce, int32_t message_activation,
                                const float *ln_weight, const float *ln_bias, float ln_eps, const float *dense_weight,
                                const float *dense_bias, int32_t dense_activation, uint16_t *out_states, void *workspace,
                                size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused aggregation (round 2): gather -> per-type Linear -> segmented reduce in ONE kernel; the [E, D]
 * message tensor of gatedmessagepassing.py:64 / mlpmessagepassing.py:100-112 is never materialised.
 *
 * Block plan: the edges sorted, stably, by (target block, edge type, target), blocks of `block_targets`
 * (<= 240, multiple of 8; ptgnn_b200_block_plan_block_targets recommends one) consecutive target nodes:
 *   group_off[ceil(N / B) * T + 1]  sorted-edge offsets of the (block, type) groups
 *   src_f[E]                        source node of the edge at sorted position j
 *   tl_f[E]                         target of that edge, relative to its block's first node
 * Built from the src32 / tgt32 arrays of ptgnn_b200_plan_build, once per minibatch, reused by all layers.
 * `status` (optional): two device-accessible int32 words (device memory or pinned host memory); the layer kernels
 * set status[0] = 1 if a node state, status[1] = 1 if an edge weight is outside the fp16 range of the fp32-exact
 * 3xFP16 split (|x| >= 65504, inf, NaN): the result is then not valid, use the unfused entry points.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t block_targets;
    const int32_t *group_off;
    const int32_t *src_f;
    const uint8_t *tl_f;
    int32_t *status;
} ptgnn_b200_block_plan;

int32_t ptgnn_b200_block_plan_block_targets(int64_t num_nodes);
size_t ptgnn_b200_block_plan_workspace_bytes(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t block_targets);
int ptgnn_b200_block_plan_build(int64_t num_nodes, int32_t num_types, const int64_t *type_off /*[host]*/,
                                const int32_t *src32, const int32_t *tgt32, int32_t block_targets, int32_t *group_off,
                                int32_t *src_f, uint8_t *tl_f, void *workspace, size_t workspace_bytes, void *stream);

/* 1 if these dimensions run on the fused kernel (message_dim == 128; state_dim in {64, 128} for fp32 states,
 * {64, 128, 256} for bf16 states); otherwise use the unfused entry points above. */
int32_t ptgnn_b200_fused_supported(int32_t bf16_states, int32_t state_dim, int32_t message_dim);

/* GatedMessagePassingLayer.forward through the fused kernel.  Same contract as ptgnn_b200_gated_forward_cached_{f32,bf16}
 * (node_states / gather_states / out_states are fp32 when bf16_states == 0, bf16 otherwise; `row_ptr` = CSR offsets of the
 * edge plan, used by reduce = mean); the edge arrays come from the block plan.  fp32 states are computed fp32-exactly with
 * three fp16 tensor-core products per term ("3xFP16", see csrc/fused_mp.cuh). */
size_t ptgnn_b200_gated_fused_workspace_bytes(int32_t bf16_states, int64_t num_nodes, int64_t num_source_nodes,
                                              int32_t num_types, int32_t state_dim, int32_t message_dim);
size_t ptgnn_b200_gated_fused_weight_cache_bytes(int32_t bf16_states, int32_t num_types, int32_t state_dim, int32_t message_dim);
int ptgnn_b200_gated_forward_fused(int32_t bf16_states, const void *node_states, const void *gather_states /* NULL: node_states */,
                                   int64_t num_nodes, int64_t num_source_nodes, int32_t state_dim, int32_t message_dim,
                                   int32_t num_types, const ptgnn_b200_block_plan *block_plan, const int32_t *row_ptr,
                                   const float *const *edge_weights /*[host] T device pointers, fp32*/, const float *gru_w_ih,
                                   const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh, int32_t reduce,
                                   void *out_states, void *workspace, size_t workspace_bytes, void *weight_cache,
                                   size_t weight_cache_bytes, int32_t cache_valid, void *stream);

/* The fp32 fused path computes on packed states: every fp32 state x as two fp16 numbers (hi | lo') in rows of 2 * state_dim
 * halfs (ptgnn_b200_packed_state_bytes per tensor).  In a stack of layers (GraphNeuralNetwork.gnn,
 * ptgnn/neuralmodels/gnn/graphneuralnetwork.py:121-131) the packing pass of layer i + 1 is redundant: the GRU kernel of layer i
 * can write its new states in both forms.  ptgnn_b200_gated_forward_fused_chained = ptgnn_b200_gated_forward_fused for fp32
 * states, plus
 *   packed_states_in  (optional): the packed form of node_states, as written through packed_states_out by the previous layer;
 *   packed_states_out (optional): [num_nodes] packed rows, receives the packed form of out_states (bit-identical to what the
 *                                 next call would derive from out_states itself). */
size_t ptgnn_b200_packed_state_bytes(int64_t num_nodes, int32_t state_dim);
int ptgnn_b200_gated_forward_fused_chained(const float *node_states, const float *gather_states /* NULL: node_states */,
                                           const void *packed_states_in, int64_t num_nodes, int64_t num_source_nodes,
                                           int32_t state_dim, int32_t message_dim, int32_t num_types,
                                           const ptgnn_b200_block_plan *block_plan, const int32_t *row_ptr,
                                           const float *const *edge_weights /*[host] T device pointers, fp32*/,
                                           const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih,
                                           const float *gru_b_hh, int32_t reduce, float *out_states, void *packed_states_out,
                                           void *workspace, size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                           int32_t cache_valid, void *stream);

/* MlpMessagePassingLayer.forward through the fused kernel (contract of ptgnn_b200_mlp_forward_{f32,bf16}); the message
 * activation and the LayerNorm run in the fused kernel's write-out. */
size_t ptgnn_b200_mlp_fused_workspace_bytes(int32_t bf16_states, int64_t num_nodes, int64_t num_source_nodes, int32_t num_types,
                                            int32_t in_dim, int32_t message_dim, int32_t out_dim, int32_t use_target_state);
int ptgnn_b200_mlp_forward_fused(int32_t bf16_states, const void *node_states, const void *gather_states /* NULL: node_states */,
                                 int64_t num_nodes, int64_t num_source_nodes, int32_t in_dim, int32_t message_dim,
                                 int32_t out_dim, int32_t num_types, const ptgnn_b200_block_plan *block_plan,
                                 const int32_t *row_ptr, const float *const *edge_weights /*[host] T device pointers, fp32*/,
                                 int32_t use_target_state, int32_t reduce, int32_t message_activation, const float *ln_weight,
                                 const float *ln_bias, float ln_eps, const float *dense_weight, const float *dense_bias,
                                 int32_t dense_activation, void *out_states, void *workspace, size_t workspace_bytes,
                                 void *stream);

/* ptgnn_b200_mlp_forward_fused with caller-owned derived weights (fp32 states): weight_cache [>= ptgnn_b200_mlp_fused_weight_cache_bytes]
 * holds the packed edge weights and the split dense weight; pass cache_valid = 0 after the parameters changed (they are then
 * re-derived into the cache), 1 to reuse them.  weight_cache == NULL or bf16 states: identical to ptgnn_b200_mlp_forward_fused. */
size_t ptgnn_b200_mlp_fused_weight_cache_bytes(int32_t bf16_states, int32_t num_types, int32_t in_dim, int32_t message_dim, int32_t out_dim,
                                               int32_t use_target_state);
int ptgnn_b200_mlp_forward_fused_cached(int32_t bf16_states, const void *node_states, const void *gather_states /* NULL: node_states */,
                                        int64_t num_nodes, int64_t num_source_nodes, int32_t in_dim, int32_t message_dim,
                                        int32_t out_dim, int32_t num_types, const ptgnn_b200_block_plan *block_plan,
                                        const int32_t *row_ptr, const float *const *edge_weights /*[host] T device pointers, fp32*/,
                                        int32_t use_target_state, int32_t reduce, int32_t message_activation, const float *ln_weight,
                                        const float *ln_bias, float ln_eps, const float *dense_weight, const float *dense_bias,
                                        int32_t dense_activation, void *out_states, void *workspace, size_t workspace_bytes,
                                        void *weight_cache, size_t weight_cache_bytes, int32_t cache_valid, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Backward support (host side: ptgnn_b200/autograd.py).  The pointwise half of the GRUCell backward
 * (torch.nn.GRUCell, gatedmessagepassing.py:69): gi / gh [N, 3H] = gate pre-activations (order r, z, n), h [N, H] the previous
 * states, grad_out [N, H] the gradient of the new states -> d_gi, d_gh [N, 3H] (gradients of the pre-activations) and
 * d_h_direct [N, H] = grad_out * z.  The GEMM-shaped products around it use ptgnn_b200_linear_f32.
 * ---------------------------------------------------------------------------------------------- */
/* Operand preparation for the parameter-gradient products dW = A^T B (K = edges or nodes; run by the host as three fp16 tensor-core
 * GEMMs per product): out row r = split(x[index ? index[r] : r] * (scale ? *scale : 1)), split(v) = (hi = rn16(v),
 * lo = rn16((v - hi) * 2^11)) -- the 3xFP16 representation of csrc/fused_mp.cuh.  x [*, cols] fp32, index int32 (NULL: identity),
 * scale: device scalar (NULL: 1), hi / lo [rows_out, cols] fp16.  cols % 8 == 0. */
int ptgnn_b200_gather_split_f16(const float *x, const int32_t *index, int64_t rows_out, int32_t cols, const float *scale, void *hi,
                                void *lo, void *stream);
int ptgnn_b200_gru_gate_grads_f32(const float *gi, const float *gh, const float *h, const float *grad_out, int64_t num_nodes,
                                  int32_t state_dim, float *d_gi, float *d_gh, float *d_h_direct, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side minibatch finalisation -- the graph-structure half of GraphNeuralNetworkModel.extend_minibatch_with /
 * finalize_minibatch (ptgnn/neuralmodels/gnn/graphneuralnetwork.py:386-493).  The host concatenates the graphs' LOCAL int32
 * ids (edge sources / targets of one edge type, or reference nodes); item_ptr [G+1] (device, int64) = where each graph's items
 * start in that concatenation, node_ptr [G+1] (device, int64) = exclusive prefix sum of the graphs' node counts.
 *   offset_ids:  out[i] = local_ids[i] + node_ptr[g(i)]   replaces `sample_adj_list + nodes_in_mb_so_far` (:419-424, :436) and the
 *                                                         np.concatenate + torch.tensor(..., int64) of :463-469, :485-491
 *   segment_ids: out[i] = g(i)                            replaces __create_node_to_graph_idx (:441-443, one Python iteration per
 *                                                         node) with item_ptr = node_ptr, and the extend() of :431-434
 * ---------------------------------------------------------------------------------------------- */
int ptgnn_b200_offset_ids(const int32_t *local_ids, int64_t num_items, const int64_t *item_ptr, const int64_t *node_ptr,
                          int32_t num_graphs, int64_t *out, void *stream);
int ptgnn_b200_segment_ids(const int64_t *item_ptr, int32_t num_segments, int64_t num_items, int64_t *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone pieces of the layers, for the configurations the fused entry points do not cover.
 *   linear:        out[rows, out_dim] = act(x W^T + b)  -- one `nn.Linear` (+ activation) of `MLP.forward` (mlp.py:79-80) or of an
 *                  Mlp layer's dense update (mlpmessagepassing.py:116); fp32-exact on the tensor cores when the dims fit.
 *   edge_messages: messages[out_row[e]] = W_{t(e)} [source_states[src32[e]] ; target_states[tgt32[e]]]  for every edge e in
 *                  cat(types) order (gatedmessagepassing.py:54-60, mlpmessagepassing.py:88-98) -- the [E, D] tensor that a module
 *                  aggregator (`AbstractMessageAggregation`, e.g. PNA: pna_aggregation.py:27-56) or a second MLP layer consumes.
 *                  out_row = the plan's `pos` (target-sorted rows) or the identity (edge order, as the reference lays them out).
 * ---------------------------------------------------------------------------------------------- */
size_t ptgnn_b200_linear_workspace_bytes(int32_t in_dim, int32_t out_dim);
int ptgnn_b200_linear_f32(const float *x, int64_t rows, int32_t in_dim, const float *weight /*[out_dim, in_dim]*/,
                          const float *bias /* NULL: none */, int32_t out_dim, int32_t activation, float *out, void *workspace,
                          size_t workspace_bytes, void *stream);
/*   grucell:       out = nn.GRUCell(input [rows, input_dim], hidden [rows, state_dim])  (gatedmessagepassing.py:69) */
size_t ptgnn_b200_grucell_workspace_bytes(int32_t state_dim, int32_t input_dim);
int ptgnn_b200_grucell_f32(const float *input, const float *hidden, int64_t rows, int32_t state_dim, int32_t input_dim,
                           const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh, float *out, void *workspace,
                           size_t workspace_bytes, void *stream);
size_t ptgnn_b200_edge_messages_workspace_bytes(int32_t num_types, int32_t in_dim, int32_t message_dim, int32_t use_target_state);
int ptgnn_b200_edge_messages_f32(const float *source_states, const float *target_states /* rows tgt32 indexes; NULL if unused */,
                                 int32_t in_dim, int32_t message_dim, int32_t num_types, const int64_t *type_off /*[host]*/,
                                 const int32_t *src32, const int32_t *tgt32, const int32_t *out_row,
                                 const float *const *edge_weights /*[host] T device pointers*/, int32_t use_target_state,
                                 float *messages, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer convenience entry point (used for the end-to-end measurement): all pointers are HOST
 * memory; copies inputs to the device, builds the plan, runs `num_layers` GatedMessagePassingLayers
 * (layer l uses weight set l; pass the same pointers to share weights), copies the final states back
 * and synchronises.  Mirrors GraphNeuralNetwork.gnn's loop (graphneuralnetwork.py:121-131) for a
 * homogeneous stack of gated layers.  Returns PTGNN_E_INDEX if an edge index is out of range.
 * ---------------------------------------------------------------------------------------------- */
int ptgnn_b200_gated_gnn_forward_host_f32(const float *node_states, int64_t num_nodes, int32_t state_dim,
                                          int32_t num_types, const int64_t *const *src_ptrs,
                                          const int64_t *const *tgt_ptrs, const int64_t *counts, int32_t num_layers,
                                          const float *const *edge_weights /* num_layers*T host pointers [D,H] */,
                                          const float *const *gru_w_ih, const float *const *gru_w_hh,
                                          const float *const *gru_b_ih, const float *const *gru_b_hh,
                                          int32_t reduce, float *out_states);

#ifdef __cplusplus
}
#endif
#endif /* PTGNN_B200_H_ */
