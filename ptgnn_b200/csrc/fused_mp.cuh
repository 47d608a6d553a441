// Fused gather -> per-edge-type Linear -> segmented reduce: the aggregation half of a message-passing layer in ONE
// persistent tcgen05 kernel that never materialises the [E, D] message tensor.
//
//   reference  ptgnn/neuralmodels/gnn/messagepassing/gatedmessagepassing.py:50-68   (F.embedding + Linear + cat + scatter)
//              ptgnn/neuralmodels/gnn/messagepassing/mlpmessagepassing.py:82-112
//              ptgnn/neuralmodels/gnn/messagepassing/abstractmessagepassing.py:38-50 (torch_scatter.scatter)
//
// Orientation.  The per-type weight W_t [D = 128, K] is the MMA's A operand (M = D) and lives in TENSOR MEMORY; the
// gathered node-state rows are the B operand (N = edges of one (target block, edge type) group, any multiple of 16) in
// shared memory; the accumulator is therefore msg^T: TMEM lane = message feature d, TMEM column = edge.  Two things
// follow: (1) a group of n edges costs an N = ceil16(n) MMA, not a padded 128-row tile -- (block, type) groups hold a
// few dozen edges; (2) the reduction over edges of the same target runs ALONG the columns of one lane, i.e. it is a
// plain sequential loop in the epilogue thread that owns feature d: no shuffles, no atomics, accumulation in the
// reference's edge order (per target: type-major, then list order), bit-reproducible.
//
// Work decomposition.  Targets are cut into blocks of B <= 256 consecutive nodes; the block plan (plan.cu) sorts the
// edges by (block, type, target).  A CTA owns a block at a time and keeps its aggregate agg_s[B][D] fp32 in shared
// memory across all edge types, then writes it once (optionally through mean / GELU / LayerNorm: the Mlp layer's
// pre-dense epilogue).  HBM traffic per layer = gathered rows (L2-resident per graph) + agg once.
//
// Arithmetic.  NPROD = 1: bf16 states and weights, one kind::f16 MMA per K-step, fp32 accumulation (the reference under
// torch.autocast(bfloat16); each message is rounded to bf16 before the fp32 reduction like the autocast Linear's
// output).  NPROD = 3: fp32-exact "3xFP16": every fp32 value x is carried as two fp16 numbers, hi = rn(x) and
// lo' = rn((x - hi) * 2^11) -- 22 significant bits, absolute error <= 2^-36 for tiny values -- and
// x*w ~= hi*hi + 2^-11 (hi*lo' + lo'*hi): three kind::f16 MMAs per K-step (half the tensor time of 3xTF32), the two
// small products in a separate correction accumulator (tensor-core accumulation truncates).  |x| >= 65504 cannot be
// represented: the packing kernels raise a status flag and the host raises (PTGNN_B200_FP32_MODE=tf32 selects the
// unfused 3xTF32 kernels).
//
// Roles (16 warps):  0-3 and 8-11 EPILOGUE, two warpgroups: thread d <-> TMEM lane d; group 0 takes the columns whose target
//   lies in the lower half of the block, group 1 the upper half (disjoint agg_s rows, no synchronisation between them) |
//   4 MMA issuer | 5-6 ROW GATHERERS (16-byte cp.async into a 3-slot ring, SWIZZLE_128B K-major) |
//   7 SCHEDULER (block -> group offsets table ring) |
//   12-15 WEIGHT LOADERS (global -> registers -> tcgen05.st, TMEM A buffers, double buffered; setmaxnreg gives them the
//   registers warps 4-7 release -- ptxas only extends a role's budget when that role is the LAST branch of the kernel).
// TMEM (512 columns): [0,256) two weight buffers | [256,512) two accumulator sets (main | correction).
#pragma once
#include "common.cuh"

namespace ptgnn {
namespace fused {

constexpr int kD = 128;                 // message dimension handled by this kernel (= MMA M)
constexpr int kMaxBlockTargets = 240;   // agg_s = B * 512 bytes of shared memory

struct Epilogue {                       // applied to the aggregated row at write-out (Mlp layers), else act = NONE / ln = null
    int act;
    const float *ln_w, *ln_b;
    float ln_eps;
};

bool supported(int nprod, int K, int D, int use_target);
// bytes of the packed edge weights (TMEM-friendly layout), of one packed state row, and of the packed-state scratch
size_t packed_weight_bytes(int nprod, int num_types, int K, int use_target);
size_t packed_state_bytes(int nprod, int64_t rows, int K);
int recommended_block_targets(int64_t num_nodes);

// weights[t]: fp32 [128, nseg*K] row-major (nn.Linear.weight) -> packed
int pack_weights(int nprod, int num_types, int K, int use_target, const float *const *weights, void *packed, int32_t *status,
                 cudaStream_t st);
// fp32 states [rows, K] -> fp16 (hi | lo') rows of 4K bytes   (NPROD = 3 only; bf16 states are gathered as they are)
int pack_states(const float *h, int64_t rows, int K, void *packed, int32_t *status, cudaStream_t st);

struct AggregateArgs {
    int nprod;                      // 3: fp32-exact (states given as packed hi|lo' rows), 1: bf16
    const void *src_rows;           // rows indexed by src_f: packed fp16 pairs (nprod 3) or bf16 (nprod 1), K elements per row
    const void *tgt_rows;           // rows indexed by target id (use_target only)
    int64_t num_nodes;              // target rows
    int K, num_types, use_target, reduce, block_targets;
    const int32_t *group_off, *src_f;
    const uint8_t *tl_f;
    const int32_t *row_ptr;         // CSR offsets over targets (mean only; may be null otherwise)
    const void *packed_weights;
    Epilogue epi;
    void *out;                      // [num_nodes, 128]: out_mode 0 = fp32, 1 = bf16, 2 = packed fp16 (hi | lo') rows of 256 halfs
    int out_mode;                   //   (2 = what the weights-stationary GRU kernel takes as its A operand)
    int32_t *status;                // optional: status[0] = 1 if an aggregate is outside the fp16 range (out_mode 2)
};
int aggregate(const AggregateArgs &a, cudaStream_t st);

}  // namespace fused
}  // namespace ptgnn
