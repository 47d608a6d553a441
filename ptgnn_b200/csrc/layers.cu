// GatedMessagePassingLayer / MlpMessagePassingLayer forward on B200 (fp32-exact path).
//
//   reference ptgnn/neuralmodels/gnn/messagepassing/gatedmessagepassing.py:37-69
//   reference ptgnn/neuralmodels/gnn/messagepassing/mlpmessagepassing.py:68-117  (+ ptgnn/neuralmodels/mlp.py:79-80)
//
// Per layer, three kernels instead of the reference's ~3*T+4 ATen/torch_scatter launches:
//   1. edge_message_kernel   gather h[src] (and h[tgt]) rows straight into the GEMM A-tile, multiply by the
//                            edge type's weight, write each message row ONCE, already at its target-sorted
//                            position (pos[e]) -- F.embedding + cat + Linear + cat(all_messages) fused.
//   2. segment_reduce_kernel (reduce.cuh) streaming CSR reduce = torch_scatter.scatter (+ GELU/LayerNorm for Mlp).
//   3. gru_update_kernel     nn.GRUCell: both GEMMs ([agg;h] x packed gate weights) + gate math in one pass, or
//      dense_update_kernel   Linear(+bias) + Tanh of the Mlp layer.
#include <stdlib.h>

#include "fused_mp.cuh"
#include "gru_ws.cuh"
#include "gemm_simt.cuh"
#include "layers_tc.cuh"
#include "reduce.cuh"

namespace ptgnn {

// =================================================================================================
// 1. per-edge messages
// =================================================================================================
struct MsgParams {
    const float *weights[PTGNN_MAX_EDGE_TYPES];  // per type: [D, K] row-major (nn.Linear.weight)
    int32_t edge_off[PTGNN_MAX_EDGE_TYPES + 1];  // edge-id prefix offsets
    int32_t tile_off[PTGNN_MAX_EDGE_TYPES + 1];  // CTA-tile prefix offsets (128 edges per tile)
    int num_types;
};

template <int TN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
edge_message_kernel(const __grid_constant__ MsgParams p, const float *__restrict__ h, const float *__restrict__ h_tgt,
                    int H, int use_target, int D,
                    const int32_t *__restrict__ src32, const int32_t *__restrict__ tgt32,
                    const int32_t *__restrict__ pos, float *__restrict__ msg) {
    using Tile = GemmTile<TN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *pipe = reinterpret_cast<float *>(smem_raw);
    int *s_idx0 = reinterpret_cast<int *>(smem_raw + (Tile::SMEM_BYTES - Tile::IDX_BYTES));
    int *s_idx1 = s_idx0 + GEMM_BM;
    int *s_out = s_idx1 + GEMM_BM;

    const int tile = blockIdx.x;
    int t = 0;
    {   // largest t with tile_off[t] <= tile
        int lo = 0, hi = p.num_types - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (p.tile_off[mid] <= tile) lo = mid; else hi = mid - 1;
        }
        t = lo;
    }
    const int e0 = p.edge_off[t] + (tile - p.tile_off[t]) * GEMM_BM;
    const int e_end = p.edge_off[t + 1];
    if (threadIdx.x < GEMM_BM) {
        const int e = e0 + threadIdx.x;
        const bool ok = e < e_end;
        s_idx0[threadIdx.x] = ok ? src32[e] : -1;
        s_idx1[threadIdx.x] = (ok && use_target) ? tgt32[e] : -1;
        s_out[threadIdx.x] = ok ? pos[e] : -1;
    }
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    AOperand A;
    A.a0 = h; A.a1 = h_tgt; A.ld0 = H; A.ld1 = H; A.K0 = H; A.K = use_target ? 2 * H : H;
    const int n0 = blockIdx.y * Tile::BN;
    gemm_mainloop<TN, 0>(acc, pipe, A, s_idx0, s_idx1, p.weights[t], A.K, n0, D);

    // stage the C tile through shared memory so that every message row leaves as one coalesced burst
    float *Cs = pipe;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) Cs[(ty + 16 * i) * Tile::CS_STRIDE + tx + 16 * j] = acc[i][j];
    __syncthreads();
    constexpr int F4_PER_ROW = Tile::BN / 4;
#pragma unroll
    for (int i = 0; i < (GEMM_BM * F4_PER_ROW) / GEMM_THREADS; ++i) {
        const int idx = threadIdx.x + i * GEMM_THREADS;
        const int row = idx / F4_PER_ROW, c4 = idx % F4_PER_ROW;
        const int orow = s_out[row];
        const int col = n0 + c4 * 4;
        if (orow >= 0 && col < D) {
            const float4 v = *reinterpret_cast<const float4 *>(Cs + row * Tile::CS_STRIDE + c4 * 4);
            *reinterpret_cast<float4 *>(msg + (size_t)orow * D + col) = v;
        }
    }
}

// =================================================================================================
// 3a. GRUCell update
// =================================================================================================
// Packed gate weights, one block of 32 hidden units per `jb`:
//   P1[jb][n][k] = weight_ih[(n/32)*H + jb*32 + n%32][k]   n in [0,96): gates r, z, n (input part),  k < D
//   P2[jb][n][k] = weight_hh[(n/32)*H + jb*32 + n%32][k]   n in [0,96): gates r, z, n (hidden part), k < H
__global__ void pack_gru_weights_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh, int H, int D,
                                        float *__restrict__ P1, float *__restrict__ P2) {
    const int nblk = H / 32;
    const int64_t n1 = (int64_t)nblk * 96 * D, n2 = (int64_t)nblk * 96 * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n1) {
            const int k = (int)(i % D);
            const int n = (int)((i / D) % 96), jb = (int)(i / ((int64_t)96 * D));
            P1[i] = w_ih[(size_t)((n / 32) * H + jb * 32 + n % 32) * D + k];
        } else {
            const int64_t r = i - n1;
            const int k = (int)(r % H);
            const int n = (int)((r / H) % 96), jb = (int)(r / ((int64_t)96 * H));
            P2[r] = w_hh[(size_t)((n / 32) * H + jb * 32 + n % 32) * H + k];
        }
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 2)
gru_update_kernel(const float *__restrict__ agg, const float *__restrict__ h, int num_nodes, int H, int D,
                  const float *__restrict__ P1, const float *__restrict__ P2, const float *__restrict__ b_ih,
                  const float *__restrict__ b_hh, float *__restrict__ out) {
    using Tile = GemmTile<6>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *pipe = reinterpret_cast<float *>(smem_raw);
    int *s_idx0 = reinterpret_cast<int *>(smem_raw + (Tile::SMEM_BYTES - Tile::IDX_BYTES));

    const int row0 = blockIdx.x * GEMM_BM;
    const int jb = blockIdx.y;
    if (threadIdx.x < GEMM_BM) {
        const int r = row0 + threadIdx.x;
        s_idx0[threadIdx.x] = r < num_nodes ? r : -1;
    }
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    // phase 1: [r z n_i] += agg x W_ih^T          (acc columns 0..5)
    AOperand A1;
    A1.a0 = agg; A1.a1 = nullptr; A1.ld0 = D; A1.ld1 = 0; A1.K0 = D; A1.K = D;
    gemm_mainloop<6, 0>(acc, pipe, A1, s_idx0, s_idx0, P1 + (size_t)jb * 96 * D, D, 0, 96);
    // phase 2: [r z] += h x W_hh^T ; n_h = h x W_hn^T   (acc columns 0..3 and 6..7)
    AOperand A2;
    A2.a0 = h; A2.a1 = nullptr; A2.ld0 = H; A2.ld1 = 0; A2.K0 = H; A2.K = H;
    gemm_mainloop<6, 2>(acc, pipe, A2, s_idx0, s_idx0, P2 + (size_t)jb * 96 * H, H, 0, 96);

    // gate math (nn.GRUCell): r = s(i_r + h_r), z = s(i_z + h_z), n = tanh(i_n + r * h_n), h' = (1 - z) n + z h
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = jb * 32 + tx + 16 * half;
        const float br = b_ih[j] + b_hh[j];
        const float bz = b_ih[H + j] + b_hh[H + j];
        const float bin = b_ih[2 * H + j], bhn = b_hh[2 * H + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + ty + 16 * i;
            if (row < num_nodes) {
                const float r = sigmoid_f(acc[i][0 + half] + br);
                const float z = sigmoid_f(acc[i][2 + half] + bz);
                const float n = tanhf(acc[i][4 + half] + bin + r * (acc[i][6 + half] + bhn));
                const float hv = h[(size_t)row * H + j];
                out[(size_t)row * H + j] = (1.0f - z) * n + z * hv;
            }
        }
    }
}

// =================================================================================================
// 3b. dense update of the Mlp layer:  out = act(y W^T + b)
// =================================================================================================
template <int TN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
dense_update_kernel(const float *__restrict__ y, int num_nodes, int D, const float *__restrict__ W /*[Hout, D]*/,
                    const float *__restrict__ bias, int Hout, int act, float *__restrict__ out) {
    using Tile = GemmTile<TN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *pipe = reinterpret_cast<float *>(smem_raw);
    int *s_idx0 = reinterpret_cast<int *>(smem_raw + (Tile::SMEM_BYTES - Tile::IDX_BYTES));
    const int row0 = blockIdx.x * GEMM_BM;
    const int n0 = blockIdx.y * Tile::BN;
    if (threadIdx.x < GEMM_BM) {
        const int r = row0 + threadIdx.x;
        s_idx0[threadIdx.x] = r < num_nodes ? r : -1;
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    AOperand A;
    A.a0 = y; A.a1 = nullptr; A.ld0 = D; A.ld1 = 0; A.K0 = D; A.K = D;
    gemm_mainloop<TN, 0>(acc, pipe, A, s_idx0, s_idx0, W, D, n0, Hout);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int col = n0 + tx + 16 * j;
        if (col >= Hout) continue;
        const float b = bias ? bias[col] : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row0 + ty + 16 * i;
            if (row < num_nodes) out[(size_t)row * Hout + col] = apply_act(acc[i][j] + b, act);
        }
    }
}

// =================================================================================================
// host-side launchers
// =================================================================================================
template <typename K>
static int set_smem(K kernel, int bytes) {
    PTGNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return PTGNN_OK;
}

static int launch_edge_messages(const float *h_src, const float *h_tgt, int H, int D, int use_target, int num_types, const int64_t *type_off,
                                const float *const *weights, const int32_t *src32, const int32_t *tgt32,
                                const int32_t *pos, float *msg, cudaStream_t st) {
    MsgParams p{};
    p.num_types = num_types;
    int tiles = 0;
    for (int t = 0; t < num_types; ++t) {
        p.weights[t] = weights[t];
        p.edge_off[t] = (int32_t)type_off[t];
        p.tile_off[t] = tiles;
        tiles += (int)ceil_div(type_off[t + 1] - type_off[t], GEMM_BM);
    }
    for (int t = num_types; t <= PTGNN_MAX_EDGE_TYPES; ++t) {
        p.edge_off[t] = (int32_t)type_off[num_types];
        p.tile_off[t] = tiles;
    }
    if (tiles == 0) return PTGNN_OK;
    if (D <= 64) {
        using Tile = GemmTile<4>;
        int rc = set_smem(edge_message_kernel<4>, Tile::SMEM_BYTES);
        if (rc) return rc;
        dim3 grid(tiles, (unsigned)ceil_div(D, Tile::BN));
        {
            TimedScope timed__(PTGNN_KERNEL_MESSAGE, st);
            edge_message_kernel<4><<<grid, GEMM_THREADS, Tile::SMEM_BYTES, st>>>(p, h_src, h_tgt, H, use_target, D, src32, tgt32, pos, msg);
        }
    } else {
        using Tile = GemmTile<8>;
        int rc = set_smem(edge_message_kernel<8>, Tile::SMEM_BYTES);
        if (rc) return rc;
        dim3 grid(tiles, (unsigned)ceil_div(D, Tile::BN));
        {
            TimedScope timed__(PTGNN_KERNEL_MESSAGE, st);
            edge_message_kernel<8><<<grid, GEMM_THREADS, Tile::SMEM_BYTES, st>>>(p, h_src, h_tgt, H, use_target, D, src32, tgt32, pos, msg);
        }
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

static int check_layer_dims(const char *who, int64_t N, int64_t E, int H, int D) {
    PTGNN_CHECK_ARG(N >= 0 && N < INT32_MAX && E >= 0 && E < INT32_MAX, "%s: sizes out of range", who);
    PTGNN_CHECK_ARG(H > 0 && H % 4 == 0 && H <= 1024, "%s: state dim %d must be a multiple of 4 (<= 1024)", who, H);
    PTGNN_CHECK_ARG(D > 0 && D % 4 == 0 && D <= 512, "%s: message dim %d must be a multiple of 4 (<= 512)", who, D);
    return PTGNN_OK;
}

// Tensor cores are the default; PTGNN_B200_DISABLE_TC=1 forces the FFMA kernels (A/B measurements, debugging).
static bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PTGNN_B200_DISABLE_TC");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// `fused` layouts (block plan given, dims supported): no [E, D] message buffer; instead the packed (hi | lo') fp16 copy
// of the source states (Ns rows) and the TMEM-layout edge weights of the fused kernel.
// out = act(y W^T + b): tensor cores (3xTF32) when the dims fit the tiles, FFMA tiles otherwise.  scratch >= tc::dense_split_bytes.
static int dense_any(const float *y, int64_t rows, int D, const float *W, const float *bias, int out_dim, int act, float *out,
                     void *scratch, cudaStream_t st, bool pack = true) {
    if (tc_enabled() && tc::supported_dense(D, out_dim)) return tc::dense_update(y, rows, D, W, bias, out_dim, act, out, scratch, st, pack);
    int rc;
    if (out_dim <= 64) {
        using Tile = GemmTile<4>;
        rc = set_smem(dense_update_kernel<4>, Tile::SMEM_BYTES);
        if (rc) return rc;
        dim3 grid((unsigned)ceil_div(rows, GEMM_BM), (unsigned)ceil_div(out_dim, Tile::BN));
        {
            TimedScope timed__(PTGNN_KERNEL_DENSE, st);
            dense_update_kernel<4><<<grid, GEMM_THREADS, Tile::SMEM_BYTES, st>>>(y, (int)rows, D, W, bias, out_dim, act, out);
        }
    } else {
        using Tile = GemmTile<8>;
        rc = set_smem(dense_update_kernel<8>, Tile::SMEM_BYTES);
        if (rc) return rc;
        dim3 grid((unsigned)ceil_div(rows, GEMM_BM), (unsigned)ceil_div(out_dim, Tile::BN));
        {
            TimedScope timed__(PTGNN_KERNEL_DENSE, st);
            dense_update_kernel<8><<<grid, GEMM_THREADS, Tile::SMEM_BYTES, st>>>(y, (int)rows, D, W, bias, out_dim, act, out);
        }
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

struct GatedWs { size_t msg, agg, p1, p2, wsplit, grupack, xpack, xpack_own, total; };
static GatedWs gated_ws_layout(int64_t N, int64_t Ns, int64_t E, int T, int H, int D, bool fused_path) {
    GatedWs w{};
    size_t o = 0;
    auto add = [&](size_t cnt) { size_t at = o; o += ws_slice(cnt, 4); return at; };
    w.msg = add(fused_path ? 4 : (size_t)E * D + 4);
    w.agg = add((size_t)N * D + 4);
    w.p1 = add((size_t)(H / 32 + 1) * 96 * D);
    w.p2 = add((size_t)(H / 32 + 1) * 96 * H);
    w.wsplit = o; o += fused_path ? fused::packed_weight_bytes(3, T, H, 0) : tc::split_edge_weights_bytes(T, D, H);
    w.grupack = o; o += tc::gru_pack_bytes(H + 32, D) + (fused_path && gruws::supported(3, H, D) ? gruws::pack_bytes(3, H, D) : 0);
    w.xpack = o; o += fused_path ? fused::packed_state_bytes(3, Ns, H) : 0;
    w.xpack_own = o; o += fused_path ? fused::packed_state_bytes(3, N, H) : 0;   // sharded run (gather_states given): this rank's rows for the GRU
    w.total = o;
    return w;
}

struct MlpWs { size_t msg, y, wsplit, dsplit, xpack, xpack_tgt, total; };
static MlpWs mlp_ws_layout(int64_t N, int64_t Ns, int64_t E, int T, int H, int D, int Hout, int use_target, bool fused_path,
                           bool separate_targets) {
    MlpWs w{};
    size_t o = 0;
    auto add = [&](size_t cnt) { size_t at = o; o += ws_slice(cnt, 4); return at; };
    w.msg = add(fused_path ? 4 : (size_t)E * D + 4);
    w.y = add((size_t)N * D + 4);
    w.wsplit = o; o += fused_path ? fused::packed_weight_bytes(3, T, H, use_target) : tc::split_edge_weights_bytes(T, D, use_target ? 2 * H : H);
    w.dsplit = o; o += tc::dense_split_bytes(Hout > 0 ? Hout : D, D);
    w.xpack = o; o += fused_path ? fused::packed_state_bytes(3, Ns, H) : 0;
    w.xpack_tgt = o; o += (fused_path && use_target && separate_targets) ? fused::packed_state_bytes(3, N, H) : 0;
    w.total = o;
    return w;
}

}  // namespace ptgnn

using namespace ptgnn;

extern "C" size_t ptgnn_b200_gated_workspace_bytes(int64_t num_nodes, int64_t num_edges, int32_t num_types,
                                                   int32_t state_dim, int32_t message_dim) {
    if (num_nodes < 0 || num_edges < 0 || num_types < 0 || state_dim <= 0 || message_dim <= 0) return 0;
    return gated_ws_layout(num_nodes, num_nodes, num_edges, num_types, state_dim, message_dim, false).total;
}

// weight cache of the tensor-core path: [split edge weights | gate-blocked GRU weights + biases]; 0 when the dims run on
// the FFMA kernels (nothing worth caching there)
static size_t gated_cache_bytes(int T, int H, int D) {
    if (!tc_enabled() || !tc::supported_message(H, D) || !tc::supported_gru(H, D)) return 0;
    return tc::split_edge_weights_bytes(T, D, H) + tc::gru_pack_bytes(H, D);
}
static bool fused_f32_ok(int H, int D) { return tc_enabled() && fused::supported(3, H, D, 0) && tc::supported_gru(H, D); }
// PTGNN_B200_GRU=tc selects the round-1 GRU pipeline (3xTF32, streams its weights per tile) behind the fused aggregation
static bool gru_ws_enabled(int nprod, int H, int D) {
    static int v = -1;
    if (v < 0) { const char *e = getenv("PTGNN_B200_GRU"); v = (e && e[0] == 't') ? 0 : 1; }
    return v == 1 && gruws::supported(nprod, H, D);
}
static size_t gated_fused_cache_bytes(int T, int H, int D) {
    return fused::packed_weight_bytes(3, T, H, 0) + (gru_ws_enabled(3, H, D) ? gruws::pack_bytes(3, H, D) : tc::gru_pack_bytes(H, D));
}

static int gated_forward_impl(const float *node_states, const float *gather_states, int64_t num_nodes, int32_t state_dim,
                              int32_t message_dim, int32_t num_types, const int64_t *type_off, const int32_t *row_ptr,
                              const int32_t *pos, const int32_t *src32, const float *const *edge_weights,
                              const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                              int32_t reduce, float *out_states, void *workspace, size_t workspace_bytes, void *weight_cache,
                              size_t weight_cache_bytes, int32_t cache_valid, void *stream,
                              const ptgnn_b200_block_plan *bp = nullptr, int64_t num_source_nodes = 0,
                              const void *packed_in = nullptr, void *packed_out = nullptr) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = state_dim, D = message_dim;
    PTGNN_CHECK_ARG(num_types >= 0 && num_types <= PTGNN_MAX_EDGE_TYPES && (type_off || bp), "gated_forward: bad num_types=%d",
                    num_types);
    const bool fused_path = bp != nullptr;
    if (fused_path && !fused_f32_ok(H, D)) {
        set_error("gated_forward_fused: dims H=%d D=%d are not supported by the fused kernel", H, D);
        return PTGNN_E_UNSUPPORTED;
    }
    const int64_t E = fused_path ? 0 : type_off[num_types];
    if (num_source_nodes <= 0) num_source_nodes = num_nodes;
    int rc = check_layer_dims("gated_forward", num_nodes, E, H, D);
    if (rc) return rc;
    if (H % 32 != 0) {
        set_error("gated_forward: state dim %d must be a multiple of 32 for the GRU kernel", H);
        return PTGNN_E_UNSUPPORTED;
    }
    PTGNN_CHECK_ARG(reduce >= PTGNN_REDUCE_SUM && reduce <= PTGNN_REDUCE_MIN, "gated_forward: bad reduce %d", reduce);
    if (num_nodes == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(node_states && out_states && row_ptr && gru_w_ih && gru_w_hh && gru_b_ih && gru_b_hh,
                    "gated_forward: null pointer");
    PTGNN_CHECK_ARG(E == 0 || (pos && src32 && edge_weights), "gated_forward: null edge arrays");
    PTGNN_CHECK_ARG(!fused_path || (bp->group_off && edge_weights && num_types > 0), "gated_forward_fused: null block plan arrays");
    const GatedWs L = gated_ws_layout(num_nodes, num_source_nodes, E, num_types, H, D, fused_path);
    if (workspace_bytes < L.total || !workspace) {
        set_error("gated_forward: workspace %zu < required %zu", workspace_bytes, L.total);
        return PTGNN_E_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    float *msg = reinterpret_cast<float *>(ws + L.msg), *agg = reinterpret_cast<float *>(ws + L.agg);
    float *P1 = reinterpret_cast<float *>(ws + L.p1), *P2 = reinterpret_cast<float *>(ws + L.p2);
    const float *gsrc = gather_states ? gather_states : node_states;   // rows that `src32` indexes (sharded runs)
    // derived weights: in the workspace (re-derived every call) or in the caller's cache (derived when !cache_valid)
    char *wsplit = ws + L.wsplit, *grupack = ws + L.grupack;
    bool pack = true;
    const size_t need_cache = fused_path ? gated_fused_cache_bytes(num_types, H, D) : gated_cache_bytes(num_types, H, D);
    if (weight_cache != nullptr && need_cache > 0) {
        if (weight_cache_bytes < need_cache) {
            set_error("gated_forward: weight cache %zu < required %zu", weight_cache_bytes, need_cache);
            return PTGNN_E_WORKSPACE;
        }
        wsplit = static_cast<char *>(weight_cache);
        grupack = wsplit + (fused_path ? fused::packed_weight_bytes(3, num_types, H, 0) : tc::split_edge_weights_bytes(num_types, D, H));
        pack = !cache_valid;
    }

    if (fused_path) {
        // 1+2. gather -> W_t -> segmented reduce in one kernel (no message buffer); fp32-exact via 3xFP16
        if (pack) {
            rc = fused::pack_weights(3, num_types, H, 0, edge_weights, wsplit, bp->status, st);
            if (rc) return rc;
        }
        // packed_in (optional): node_states already as fp16 (hi | lo') rows -- the previous layer's GRU wrote them next to its
        // fp32 output -- so the packing pass is skipped.  In a sharded run the gathered rows are a different tensor: they are
        // packed here, and packed_in (this rank's rows) feeds the GRU.
        const bool sharded = gather_states != nullptr && gather_states != node_states;
        const void *src_rows = packed_in;
        if (sharded || packed_in == nullptr) {
            rc = fused::pack_states(gsrc, num_source_nodes, H, ws + L.xpack, bp->status, st);
            if (rc) return rc;
            src_rows = ws + L.xpack;
        }
        fused::AggregateArgs a{};
        a.nprod = 3; a.src_rows = src_rows; a.tgt_rows = nullptr; a.num_nodes = num_nodes; a.K = H; a.num_types = num_types;
        a.use_target = 0; a.reduce = reduce; a.block_targets = bp->block_targets; a.group_off = bp->group_off; a.src_f = bp->src_f;
        a.tl_f = bp->tl_f; a.row_ptr = row_ptr; a.packed_weights = wsplit; a.epi = fused::Epilogue{PTGNN_ACT_NONE, nullptr, nullptr, 0.0f};
        a.status = bp->status;
        const bool ws_gru = gru_ws_enabled(3, H, D);
        a.out = agg; a.out_mode = ws_gru ? 2 : 0;
        rc = fused::aggregate(a, st);
        if (rc) return rc;
        if (!ws_gru) {
            rc = tc::gru_update(agg, node_states, num_nodes, H, D, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, out_states, grupack, pack, st);
            if (rc || packed_out == nullptr) return rc;
            return fused::pack_states(out_states, num_nodes, H, packed_out, bp->status, st);
        }
        // 3. GRUCell, weights-stationary, on the packed aggregate and the packed states (3xFP16)
        if (pack) {
            rc = gruws::pack(3, H, D, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, grupack, st);
            if (rc) return rc;
        }
        const void *h_rows = src_rows;
        if (sharded) {      // the packed copy above holds the GATHERED rows; the GRU needs this rank's
            h_rows = packed_in;
            if (h_rows == nullptr) {
                rc = fused::pack_states(node_states, num_nodes, H, ws + L.xpack_own, bp->status, st);
                if (rc) return rc;
                h_rows = ws + L.xpack_own;
            }
        }
        return gruws::update(3, agg, h_rows, node_states, num_nodes, H, D, grupack, out_states, packed_out, bp->status, st);
    }

    // 1. per-edge messages, written at their target-sorted positions
    if (tc_enabled() && tc::supported_message(H, D)) {
        rc = tc::edge_messages(gsrc, node_states, H, D, 0, num_types, type_off, edge_weights, src32, nullptr, pos, msg,
                               wsplit, pack, st);
    } else {
        rc = launch_edge_messages(gsrc, node_states, H, D, 0, num_types, type_off, edge_weights, src32, nullptr, pos, msg,
                                  st);
    }
    if (rc) return rc;
    // 2. streaming segmented reduce
    rc = launch_segment_reduce(msg, row_ptr, nullptr, num_nodes, E, D, reduce, agg, nullptr, nullptr, st);
    if (rc) return rc;
    // 3. GRUCell
    if (tc_enabled() && tc::supported_gru(H, D)) {
        return tc::gru_update(agg, node_states, num_nodes, H, D, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, out_states,
                              grupack, pack, st);
    }
    {
        TimedScope timed__(PTGNN_KERNEL_PACK, st);
        pack_gru_weights_kernel<<<148, 256, 0, st>>>(gru_w_ih, gru_w_hh, H, D, P1, P2);
    }
    PTGNN_LAUNCHED();
    using Tile = GemmTile<6>;
    rc = set_smem(gru_update_kernel, Tile::SMEM_BYTES);
    if (rc) return rc;
    dim3 grid((unsigned)ceil_div(num_nodes, GEMM_BM), H / 32);
    {
        TimedScope timed__(PTGNN_KERNEL_GRU, st);
        gru_update_kernel<<<grid, GEMM_THREADS, Tile::SMEM_BYTES, st>>>(agg, node_states, (int)num_nodes, H, D, P1, P2,
                                                                     gru_b_ih, gru_b_hh, out_states);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

extern "C" int ptgnn_b200_gated_forward_f32(const float *node_states, const float *gather_states, int64_t num_nodes,
                                            int32_t state_dim, int32_t message_dim, int32_t num_types,
                                            const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                            const int32_t *src32, const float *const *edge_weights, const float *gru_w_ih,
                                            const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                                            int32_t reduce, float *out_states, void *workspace, size_t workspace_bytes,
                                            void *stream) {
    return gated_forward_impl(node_states, gather_states, num_nodes, state_dim, message_dim, num_types, type_off, row_ptr, pos,
                              src32, edge_weights, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, reduce, out_states, workspace,
                              workspace_bytes, nullptr, 0, 0, stream);
}

extern "C" size_t ptgnn_b200_gated_weight_cache_bytes(int32_t num_types, int32_t state_dim, int32_t message_dim) {
    if (num_types < 0 || num_types > PTGNN_MAX_EDGE_TYPES || state_dim <= 0 || message_dim <= 0) return 0;
    return gated_cache_bytes(num_types, state_dim, message_dim);
}

extern "C" int ptgnn_b200_gated_forward_cached_f32(const float *node_states, const float *gather_states, int64_t num_nodes,
                                                   int32_t state_dim, int32_t message_dim, int32_t num_types,
                                                   const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                                   const int32_t *src32, const float *const *edge_weights,
                                                   const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih,
                                                   const float *gru_b_hh, int32_t reduce, float *out_states, void *workspace,
                                                   size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                                   int32_t cache_valid, void *stream) {
    return gated_forward_impl(node_states, gather_states, num_nodes, state_dim, message_dim, num_types, type_off, row_ptr, pos,
                              src32, edge_weights, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, reduce, out_states, workspace,
                              workspace_bytes, weight_cache, weight_cache_bytes, cache_valid, stream);
}

extern "C" size_t ptgnn_b200_mlp_workspace_bytes(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t in_dim,
                                                 int32_t message_dim, int32_t out_dim, int32_t use_target_state) {
    if (num_nodes < 0 || num_edges < 0 || num_types < 0 || in_dim <= 0 || message_dim <= 0) return 0;
    return mlp_ws_layout(num_nodes, num_nodes, num_edges, num_types, in_dim, message_dim, out_dim, use_target_state, false, false).total;
}

static size_t mlp_fused_cache_bytes(int T, int H, int D, int out_dim, int ut) {
    return ws_slice(fused::packed_weight_bytes(3, T, H, ut), 1) + tc::dense_split_bytes(out_dim > 0 ? out_dim : D, D) + 256;
}

static int mlp_forward_impl(const float *node_states, const float *gather_states, int64_t num_nodes,
                                          int32_t in_dim,
                                          int32_t message_dim, int32_t out_dim, int32_t num_types,
                                          const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                          const int32_t *src32, const int32_t *tgt32, const float *const *edge_weights,
                                          int32_t use_target_state, int32_t reduce, int32_t message_activation,
                                          const float *ln_weight, const float *ln_bias, float ln_eps,
                                          const float *dense_weight, const float *dense_bias, int32_t dense_activation,
                                          float *out_states, void *workspace, size_t workspace_bytes, void *stream,
                                          const ptgnn_b200_block_plan *bp, int64_t num_source_nodes, void *weight_cache = nullptr,
                                          size_t weight_cache_bytes = 0, int cache_valid = 0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = in_dim, D = message_dim;
    PTGNN_CHECK_ARG(num_types >= 0 && num_types <= PTGNN_MAX_EDGE_TYPES && (type_off || bp), "mlp_forward: bad num_types=%d",
                    num_types);
    const bool fused_path = bp != nullptr;
    if (fused_path && !(tc_enabled() && fused::supported(3, H, D, use_target_state))) {
        set_error("mlp_forward_fused: dims H=%d D=%d are not supported by the fused kernel", H, D);
        return PTGNN_E_UNSUPPORTED;
    }
    const int64_t E = fused_path ? 0 : type_off[num_types];
    if (num_source_nodes <= 0) num_source_nodes = num_nodes;
    int rc = check_layer_dims("mlp_forward", num_nodes, E, H, D);
    if (rc) return rc;
    PTGNN_CHECK_ARG(reduce >= PTGNN_REDUCE_SUM && reduce <= PTGNN_REDUCE_MIN, "mlp_forward: bad reduce %d", reduce);
    PTGNN_CHECK_ARG(message_activation >= PTGNN_ACT_NONE && message_activation <= PTGNN_ACT_RELU &&
                        dense_activation >= PTGNN_ACT_NONE && dense_activation <= PTGNN_ACT_RELU,
                    "mlp_forward: bad activation");
    PTGNN_CHECK_ARG((ln_weight == nullptr) == (ln_bias == nullptr), "mlp_forward: ln_weight/ln_bias must both be set");
    PTGNN_CHECK_ARG(dense_weight ? out_dim > 0 : out_dim == D, "mlp_forward: out_dim=%d inconsistent", out_dim);
    if (num_nodes == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(node_states && out_states && row_ptr, "mlp_forward: null pointer");
    PTGNN_CHECK_ARG(E == 0 || (pos && src32 && edge_weights && (!use_target_state || tgt32)),
                    "mlp_forward: null edge arrays");
    PTGNN_CHECK_ARG(!fused_path || (bp->group_off && edge_weights && num_types > 0), "mlp_forward_fused: null block plan arrays");
    const MlpWs L = mlp_ws_layout(num_nodes, num_source_nodes, E, num_types, H, D, out_dim, use_target_state, fused_path,
                                  gather_states != nullptr);
    if (workspace_bytes < L.total || !workspace) {
        set_error("mlp_forward: workspace %zu < required %zu", workspace_bytes, L.total);
        return PTGNN_E_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    float *msg = reinterpret_cast<float *>(ws + L.msg);
    float *y = dense_weight ? reinterpret_cast<float *>(ws + L.y) : out_states;
    const int ut = use_target_state ? 1 : 0;
    const float *gsrc = gather_states ? gather_states : node_states;   // rows that `src32` indexes (sharded runs)

    // derived weights: in the workspace (re-derived every call) or in the caller's cache (derived when !cache_valid); fused path only
    char *wsplit = ws + L.wsplit, *dsplit = ws + L.dsplit;
    bool pack = true;
    if (fused_path && weight_cache != nullptr) {
        const size_t need = mlp_fused_cache_bytes(num_types, H, D, out_dim, ut);
        if (weight_cache_bytes < need) {
            set_error("mlp_forward_fused: weight cache %zu < required %zu", weight_cache_bytes, need);
            return PTGNN_E_WORKSPACE;
        }
        wsplit = static_cast<char *>(weight_cache);
        dsplit = wsplit + ws_slice(fused::packed_weight_bytes(3, num_types, H, ut), 1);
        pack = !cache_valid;
    }
    if (fused_path) {
        if (pack) {
            rc = fused::pack_weights(3, num_types, H, ut, edge_weights, wsplit, bp->status, st);
            if (rc) return rc;
        }
        rc = fused::pack_states(gsrc, num_source_nodes, H, ws + L.xpack, bp->status, st);
        if (rc) return rc;
        const void *tgt_rows = ws + L.xpack;
        if (ut && gather_states != nullptr) {      // sharded run: targets are this rank's rows, not the gathered ones
            rc = fused::pack_states(node_states, num_nodes, H, ws + L.xpack_tgt, bp->status, st);
            if (rc) return rc;
            tgt_rows = ws + L.xpack_tgt;
        }
        fused::AggregateArgs a{};
        a.nprod = 3; a.src_rows = ws + L.xpack; a.tgt_rows = tgt_rows; a.num_nodes = num_nodes; a.K = H; a.num_types = num_types;
        a.use_target = ut; a.reduce = reduce; a.block_targets = bp->block_targets; a.group_off = bp->group_off; a.src_f = bp->src_f;
        a.tl_f = bp->tl_f; a.row_ptr = row_ptr; a.packed_weights = wsplit;
        a.epi = fused::Epilogue{message_activation, ln_weight, ln_bias, ln_eps};
        a.out = y; a.out_mode = 0; a.status = bp->status;
        rc = fused::aggregate(a, st);
        if (rc) return rc;
    } else {
        if (tc_enabled() && tc::supported_message(H, D)) {
            rc = tc::edge_messages(gsrc, node_states, H, D, ut, num_types, type_off, edge_weights, src32, tgt32, pos, msg,
                                   ws + L.wsplit, true, st);
        } else {
            rc = launch_edge_messages(gsrc, node_states, H, D, ut, num_types, type_off, edge_weights, src32, tgt32, pos, msg,
                                      st);
        }
        if (rc) return rc;
        ReduceEpilogue epi{0, message_activation, ln_weight, ln_bias, ln_eps};
        rc = launch_segment_reduce(msg, row_ptr, nullptr, num_nodes, E, D, reduce, y, nullptr, &epi, st);
        if (rc) return rc;
    }
    if (!dense_weight) return PTGNN_OK;
    return dense_any(y, num_nodes, D, dense_weight, dense_bias, out_dim, dense_activation, out_states, dsplit, st, pack);
}

extern "C" int ptgnn_b200_mlp_forward_f32(const float *node_states, const float *gather_states, int64_t num_nodes,
                                          int32_t in_dim, int32_t message_dim, int32_t out_dim, int32_t num_types,
                                          const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                          const int32_t *src32, const int32_t *tgt32, const float *const *edge_weights,
                                          int32_t use_target_state, int32_t reduce, int32_t message_activation,
                                          const float *ln_weight, const float *ln_bias, float ln_eps,
                                          const float *dense_weight, const float *dense_bias, int32_t dense_activation,
                                          float *out_states, void *workspace, size_t workspace_bytes, void *stream) {
    return mlp_forward_impl(node_states, gather_states, num_nodes, in_dim, message_dim, out_dim, num_types, type_off, row_ptr, pos,
                            src32, tgt32, edge_weights, use_target_state, reduce, message_activation, ln_weight, ln_bias, ln_eps,
                            dense_weight, dense_bias, dense_activation, out_states, workspace, workspace_bytes, stream, nullptr, 0);
}

// ---- fused entry points (fp32 states here, bf16 states in layers_bf16.cu) --------------------------------------------------
namespace ptgnn {
namespace tcb {
size_t gated_fused_workspace_bytes_bf16(int64_t N, int T, int H, int D);
size_t gated_fused_cache_bytes_bf16(int T, int H, int D);
int gated_forward_fused_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes, int32_t state_dim,
                             int32_t message_dim, int32_t num_types, const ptgnn_b200_block_plan *bp, const int32_t *row_ptr,
                             const float *const *edge_weights, const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih,
                             const float *gru_b_hh, int32_t reduce, uint16_t *out_states, void *workspace, size_t workspace_bytes,
                             void *weight_cache, size_t weight_cache_bytes, int32_t cache_valid, void *stream);
size_t mlp_fused_workspace_bytes_bf16(int64_t N, int T, int H, int D, int Hout, int use_target);
int mlp_forward_fused_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes, int32_t in_dim,
                           int32_t message_dim, int32_t out_dim, int32_t num_types, const ptgnn_b200_block_plan *bp,
                           const int32_t *row_ptr, const float *const *edge_weights, int32_t use_target_state, int32_t reduce,
                           int32_t message_activation, const float *ln_weight, const float *ln_bias, float ln_eps,
                           const float *dense_weight, const float *dense_bias, int32_t dense_activation, uint16_t *out_states,
                           void *workspace, size_t workspace_bytes, void *stream);
}  // namespace tcb
}  // namespace ptgnn

extern "C" int32_t ptgnn_b200_block_plan_block_targets(int64_t num_nodes) { return fused::recommended_block_targets(num_nodes); }

extern "C" int32_t ptgnn_b200_fused_supported(int32_t bf16_states, int32_t state_dim, int32_t message_dim) {
    if (!tc_enabled()) return 0;
    if (bf16_states) return fused::supported(1, state_dim, message_dim, 0) && state_dim % 32 == 0 ? 1 : 0;
    return fused_f32_ok(state_dim, message_dim) ? 1 : 0;
}

extern "C" size_t ptgnn_b200_gated_fused_workspace_bytes(int32_t bf16_states, int64_t num_nodes, int64_t num_source_nodes,
                                                         int32_t num_types, int32_t state_dim, int32_t message_dim) {
    if (num_nodes < 0 || num_types < 0 || state_dim <= 0 || message_dim <= 0) return 0;
    if (num_source_nodes <= 0) num_source_nodes = num_nodes;
    if (bf16_states) return tcb::gated_fused_workspace_bytes_bf16(num_nodes, num_types, state_dim, message_dim);
    return gated_ws_layout(num_nodes, num_source_nodes, 0, num_types, state_dim, message_dim, true).total;
}
extern "C" size_t ptgnn_b200_gated_fused_weight_cache_bytes(int32_t bf16_states, int32_t num_types, int32_t state_dim,
                                                            int32_t message_dim) {
    if (num_types < 0 || num_types > PTGNN_MAX_EDGE_TYPES || state_dim <= 0 || message_dim <= 0) return 0;
    if (bf16_states) return tcb::gated_fused_cache_bytes_bf16(num_types, state_dim, message_dim);
    return gated_fused_cache_bytes(num_types, state_dim, message_dim);
}
extern "C" int ptgnn_b200_gated_forward_fused(int32_t bf16_states, const void *node_states, const void *gather_states,
                                              int64_t num_nodes, int64_t num_source_nodes, int32_t state_dim, int32_t message_dim,
                                              int32_t num_types, const ptgnn_b200_block_plan *block_plan, const int32_t *row_ptr,
                                              const float *const *edge_weights, const float *gru_w_ih, const float *gru_w_hh,
                                              const float *gru_b_ih, const float *gru_b_hh, int32_t reduce, void *out_states,
                                              void *workspace, size_t workspace_bytes, void *weight_cache,
                                              size_t weight_cache_bytes, int32_t cache_valid, void *stream) {
    PTGNN_CHECK_ARG(block_plan != nullptr, "gated_forward_fused: null block plan");
    if (bf16_states)
        return tcb::gated_forward_fused_bf16(static_cast<const uint16_t *>(node_states), static_cast<const uint16_t *>(gather_states),
                                             num_nodes, state_dim, message_dim, num_types, block_plan, row_ptr, edge_weights, gru_w_ih,
                                             gru_w_hh, gru_b_ih, gru_b_hh, reduce, static_cast<uint16_t *>(out_states), workspace,
                                             workspace_bytes, weight_cache, weight_cache_bytes, cache_valid, stream);
    return gated_forward_impl(static_cast<const float *>(node_states), static_cast<const float *>(gather_states), num_nodes, state_dim,
                              message_dim, num_types, nullptr, row_ptr, nullptr, nullptr, edge_weights, gru_w_ih, gru_w_hh, gru_b_ih,
                              gru_b_hh, reduce, static_cast<float *>(out_states), workspace, workspace_bytes, weight_cache,
                              weight_cache_bytes, cache_valid, stream, block_plan, num_source_nodes);
}

extern "C" size_t ptgnn_b200_packed_state_bytes(int64_t num_nodes, int32_t state_dim) {
    if (num_nodes < 0 || state_dim <= 0) return 0;
    return fused::packed_state_bytes(3, num_nodes, state_dim);
}
extern "C" int ptgnn_b200_gated_forward_fused_chained(const float *node_states, const float *gather_states, const void *packed_states_in,
                                                      int64_t num_nodes, int64_t num_source_nodes, int32_t state_dim,
                                                      int32_t message_dim, int32_t num_types, const ptgnn_b200_block_plan *block_plan,
                                                      const int32_t *row_ptr, const float *const *edge_weights, const float *gru_w_ih,
                                                      const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                                                      int32_t reduce, float *out_states, void *packed_states_out, void *workspace,
                                                      size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                                      int32_t cache_valid, void *stream) {
    PTGNN_CHECK_ARG(block_plan != nullptr, "gated_forward_fused_chained: null block plan");
    return gated_forward_impl(node_states, gather_states, num_nodes, state_dim, message_dim, num_types, nullptr, row_ptr, nullptr, nullptr,
                              edge_weights, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, reduce, out_states, workspace, workspace_bytes,
                              weight_cache, weight_cache_bytes, cache_valid, stream, block_plan, num_source_nodes, packed_states_in,
                              packed_states_out);
}

extern "C" size_t ptgnn_b200_mlp_fused_workspace_bytes(int32_t bf16_states, int64_t num_nodes, int64_t num_source_nodes,
                                                       int32_t num_types, int32_t in_dim, int32_t message_dim, int32_t out_dim,
                                                       int32_t use_target_state) {
    if (num_nodes < 0 || num_types < 0 || in_dim <= 0 || message_dim <= 0) return 0;
    if (num_source_nodes <= 0) num_source_nodes = num_nodes;
    if (bf16_states) return tcb::mlp_fused_workspace_bytes_bf16(num_nodes, num_types, in_dim, message_dim, out_dim, use_target_state);
    return mlp_ws_layout(num_nodes, num_source_nodes, 0, num_types, in_dim, message_dim, out_dim, use_target_state, true, true).total;
}
extern "C" int ptgnn_b200_mlp_forward_fused(int32_t bf16_states, const void *node_states, const void *gather_states,
                                            int64_t num_nodes, int64_t num_source_nodes, int32_t in_dim, int32_t message_dim,
                                            int32_t out_dim, int32_t num_types, const ptgnn_b200_block_plan *block_plan,
                                            const int32_t *row_ptr, const float *const *edge_weights, int32_t use_target_state,
                                            int32_t reduce, int32_t message_activation, const float *ln_weight,
                                            const float *ln_bias, float ln_eps, const float *dense_weight, const float *dense_bias,
                                            int32_t dense_activation, void *out_states, void *workspace, size_t workspace_bytes,
                                            void *stream) {
    PTGNN_CHECK_ARG(block_plan != nullptr, "mlp_forward_fused: null block plan");
    if (bf16_states)
        return tcb::mlp_forward_fused_bf16(static_cast<const uint16_t *>(node_states), static_cast<const uint16_t *>(gather_states),
                                           num_nodes, in_dim, message_dim, out_dim, num_types, block_plan, row_ptr, edge_weights,
                                           use_target_state, reduce, message_activation, ln_weight, ln_bias, ln_eps, dense_weight,
                                           dense_bias, dense_activation, static_cast<uint16_t *>(out_states), workspace,
                                           workspace_bytes, stream);
    return mlp_forward_impl(static_cast<const float *>(node_states), static_cast<const float *>(gather_states), num_nodes, in_dim,
                            message_dim, out_dim, num_types, nullptr, row_ptr, nullptr, nullptr, nullptr, edge_weights,
                            use_target_state, reduce, message_activation, ln_weight, ln_bias, ln_eps, dense_weight, dense_bias,
                            dense_activation, static_cast<float *>(out_states), workspace, workspace_bytes, stream, block_plan,
                            num_source_nodes);
}

extern "C" size_t ptgnn_b200_mlp_fused_weight_cache_bytes(int32_t bf16_states, int32_t num_types, int32_t in_dim, int32_t message_dim,
                                                          int32_t out_dim, int32_t use_target_state) {
    if (bf16_states || num_types <= 0 || in_dim <= 0 || message_dim <= 0) return 0;
    if (!(tc_enabled() && fused::supported(3, in_dim, message_dim, use_target_state))) return 0;
    return mlp_fused_cache_bytes(num_types, in_dim, message_dim, out_dim, use_target_state ? 1 : 0);
}
extern "C" int ptgnn_b200_mlp_forward_fused_cached(int32_t bf16_states, const void *node_states, const void *gather_states, int64_t num_nodes,
                                                   int64_t num_source_nodes, int32_t in_dim, int32_t message_dim, int32_t out_dim,
                                                   int32_t num_types, const ptgnn_b200_block_plan *block_plan, const int32_t *row_ptr,
                                                   const float *const *edge_weights, int32_t use_target_state, int32_t reduce,
                                                   int32_t message_activation, const float *ln_weight, const float *ln_bias, float ln_eps,
                                                   const float *dense_weight, const float *dense_bias, int32_t dense_activation,
                                                   void *out_states, void *workspace, size_t workspace_bytes, void *weight_cache,
                                                   size_t weight_cache_bytes, int32_t cache_valid, void *stream) {
    if (bf16_states || weight_cache == nullptr)
        return ptgnn_b200_mlp_forward_fused(bf16_states, node_states, gather_states, num_nodes, num_source_nodes, in_dim, message_dim, out_dim,
                                            num_types, block_plan, row_ptr, edge_weights, use_target_state, reduce, message_activation,
                                            ln_weight, ln_bias, ln_eps, dense_weight, dense_bias, dense_activation, out_states, workspace,
                                            workspace_bytes, stream);
    PTGNN_CHECK_ARG(block_plan != nullptr, "mlp_forward_fused_cached: null block plan");
    return mlp_forward_impl(static_cast<const float *>(node_states), static_cast<const float *>(gather_states), num_nodes, in_dim, message_dim,
                            out_dim, num_types, nullptr, row_ptr, nullptr, nullptr, nullptr, edge_weights, use_target_state, reduce,
                            message_activation, ln_weight, ln_bias, ln_eps, dense_weight, dense_bias, dense_activation,
                            static_cast<float *>(out_states), workspace, workspace_bytes, stream, block_plan, num_source_nodes, weight_cache,
                            weight_cache_bytes, cache_valid);
}

/* ---- stand-alone pieces (MLP.forward, message MLPs with hidden layers, module aggregators) -------------------------------------- */
extern "C" size_t ptgnn_b200_linear_workspace_bytes(int32_t in_dim, int32_t out_dim) {
    if (in_dim <= 0 || out_dim <= 0) return 0;
    return tc::dense_split_bytes(out_dim, in_dim) + 256;
}
extern "C" int ptgnn_b200_linear_f32(const float *x, int64_t rows, int32_t in_dim, const float *weight, const float *bias,
                                     int32_t out_dim, int32_t activation, float *out, void *workspace, size_t workspace_bytes,
                                     void *stream) {
    PTGNN_CHECK_ARG(rows >= 0 && rows < INT32_MAX, "linear: rows out of range");
    PTGNN_CHECK_ARG(in_dim > 0 && in_dim % 4 == 0 && out_dim > 0 && out_dim % 4 == 0 && in_dim <= 4096 && out_dim <= 4096,
                    "linear: dims %d -> %d must be multiples of 4 (<= 4096)", in_dim, out_dim);
    PTGNN_CHECK_ARG(activation >= PTGNN_ACT_NONE && activation <= PTGNN_ACT_RELU, "linear: bad activation %d", activation);
    if (rows == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(x && weight && out, "linear: null pointer");
    if (workspace_bytes < ptgnn_b200_linear_workspace_bytes(in_dim, out_dim) || !workspace) {
        set_error("linear: workspace too small");
        return PTGNN_E_WORKSPACE;
    }
    return dense_any(x, rows, in_dim, weight, bias, out_dim, activation, out, workspace, static_cast<cudaStream_t>(stream));
}

extern "C" size_t ptgnn_b200_edge_messages_workspace_bytes(int32_t num_types, int32_t in_dim, int32_t message_dim, int32_t use_target_state) {
    if (num_types < 0 || in_dim <= 0 || message_dim <= 0) return 0;
    return tc::split_edge_weights_bytes(num_types, message_dim, use_target_state ? 2 * in_dim : in_dim) + 256;
}
extern "C" int ptgnn_b200_edge_messages_f32(const float *source_states, const float *target_states, int32_t in_dim, int32_t message_dim,
                                            int32_t num_types, const int64_t *type_off, const int32_t *src32, const int32_t *tgt32,
                                            const int32_t *out_row, const float *const *edge_weights, int32_t use_target_state,
                                            float *messages, void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PTGNN_CHECK_ARG(num_types >= 0 && num_types <= PTGNN_MAX_EDGE_TYPES && type_off, "edge_messages: bad num_types=%d", num_types);
    const int64_t E = type_off[num_types];
    int rc = check_layer_dims("edge_messages", 0, E, in_dim, message_dim);
    if (rc) return rc;
    if (E == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(source_states && src32 && out_row && edge_weights && messages && (!use_target_state || (target_states && tgt32)),
                    "edge_messages: null pointer");
    if (workspace_bytes < ptgnn_b200_edge_messages_workspace_bytes(num_types, in_dim, message_dim, use_target_state) || !workspace) {
        set_error("edge_messages: workspace too small");
        return PTGNN_E_WORKSPACE;
    }
    const int ut = use_target_state ? 1 : 0;
    if (tc_enabled() && tc::supported_message(in_dim, message_dim))
        return tc::edge_messages(source_states, target_states, in_dim, message_dim, ut, num_types, type_off, edge_weights, src32, tgt32,
                                 out_row, messages, workspace, true, st);
    return launch_edge_messages(source_states, target_states, in_dim, message_dim, ut, num_types, type_off, edge_weights, src32, tgt32,
                                out_row, messages, st);
}

extern "C" size_t ptgnn_b200_grucell_workspace_bytes(int32_t state_dim, int32_t input_dim) {
    if (state_dim <= 0 || input_dim <= 0) return 0;
    return ws_slice((size_t)(state_dim / 32 + 1) * 96 * input_dim, 4) + ws_slice((size_t)(state_dim / 32 + 1) * 96 * state_dim, 4) +
           tc::gru_pack_bytes(state_dim + 32, input_dim) + 256;
}
extern "C" int ptgnn_b200_grucell_f32(const float *input, const float *hidden, int64_t rows, int32_t state_dim, int32_t input_dim,
                                      const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh, float *out,
                                      void *workspace, size_t workspace_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = state_dim, D = input_dim;
    int rc = check_layer_dims("grucell", rows, 0, H, D);
    if (rc) return rc;
    if (H % 32 != 0) { set_error("grucell: state dim %d must be a multiple of 32", H); return PTGNN_E_UNSUPPORTED; }
    if (rows == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(input && hidden && w_ih && w_hh && b_ih && b_hh && out, "grucell: null pointer");
    if (workspace_bytes < ptgnn_b200_grucell_workspace_bytes(H, D) || !workspace) { set_error("grucell: workspace too small"); return PTGNN_E_WORKSPACE; }
    char *ws = static_cast<char *>(workspace);
    const size_t o1 = ws_slice((size_t)(H / 32 + 1) * 96 * D, 4), o2 = ws_slice((size_t)(H / 32 + 1) * 96 * H, 4);
    if (tc_enabled() && tc::supported_gru(H, D))
        return tc::gru_update(input, hidden, rows, H, D, w_ih, w_hh, b_ih, b_hh, out, ws + o1 + o2, true, st);
    float *P1 = reinterpret_cast<float *>(ws), *P2 = reinterpret_cast<float *>(ws + o1);
    {
        TimedScope timed__(PTGNN_KERNEL_PACK, st);
        pack_gru_weights_kernel<<<148, 256, 0, st>>>(w_ih, w_hh, H, D, P1, P2);
    }
    PTGNN_LAUNCHED();
    using Tile = GemmTile<6>;
    rc = set_smem(gru_update_kernel, Tile::SMEM_BYTES);
    if (rc) return rc;
    dim3 grid((unsigned)ceil_div(rows, GEMM_BM), H / 32);
    {
        TimedScope timed__(PTGNN_KERNEL_GRU, st);
        gru_update_kernel<<<grid, GEMM_THREADS, Tile::SMEM_BYTES, st>>>(input, hidden, (int)rows, H, D, P1, P2, b_ih, b_hh, out);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}
