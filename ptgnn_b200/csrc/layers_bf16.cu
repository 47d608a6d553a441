// bf16 GatedMessagePassingLayer forward (BASELINE.json configs[3]: bf16 states): bf16 node states / messages / weights,
// fp32 accumulation everywhere (tensor-core accumulators, segmented reduce, gate math) -- the arithmetic of the
// reference under torch.autocast(bfloat16), whose scatter is always fp32 (abstractmessagepassing.py:43-50).
//   reference ptgnn/neuralmodels/gnn/messagepassing/gatedmessagepassing.py:37-69
// Kernels: weight conversion/packing -> tc_pipeline_bf16_kernel<MsgPolicyB> -> segment_reduce_bf16_kernel ->
// tc_pipeline_bf16_kernel<GruPolicyB>.
#include <float.h>
#include <stdlib.h>

#include "layers_tc.cuh"
#include "fused_mp.cuh"
#include "gru_ws.cuh"
#include "tc_pipeline_bf16.cuh"

namespace ptgnn {
namespace tcb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
// bf16 row-major [rows, cols], box = {64 columns (128 bytes), box_rows}, SWIZZLE_128B, OOB zero fill
static int make_map_bf16(CUtensorMap *map, const __nv_bfloat16 *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return PTGNN_E_CUDA; }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)CHUNK_K, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16 *>(base), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r); return PTGNN_E_CUDA; }
    return PTGNN_OK;
}

// ---- weights: fp32 module parameters -> bf16 working copies ---------------------------------------------------
struct ConvSrc {
    const float *w[PTGNN_MAX_EDGE_TYPES];
    int num, elems;
};
__global__ void convert_weights_kernel(const __grid_constant__ ConvSrc s, __nv_bfloat16 *__restrict__ out) {
    const int64_t total = (int64_t)s.num * s.elems;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(s.w[i / s.elems][i % s.elems]);
}
// same gate-blocked layout as the fp32 path: P1[jb] = [W_ir; W_iz; W_in; 0], P2[jb] = [W_hr; W_hz; 0; W_hn] (128 rows each)
__global__ void pack_gru_bias_bf16_kernel(const float *__restrict__ b_ih, const float *__restrict__ b_hh, int H, float4 *__restrict__ bias4) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < H) bias4[j] = make_float4(b_ih[j] + b_hh[j], b_ih[H + j] + b_hh[H + j], b_ih[2 * H + j], b_hh[2 * H + j]);
}
__global__ void pack_gru_bf16_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh, int H, int D,
                                     __nv_bfloat16 *__restrict__ p1, __nv_bfloat16 *__restrict__ p2) {
    const int nblk = H / 32;
    const int64_t n1 = (int64_t)nblk * 128 * D, n2 = (int64_t)nblk * 128 * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n1) {
            const int k = (int)(i % D), n = (int)((i / D) % 128), jb = (int)(i / ((int64_t)128 * D));
            const int gate = n / 32;
            p1[i] = __float2bfloat16_rn(gate < 3 ? w_ih[(size_t)(gate * H + jb * 32 + n % 32) * D + k] : 0.0f);
        } else {
            const int64_t r = i - n1;
            const int k = (int)(r % H), n = (int)((r / H) % 128), jb = (int)(r / ((int64_t)128 * H));
            const int gate = n / 32;
            p2[r] = __float2bfloat16_rn(gate == 2 ? 0.0f : w_hh[(size_t)((gate == 3 ? 2 : gate) * H + jb * 32 + n % 32) * H + k]);
        }
    }
}

// ---- policy: per-edge messages -------------------------------------------------------------------------------
struct MsgPolicyB {
    struct Params {
        CUtensorMap map_w;                 // [T*D, Kw] bf16 (Kw = H, or 2H with target states), box {64, min(128, D)}
        const __nv_bfloat16 *h, *h_tgt;    // rows indexed by src32 / by tgt32 (Mlp layers with use_target_state)
        const int32_t *src32, *tgt32, *pos;
        int use_target;
        __nv_bfloat16 *msg;                // [E, D] bf16 at target-sorted rows
        unsigned long long *trace;
        int H, D, num_types, n_blocks, dbg;   // dbg: PTGNN_TC_DEBUG ablation bits (1 no MMA, 2 no store, 4 no loads, 8 no drain)
        int32_t edge_off[PTGNN_MAX_EDGE_TYPES + 1];
        int32_t tile_off[PTGNN_MAX_EDGE_TYPES + 1];
    };
    struct Tile { int t, e0, e_end, n0, b_rows; };
    __device__ static int num_tiles(const Params &p) { return p.tile_off[p.num_types] * p.n_blocks; }
    // Tiles are visited in increasing order by every role, so the edge type only ever moves forward from the previous
    // tile's: an amortised O(1) walk over tile_off instead of a binary search of dependent constant loads per tile.
    __device__ static void tile_init(Tile &ti) { ti.t = 0; }
    __device__ static void tile_setup(const Params &p, int tile, Tile &ti) {
        int mt = tile, nb = 0;
        if (p.n_blocks > 1) { mt = tile / p.n_blocks; nb = tile - mt * p.n_blocks; }
        int t = ti.t;
        while (p.tile_off[t + 1] <= mt) ++t;
        ti.t = t;
        ti.e0 = p.edge_off[t] + (mt - p.tile_off[t]) * TILE_M;
        ti.e_end = p.edge_off[t + 1];
        ti.n0 = nb * 128;
        ti.b_rows = min(128, p.D - ti.n0);
    }
    __device__ static int num_segments(const Params &p, const Tile &) { return p.use_target ? 2 : 1; }
    __device__ static Segment segment(const Params &p, const Tile &ti, int seg) {
        Segment s;
        s.a = seg == 0 ? p.h : p.h_tgt; s.lda = p.H; s.K = p.H; s.a_map = nullptr; s.a_row0 = 0;
        s.b_map = &p.map_w; s.b_row0 = ti.t * p.D + ti.n0; s.b_col0 = seg * p.H; s.b_box_rows = min(128, p.D);
        return s;
    }
    __device__ static int gather_row(const Params &p, const Tile &ti, int seg, int r) {
        const int e = ti.e0 + r;
        if (e >= ti.e_end) return -1;
        return seg == 0 ? p.src32[e] : p.tgt32[e];
    }
    __device__ static int mma_groups(const Params &, const Tile &ti, int seg, MmaGroup (&g)[2]) {
        g[0] = MmaGroup{ti.b_rows, 0, 0, seg == 0};
        return 1;
    }
    __device__ static void drain(const Params &, const Tile &ti, uint32_t tmem_lane, int half, float (&acc)[64]) {
        drain_2x32(tmem_lane, 64 * half, ti.b_rows, acc);
    }
    // only the raw load is issued a tile ahead: arithmetic on the loaded value would stall the in-order issue right there
    struct Pre { int32_t pos; };
    __device__ static void prefetch(const Params &p, const Tile &ti, int quarter, int, int lane, Pre &pre) {
        const int e = ti.e0 + quarter * 32 + lane;
        pre.pos = -1;
        if (e < ti.e_end) pre.pos = __ldg(p.pos + e);
    }
    __device__ static void smem_init(const Params &, float *) {}
    __device__ static void store(const Params &p, const Tile &ti, float (&acc)[64], const Pre &pre, int half, int lane, float *stage, float *) {
        // offsets are in 4-byte words of the bf16 message array (2 bf16 per word)
        const long long row_off = pre.pos >= 0 ? ((long long)pre.pos * p.D + ti.n0) / 2 : -1;
        const int c0 = 64 * half;            // this warp's 64 accumulator columns -> 32 packed words
        if (c0 >= ti.b_rows) return;
        float w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = pack_bf16x2(acc[2 * i], acc[2 * i + 1]);
        float *dst = reinterpret_cast<float *>(p.msg) + c0 / 2;
        if (ti.b_rows - c0 >= 64) tc::warp_store_rows<32>(stage, w, dst, row_off, lane);
        else if (ti.b_rows - c0 >= 32) tc::warp_store_rows<16>(stage, w, dst, row_off, lane);
        else tc::warp_store_rows<8>(stage, w, dst, row_off, lane);   // 16 columns
    }
};

// ---- policy: GRUCell ---------------------------------------------------------------------------------------------
struct GruPolicyB {
    struct Params {
        CUtensorMap map_agg, map_h, map_p1, map_p2;
        const __nv_bfloat16 *h;
        const float4 *bias4;   // (b_ir + b_hr, b_iz + b_hz, b_in, b_hn) per hidden unit
        __nv_bfloat16 *out;
        unsigned long long *trace;
        int num_nodes, H, D, n_jb, dbg;
    };
    struct Tile { int row0, jb; };
    __device__ static int num_tiles(const Params &p) { return ((p.num_nodes + TILE_M - 1) / TILE_M) * p.n_jb; }
    __device__ static void tile_init(Tile &) {}
    __device__ static void tile_setup(const Params &p, int tile, Tile &ti) {
        const int rb = tile / p.n_jb;
        ti.row0 = rb * TILE_M;
        ti.jb = tile - rb * p.n_jb;
    }
    __device__ static int num_segments(const Params &, const Tile &) { return 2; }
    __device__ static Segment segment(const Params &p, const Tile &ti, int seg) {
        Segment s;
        s.a = nullptr; s.lda = 0; s.a_row0 = ti.row0; s.b_row0 = ti.jb * 128; s.b_col0 = 0; s.b_box_rows = 128;
        if (seg == 0) { s.a_map = &p.map_agg; s.K = p.D; s.b_map = &p.map_p1; }
        else { s.a_map = &p.map_h; s.K = p.H; s.b_map = &p.map_p2; }
        return s;
    }
    __device__ static int gather_row(const Params &, const Tile &, int, int) { return -1; }
    __device__ static int mma_groups(const Params &, const Tile &, int seg, MmaGroup (&g)[2]) {
        g[0] = MmaGroup{128, 0, 0, seg == 0};
        return 1;
    }
    __device__ static void drain(const Params &, const Tile &, uint32_t tmem_lane, int half, float (&acc)[64]) {
        drain_4x16(tmem_lane, 16 * half, acc);
    }
    // a lane owns one node row and 16 hidden units of it: 32 bytes (one sector) of h in, 32 bytes of h' out
    struct Pre { long long off; uint4 h0, h1; };
    __device__ static void prefetch(const Params &p, const Tile &ti, int quarter, int half, int lane, Pre &pre) {
        const int row = ti.row0 + quarter * 32 + lane;
        pre.off = row < p.num_nodes ? (long long)row * p.H + ti.jb * 32 + 16 * half : -1;   // bf16 elements
        pre.h0 = pre.h1 = make_uint4(0u, 0u, 0u, 0u);
        if (pre.off >= 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.h + pre.off);
            pre.h0 = __ldg(src);
            pre.h1 = __ldg(src + 1);
        }
    }
    // the epilogue's transpose buffers are unused by this policy (a lane stores its own 32 bytes): they hold bias4
    __device__ static void smem_init(const Params &p, float *tables) {
        float4 *b = reinterpret_cast<float4 *>(tables);
        for (int j = threadIdx.x; j < p.H; j += blockDim.x) b[j] = p.bias4[j];
    }
    __device__ static void store(const Params &p, const Tile &ti, float (&acc)[64], const Pre &pre, int half, int, float *, float *tables) {
        const float4 *bias_s = reinterpret_cast<const float4 *>(tables);
        const int j0 = ti.jb * 32 + 16 * half;
        const uint32_t hw[8] = {pre.h0.x, pre.h0.y, pre.h0.z, pre.h0.w, pre.h1.x, pre.h1.y, pre.h1.z, pre.h1.w};
        uint32_t ow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 hp = *reinterpret_cast<const __nv_bfloat162 *>(&hw[i]);
            const float hv[2] = {__low2float(hp), __high2float(hp)};
            float o[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int ii = 2 * i + u;
                const float4 b = (p.dbg & 16) ? make_float4(0.1f, 0.2f, 0.3f, 0.4f) : bias_s[j0 + ii];
                if (p.dbg & 32) { o[u] = acc[ii] + acc[16 + ii] + acc[32 + ii] + acc[48 + ii] + b.x + hv[u]; continue; }
                const float rr = sigmoid_mufu(acc[ii] + b.x);
                const float zz = sigmoid_mufu(acc[16 + ii] + b.y);
                const float nn = tanh_mufu(fmaf(rr, acc[48 + ii] + b.w, acc[32 + ii] + b.z));
                o[u] = fmaf(zz, hv[u] - nn, nn);
            }
            ow[i] = __float_as_uint(pack_bf16x2(o[0], o[1]));
        }
        if (pre.off >= 0 && !(p.dbg & 64)) {
            uint4 *dst = reinterpret_cast<uint4 *>(p.out + pre.off);
            dst[0] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            dst[1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
    }
};

// ---- policy: Mlp dense update  out = act(y W^T + b), bf16 in / out, fp32 accumulate -------------------------------
struct DensePolicyB {
    struct Params {
        CUtensorMap map_y, map_w;          // [N, D] box {64, 128}; [Hout, D] box {64, min(128, Hout)}
        const float *bias;                 // fp32 [Hout] or nullptr
        __nv_bfloat16 *out;                // [N, Hout]
        unsigned long long *trace;
        int num_nodes, D, Hout, act, n_blocks, dbg;
    };
    struct Tile { int row0, n0, b_rows; };
    __device__ static int num_tiles(const Params &p) { return ((p.num_nodes + TILE_M - 1) / TILE_M) * p.n_blocks; }
    __device__ static void tile_init(Tile &) {}
    __device__ static void tile_setup(const Params &p, int tile, Tile &ti) {
        const int rb = tile / p.n_blocks;
        ti.row0 = rb * TILE_M;
        ti.n0 = (tile - rb * p.n_blocks) * 128;
        ti.b_rows = min(128, p.Hout - ti.n0);
    }
    __device__ static int num_segments(const Params &, const Tile &) { return 1; }
    __device__ static Segment segment(const Params &p, const Tile &ti, int) {
        Segment s;
        s.a = nullptr; s.lda = 0; s.a_map = &p.map_y; s.a_row0 = ti.row0; s.K = p.D;
        s.b_map = &p.map_w; s.b_row0 = ti.n0; s.b_col0 = 0; s.b_box_rows = min(128, p.Hout);
        return s;
    }
    __device__ static int gather_row(const Params &, const Tile &, int, int) { return -1; }
    __device__ static int mma_groups(const Params &, const Tile &ti, int, MmaGroup (&g)[2]) {
        g[0] = MmaGroup{ti.b_rows, 0, 0, true};
        return 1;
    }
    __device__ static void drain(const Params &, const Tile &ti, uint32_t tmem_lane, int half, float (&acc)[64]) {
        drain_2x32(tmem_lane, 64 * half, ti.b_rows, acc);
    }
    struct Pre { long long row_off; };     // 4-byte words of the bf16 output (no global load needed)
    __device__ static void prefetch(const Params &p, const Tile &ti, int quarter, int, int lane, Pre &pre) {
        const int row = ti.row0 + quarter * 32 + lane;
        pre.row_off = row < p.num_nodes ? ((long long)row * p.Hout + ti.n0) / 2 : -1;
    }
    __device__ static void smem_init(const Params &, float *) {}
    __device__ static void store(const Params &p, const Tile &ti, float (&acc)[64], const Pre &pre, int half, int lane, float *stage, float *) {
        const int c0 = 64 * half;
        if (c0 >= ti.b_rows) return;
        float w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float v0 = acc[2 * i], v1 = acc[2 * i + 1];
            if (c0 + 2 * i < ti.b_rows) {      // pairs never straddle b_rows (Hout % 16 == 0)
                if (p.bias) { v0 += p.bias[ti.n0 + c0 + 2 * i]; v1 += p.bias[ti.n0 + c0 + 2 * i + 1]; }
                v0 = apply_act(v0, p.act); v1 = apply_act(v1, p.act);
            }
            w[i] = pack_bf16x2(v0, v1);
        }
        float *dst = reinterpret_cast<float *>(p.out) + c0 / 2;
        if (ti.b_rows - c0 >= 64) tc::warp_store_rows<32>(stage, w, dst, pre.row_off, lane);
        else if (ti.b_rows - c0 >= 32) tc::warp_store_rows<16>(stage, w, dst, pre.row_off, lane);
        else tc::warp_store_rows<8>(stage, w, dst, pre.row_off, lane);   // 16 columns
    }
};

// ---- segmented reduce over bf16 message rows (fp32 accumulation, bf16 result) ------------------------------------
// Same flat streaming walk as segment_reduce_stream_kernel (reduce.cuh); a lane owns 4 consecutive bf16 columns
// (8-byte loads), CHUNKS x 128 columns per row.
// WITH_EPI (Mlp layers): activation + LayerNorm of the aggregated row in fp32 before the single rounding to bf16
// (mlpmessagepassing.py:114-116; the same epilogue as segment_reduce_stream_kernel).
struct ReduceEpilogueB { int act; const float *ln_w, *ln_b; float ln_eps; };
template <int RED, int CHUNKS, bool WITH_EPI>
__global__ void __launch_bounds__(256)
segment_reduce_bf16_kernel(const __nv_bfloat16 *__restrict__ msg, const int32_t *__restrict__ row_ptr, int num_nodes, int D,
                           __nv_bfloat16 *__restrict__ out, const ReduceEpilogueB epi) {
    constexpr int ROWS_PER_WARP = 16;
    constexpr int UNROLL = CHUNKS == 1 ? 8 : 4;
    const int lane = threadIdx.x & 31;
    const int r0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS_PER_WARP;
    if (r0 >= num_nodes) return;
    const int nrows = min(ROWS_PER_WARP, num_nodes - r0);
    const int bound = row_ptr[r0 + min(lane, nrows)];
    const int j_begin = __shfl_sync(0xffffffffu, bound, 0), j_end = __shfl_sync(0xffffffffu, bound, nrows);
    bool col_ok[CHUNKS];
    float4 acc[CHUNKS];
    const float init = RED == PTGNN_REDUCE_MAX ? -FLT_MAX : (RED == PTGNN_REDUCE_MIN ? FLT_MAX : 0.0f);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) { col_ok[c] = (c * 32 + lane) * 4 < D; acc[c] = make_float4(init, init, init, init); }
    const size_t ld2 = (size_t)D / 4;   // row pitch in uint2 (4 bf16)
    const uint2 *msg2 = reinterpret_cast<const uint2 *>(msg);
    uint2 *out2 = reinterpret_cast<uint2 *>(out);
    int cur = 0, cur_end = __shfl_sync(0xffffffffu, bound, 1);

    auto flush = [&](int row, int count) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            float4 a = acc[c];
            if (RED == PTGNN_REDUCE_MEAN) { const float n = (float)(count < 1 ? 1 : count); a.x /= n; a.y /= n; a.z /= n; a.w /= n; }
            if (RED == PTGNN_REDUCE_MAX || RED == PTGNN_REDUCE_MIN) {
                if (a.x == init) a.x = 0.f; if (a.y == init) a.y = 0.f; if (a.z == init) a.z = 0.f; if (a.w == init) a.w = 0.f;
            }
            if (WITH_EPI) { a.x = apply_act(a.x, epi.act); a.y = apply_act(a.y, epi.act); a.z = apply_act(a.z, epi.act); a.w = apply_act(a.w, epi.act); }
            acc[c] = a;
        }
        if (WITH_EPI && epi.ln_w != nullptr) {
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
                if (col_ok[c]) s += (acc[c].x + acc[c].y) + (acc[c].z + acc[c].w);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)D;
            float q = 0.0f;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
                if (col_ok[c]) {
                    const float dx = acc[c].x - mean, dy = acc[c].y - mean, dz = acc[c].z - mean, dw = acc[c].w - mean;
                    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q / (float)D + epi.ln_eps);
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
                if (col_ok[c]) {
                    const int col = (c * 32 + lane) * 4;
                    const float4 w = *reinterpret_cast<const float4 *>(epi.ln_w + col);
                    const float4 b = *reinterpret_cast<const float4 *>(epi.ln_b + col);
                    acc[c].x = (acc[c].x - mean) * rstd * w.x + b.x; acc[c].y = (acc[c].y - mean) * rstd * w.y + b.y;
                    acc[c].z = (acc[c].z - mean) * rstd * w.z + b.z; acc[c].w = (acc[c].w - mean) * rstd * w.w + b.w;
                }
        }
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const float4 a = acc[c];
            if (col_ok[c]) {
                uint2 o;
                o.x = __float_as_uint(pack_bf16x2(a.x, a.y));
                o.y = __float_as_uint(pack_bf16x2(a.z, a.w));
                out2[(size_t)(r0 + row) * ld2 + c * 32 + lane] = o;
            }
            acc[c] = make_float4(init, init, init, init);
        }
    };
    auto comb = [&](float &a, float m) {
        if (RED == PTGNN_REDUCE_MAX) { if (m > a) a = m; }
        else if (RED == PTGNN_REDUCE_MIN) { if (m < a) a = m; }
        else a += m;
    };
    for (int j = j_begin; j < j_end; j += UNROLL) {
        uint2 m[UNROLL][CHUNKS];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (j + u < j_end) {
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (col_ok[c]) m[u][c] = __ldg(msg2 + (size_t)(j + u) * ld2 + c * 32 + lane);
            }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int jj = j + u;
            if (jj < j_end) {
                while (jj >= cur_end) {
                    const int beg = __shfl_sync(0xffffffffu, bound, cur);
                    flush(cur, cur_end - beg);
                    ++cur;
                    cur_end = __shfl_sync(0xffffffffu, bound, cur + 1);
                }
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c)
                    if (col_ok[c]) {
                        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&m[u][c].x);
                        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162 *>(&m[u][c].y);
                        comb(acc[c].x, __low2float(lo)); comb(acc[c].y, __high2float(lo));
                        comb(acc[c].z, __low2float(hi)); comb(acc[c].w, __high2float(hi));
                    }
            }
        }
    }
    for (; cur < nrows; ++cur) {
        const int beg = __shfl_sync(0xffffffffu, bound, cur), end = __shfl_sync(0xffffffffu, bound, cur + 1);
        flush(cur, end - beg);
    }
}

template <int RED>
static int launch_reduce_bf16(const __nv_bfloat16 *msg, const int32_t *row_ptr, int64_t N, int D, __nv_bfloat16 *out,
                              const ReduceEpilogueB *epi, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(N, 8 * 16);
    const ReduceEpilogueB e = epi ? *epi : ReduceEpilogueB{PTGNN_ACT_NONE, nullptr, nullptr, 0.0f};
    {
        TimedScope timed__(PTGNN_KERNEL_REDUCE, st);
        if (epi) {
            if (D <= 128) segment_reduce_bf16_kernel<RED, 1, true><<<grid, 256, 0, st>>>(msg, row_ptr, (int)N, D, out, e);
            else segment_reduce_bf16_kernel<RED, 2, true><<<grid, 256, 0, st>>>(msg, row_ptr, (int)N, D, out, e);
        } else {
            if (D <= 128) segment_reduce_bf16_kernel<RED, 1, false><<<grid, 256, 0, st>>>(msg, row_ptr, (int)N, D, out, e);
            else segment_reduce_bf16_kernel<RED, 2, false><<<grid, 256, 0, st>>>(msg, row_ptr, (int)N, D, out, e);
        }
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}
static int reduce_bf16(int reduce, const __nv_bfloat16 *msg, const int32_t *row_ptr, int64_t N, int D, __nv_bfloat16 *out,
                       const ReduceEpilogueB *epi, cudaStream_t st) {
    switch (reduce) {
        case PTGNN_REDUCE_SUM: return launch_reduce_bf16<PTGNN_REDUCE_SUM>(msg, row_ptr, N, D, out, epi, st);
        case PTGNN_REDUCE_MEAN: return launch_reduce_bf16<PTGNN_REDUCE_MEAN>(msg, row_ptr, N, D, out, epi, st);
        case PTGNN_REDUCE_MAX: return launch_reduce_bf16<PTGNN_REDUCE_MAX>(msg, row_ptr, N, D, out, epi, st);
        default: return launch_reduce_bf16<PTGNN_REDUCE_MIN>(msg, row_ptr, N, D, out, epi, st);
    }
}

static int debug_bits() {
    const char *e = getenv("PTGNN_TC_DEBUG");
    return e ? atoi(e) : 0;
}
static int sm_count() {   // of the CURRENT device: a process may drive several GPUs (nothing cached across devices)
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}
template <class Policy>
static int launch_pipeline(const typename Policy::Params &p, int total_tiles, int category, cudaStream_t st) {
    if (total_tiles <= 0) return PTGNN_OK;
    // per launch, not once per process: the attribute is per device (and per context)
    PTGNN_CUDA(cudaFuncSetAttribute(tc_pipeline_bf16_kernel<Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int sms = sm_count();
    const int grid = total_tiles < sms ? total_tiles : sms;
    {
        TimedScope timed__(category, st);
        tc_pipeline_bf16_kernel<Policy><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

struct WsB { size_t msg, agg, w, p1, p2, bias, gws, total; };
static WsB ws_layout(int64_t N, int64_t E, int T, int H, int D) {
    WsB w{};
    size_t o = 0;
    auto add = [&](size_t cnt) { size_t at = o; o += ws_slice(cnt, 2); return at; };
    w.msg = add((size_t)E * D + 8);
    w.agg = add((size_t)N * D + 8);
    w.w = add((size_t)T * D * H + 8);
    w.p1 = add((size_t)(H / 32 + 1) * 128 * D);
    w.p2 = add((size_t)(H / 32 + 1) * 128 * H);
    w.bias = add((size_t)H * 8 + 8);
    w.gws = add(gruws::supported(1, H, D) ? gruws::pack_bytes(1, H, D) / 2 + 8 : 8);   // weights-stationary GRU packing (fused path)
    w.total = o;
    return w;
}

}  // namespace tcb
}  // namespace ptgnn

using namespace ptgnn;
using namespace ptgnn::tcb;

extern "C" size_t ptgnn_b200_gated_workspace_bytes_bf16(int64_t num_nodes, int64_t num_edges, int32_t num_types,
                                                        int32_t state_dim, int32_t message_dim) {
    if (num_nodes < 0 || num_edges < 0 || num_types < 0 || state_dim <= 0 || message_dim <= 0) return 0;
    return ws_layout(num_nodes, num_edges, num_types, state_dim, message_dim).total;
}

// weight cache of the bf16 path: the tail of the workspace layout [bf16 edge weights | P1 | P2 | bias4]
static size_t gated_cache_bytes_bf16(int T, int H, int D) {
    const WsB L = ws_layout(0, 0, T, H, D);
    return L.total - L.w;
}

static int gated_forward_bf16_impl(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes,
                                   int32_t state_dim, int32_t message_dim, int32_t num_types, const int64_t *type_off,
                                   const int32_t *row_ptr, const int32_t *pos, const int32_t *src32,
                                   const float *const *edge_weights, const float *gru_w_ih, const float *gru_w_hh,
                                   const float *gru_b_ih, const float *gru_b_hh, int32_t reduce, uint16_t *out_states,
                                   void *workspace, size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                   int32_t cache_valid, void *stream, const ptgnn_b200_block_plan *bp = nullptr) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = state_dim, D = message_dim;
    PTGNN_CHECK_ARG(num_types >= 0 && num_types <= PTGNN_MAX_EDGE_TYPES && (type_off || bp), "gated_forward_bf16: bad num_types=%d", num_types);
    const bool fused_path = bp != nullptr;
    if (fused_path && !fused::supported(1, H, D, 0)) {
        set_error("gated_forward_fused (bf16): dims H=%d D=%d are not supported by the fused kernel", H, D);
        return PTGNN_E_UNSUPPORTED;
    }
    const int64_t E = fused_path ? 0 : type_off[num_types];
    PTGNN_CHECK_ARG(num_nodes >= 0 && num_nodes < INT32_MAX && E >= 0 && E < INT32_MAX, "gated_forward_bf16: sizes out of range");
    if (H % 32 != 0 || D % 16 != 0 || H < 64 || D < 64 || D > 256 || H > 1024) {
        set_error("gated_forward_bf16: needs state dim %% 32 == 0 (>= 64) and message dim %% 16 == 0 in [64, 256]; got %d, %d", H, D);
        return PTGNN_E_UNSUPPORTED;
    }
    PTGNN_CHECK_ARG(reduce >= PTGNN_REDUCE_SUM && reduce <= PTGNN_REDUCE_MIN, "gated_forward_bf16: bad reduce %d", reduce);
    if (num_nodes == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(node_states && out_states && row_ptr && gru_w_ih && gru_w_hh && gru_b_ih && gru_b_hh, "gated_forward_bf16: null pointer");
    PTGNN_CHECK_ARG(E == 0 || (pos && src32 && edge_weights), "gated_forward_bf16: null edge arrays");
    PTGNN_CHECK_ARG(!fused_path || (bp->group_off && edge_weights && num_types > 0), "gated_forward_fused (bf16): null block plan arrays");
    const WsB L = ws_layout(num_nodes, E, num_types, H, D);
    if (workspace_bytes < L.total || !workspace) {
        set_error("gated_forward_bf16: workspace %zu < required %zu", workspace_bytes, L.total);
        return PTGNN_E_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    auto b16 = [&](size_t off) { return reinterpret_cast<__nv_bfloat16 *>(ws + off); };
    const __nv_bfloat16 *h = reinterpret_cast<const __nv_bfloat16 *>(node_states);
    const __nv_bfloat16 *hsrc = gather_states ? reinterpret_cast<const __nv_bfloat16 *>(gather_states) : h;
    __nv_bfloat16 *msg = b16(L.msg), *agg = b16(L.agg);
    // derived weights live in the workspace (re-derived every call) or in the caller's cache (derived when !cache_valid)
    char *wbase = ws + L.w;
    bool pack = true;
    if (weight_cache != nullptr) {
        const size_t need = gated_cache_bytes_bf16(num_types, H, D);
        if (weight_cache_bytes < need) {
            set_error("gated_forward_bf16: weight cache %zu < required %zu", weight_cache_bytes, need);
            return PTGNN_E_WORKSPACE;
        }
        wbase = static_cast<char *>(weight_cache);
        pack = !cache_valid;
    }
    __nv_bfloat16 *wb = reinterpret_cast<__nv_bfloat16 *>(wbase);
    __nv_bfloat16 *p1 = reinterpret_cast<__nv_bfloat16 *>(wbase + (L.p1 - L.w)), *p2 = reinterpret_cast<__nv_bfloat16 *>(wbase + (L.p2 - L.w));
    float4 *bias4 = reinterpret_cast<float4 *>(wbase + (L.bias - L.w));

    // 0. weights -> bf16 (edge weights [T][D][H]; GRU gate blocks)
    if (pack) {
        if (fused_path) {       // same bytes as the plain bf16 copy, in the fused kernel's TMEM-lane layout
            const int prc = fused::pack_weights(1, num_types, H, 0, edge_weights, wb, bp->status, st);
            if (prc) return prc;
        } else {
            ConvSrc cs{};
            cs.num = num_types; cs.elems = D * H;
            for (int t = 0; t < num_types; ++t) cs.w[t] = edge_weights[t];
            {
                TimedScope timed__(PTGNN_KERNEL_PACK, st);
                convert_weights_kernel<<<148, 256, 0, st>>>(cs, wb);
            }
            PTGNN_LAUNCHED();
        }
        {
            TimedScope timed__(PTGNN_KERNEL_PACK, st);
            pack_gru_bf16_kernel<<<148, 256, 0, st>>>(gru_w_ih, gru_w_hh, H, D, p1, p2);
        }
        PTGNN_LAUNCHED();
        {
            TimedScope timed__(PTGNN_KERNEL_PACK, st);
            pack_gru_bias_bf16_kernel<<<(H + 127) / 128, 128, 0, st>>>(gru_b_ih, gru_b_hh, H, bias4);
        }
        PTGNN_LAUNCHED();
    }

    int rc = PTGNN_OK;
    if (fused_path) {
        // 1+2. gather -> W_t -> segmented reduce in one kernel (no message buffer)
        fused::AggregateArgs a{};
        a.nprod = 1; a.src_rows = hsrc; a.tgt_rows = nullptr; a.num_nodes = num_nodes; a.K = H; a.num_types = num_types;
        a.use_target = 0; a.reduce = reduce; a.block_targets = bp->block_targets; a.group_off = bp->group_off; a.src_f = bp->src_f;
        a.tl_f = bp->tl_f; a.row_ptr = row_ptr; a.packed_weights = wb; a.epi = fused::Epilogue{PTGNN_ACT_NONE, nullptr, nullptr, 0.0f};
        a.out = agg; a.out_mode = 1; a.status = bp->status;
        rc = fused::aggregate(a, st);
        if (rc) return rc;
        static const bool ws_gru_on = [] { const char *e = getenv("PTGNN_B200_GRU"); return !(e && e[0] == 't'); }();
        if (ws_gru_on && gruws::supported(1, H, D)) {
            // 3. GRUCell, weights-stationary (the gate weights stay in shared memory, only node rows stream)
            char *gws = wbase + (L.gws - L.w);
            if (pack) {
                rc = gruws::pack(1, H, D, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, gws, st);
                if (rc) return rc;
            }
            return gruws::update(1, agg, h, h, num_nodes, H, D, gws, out_states, nullptr, nullptr, st);
        }
    } else {
    // 1. messages
    MsgPolicyB::Params mp{};
    rc = make_map_bf16(&mp.map_w, wb, (uint64_t)num_types * D, H, D < 128 ? D : 128);
    if (rc) return rc;
    mp.h = hsrc; mp.h_tgt = h; mp.src32 = src32; mp.tgt32 = nullptr; mp.use_target = 0; mp.pos = pos; mp.msg = msg; mp.H = H; mp.D = D;
    mp.num_types = num_types;
    mp.n_blocks = (D + 127) / 128;
    mp.dbg = debug_bits(); mp.trace = tc::trace_buffer(PTGNN_KERNEL_MESSAGE + 10);
    int tiles = 0;
    for (int t = 0; t < num_types; ++t) {
        mp.edge_off[t] = (int32_t)type_off[t];
        mp.tile_off[t] = tiles;
        tiles += (int)ceil_div(type_off[t + 1] - type_off[t], TILE_M);
    }
    for (int t = num_types; t <= PTGNN_MAX_EDGE_TYPES; ++t) { mp.edge_off[t] = (int32_t)type_off[num_types]; mp.tile_off[t] = tiles; }
    rc = launch_pipeline<MsgPolicyB>(mp, tiles * mp.n_blocks, PTGNN_KERNEL_MESSAGE, st);
    if (rc) return rc;

    // 2. segmented reduce (fp32 accumulate, bf16 result)
    rc = reduce_bf16(reduce, msg, row_ptr, num_nodes, D, agg, nullptr, st);
    if (rc) return rc;
    }

    // 3. GRUCell
    GruPolicyB::Params gp{};
    const uint64_t prow = (uint64_t)(H / 32) * 128;
    rc = make_map_bf16(&gp.map_agg, agg, num_nodes, D, 128);
    if (!rc) rc = make_map_bf16(&gp.map_h, h, num_nodes, H, 128);
    if (!rc) rc = make_map_bf16(&gp.map_p1, p1, prow, D, 128);
    if (!rc) rc = make_map_bf16(&gp.map_p2, p2, prow, H, 128);
    if (rc) return rc;
    gp.h = h; gp.bias4 = bias4; gp.out = reinterpret_cast<__nv_bfloat16 *>(out_states);
    gp.num_nodes = (int)num_nodes; gp.H = H; gp.D = D; gp.n_jb = H / 32; gp.dbg = debug_bits(); gp.trace = tc::trace_buffer(PTGNN_KERNEL_GRU + 10);
    return launch_pipeline<GruPolicyB>(gp, (int)ceil_div(num_nodes, TILE_M) * gp.n_jb, PTGNN_KERNEL_GRU, st);
}

extern "C" int ptgnn_b200_gated_forward_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes,
                                             int32_t state_dim, int32_t message_dim, int32_t num_types,
                                             const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                             const int32_t *src32, const float *const *edge_weights, const float *gru_w_ih,
                                             const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                                             int32_t reduce, uint16_t *out_states, void *workspace, size_t workspace_bytes,
                                             void *stream) {
    return gated_forward_bf16_impl(node_states, gather_states, num_nodes, state_dim, message_dim, num_types, type_off, row_ptr,
                                   pos, src32, edge_weights, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, reduce, out_states,
                                   workspace, workspace_bytes, nullptr, 0, 0, stream);
}

extern "C" size_t ptgnn_b200_gated_weight_cache_bytes_bf16(int32_t num_types, int32_t state_dim, int32_t message_dim) {
    if (num_types < 0 || num_types > PTGNN_MAX_EDGE_TYPES || state_dim <= 0 || message_dim <= 0) return 0;
    return gated_cache_bytes_bf16(num_types, state_dim, message_dim);
}

extern "C" int ptgnn_b200_gated_forward_cached_bf16(const uint16_t *node_states, const uint16_t *gather_states,
                                                    int64_t num_nodes, int32_t state_dim, int32_t message_dim,
                                                    int32_t num_types, const int64_t *type_off, const int32_t *row_ptr,
                                                    const int32_t *pos, const int32_t *src32,
                                                    const float *const *edge_weights, const float *gru_w_ih,
                                                    const float *gru_w_hh, const float *gru_b_ih, const float *gru_b_hh,
                                                    int32_t reduce, uint16_t *out_states, void *workspace,
                                                    size_t workspace_bytes, void *weight_cache, size_t weight_cache_bytes,
                                                    int32_t cache_valid, void *stream) {
    return gated_forward_bf16_impl(node_states, gather_states, num_nodes, state_dim, message_dim, num_types, type_off, row_ptr,
                                   pos, src32, edge_weights, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, reduce, out_states,
                                   workspace, workspace_bytes, weight_cache, weight_cache_bytes, cache_valid, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// MlpMessagePassingLayer with bf16 states (mlpmessagepassing.py:68-117 under torch.autocast(bfloat16)): bf16 messages,
// fp32 aggregation + GELU + LayerNorm, bf16 dense update with fp32 accumulation.  Parameters arrive in fp32.
// ---------------------------------------------------------------------------------------------------------------
namespace ptgnn {
namespace tcb {
struct MlpWsB { size_t msg, y, w, wd, total; };
static MlpWsB mlp_ws_layout(int64_t N, int64_t E, int T, int H, int D, int Hout, int use_target) {
    MlpWsB w{};
    size_t o = 0;
    auto add = [&](size_t cnt) { size_t at = o; o += ws_slice(cnt, 2); return at; };
    w.msg = add((size_t)E * D + 8);
    w.y = add((size_t)N * D + 8);
    w.w = add((size_t)T * D * H * (use_target ? 2 : 1) + 8);
    w.wd = add((size_t)Hout * D + 8);
    w.total = o;
    return w;
}
}  // namespace tcb
}  // namespace ptgnn

extern "C" size_t ptgnn_b200_mlp_workspace_bytes_bf16(int64_t num_nodes, int64_t num_edges, int32_t num_types, int32_t in_dim,
                                                      int32_t message_dim, int32_t out_dim, int32_t use_target_state) {
    if (num_nodes < 0 || num_edges < 0 || num_types < 0 || in_dim <= 0 || message_dim <= 0 || out_dim <= 0) return 0;
    return mlp_ws_layout(num_nodes, num_edges, num_types, in_dim, message_dim, out_dim, use_target_state).total;
}

static int mlp_forward_bf16_impl(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes,
                                           int32_t in_dim, int32_t message_dim, int32_t out_dim, int32_t num_types,
                                           const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                           const int32_t *src32, const int32_t *tgt32, const float *const *edge_weights,
                                           int32_t use_target_state, int32_t reduce, int32_t message_activation,
                                           const float *ln_weight, const float *ln_bias, float ln_eps,
                                           const float *dense_weight, const float *dense_bias, int32_t dense_activation,
                                           uint16_t *out_states, void *workspace, size_t workspace_bytes, void *stream,
                                           const ptgnn_b200_block_plan *bp) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int H = in_dim, D = message_dim, ut = use_target_state ? 1 : 0, Kw = H * (1 + ut);
    PTGNN_CHECK_ARG(num_types >= 0 && num_types <= PTGNN_MAX_EDGE_TYPES && (type_off || bp), "mlp_forward_bf16: bad num_types=%d", num_types);
    const bool fused_path = bp != nullptr;
    if (fused_path && !fused::supported(1, H, D, ut)) {
        set_error("mlp_forward_fused (bf16): dims H=%d D=%d are not supported by the fused kernel", H, D);
        return PTGNN_E_UNSUPPORTED;
    }
    const int64_t E = fused_path ? 0 : type_off[num_types];
    PTGNN_CHECK_ARG(num_nodes >= 0 && num_nodes < INT32_MAX && E >= 0 && E < INT32_MAX, "mlp_forward_bf16: sizes out of range");
    PTGNN_CHECK_ARG(dense_weight ? out_dim > 0 : out_dim == D, "mlp_forward_bf16: out_dim=%d inconsistent", out_dim);
    if (H % 32 != 0 || D % 16 != 0 || H < 64 || D < 64 || D > 256 || H > 1024 || (dense_weight && (out_dim % 16 != 0 || out_dim < 64))) {
        set_error("mlp_forward_bf16: needs state dim %% 32 == 0 (>= 64), message dim %% 16 == 0 in [64, 256], output dim %% 16 == 0 (>= 64); "
                  "got %d, %d, %d", H, D, out_dim);
        return PTGNN_E_UNSUPPORTED;
    }
    PTGNN_CHECK_ARG(reduce >= PTGNN_REDUCE_SUM && reduce <= PTGNN_REDUCE_MIN, "mlp_forward_bf16: bad reduce %d", reduce);
    PTGNN_CHECK_ARG(message_activation >= PTGNN_ACT_NONE && message_activation <= PTGNN_ACT_RELU &&
                        dense_activation >= PTGNN_ACT_NONE && dense_activation <= PTGNN_ACT_RELU, "mlp_forward_bf16: bad activation");
    PTGNN_CHECK_ARG((ln_weight == nullptr) == (ln_bias == nullptr), "mlp_forward_bf16: ln_weight/ln_bias must both be set");
    if (num_nodes == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(node_states && out_states && row_ptr, "mlp_forward_bf16: null pointer");
    PTGNN_CHECK_ARG(E == 0 || (pos && src32 && edge_weights && (!ut || tgt32)), "mlp_forward_bf16: null edge arrays");
    const MlpWsB L = mlp_ws_layout(num_nodes, E, num_types, H, D, out_dim, ut);
    if (workspace_bytes < L.total || !workspace) {
        set_error("mlp_forward_bf16: workspace %zu < required %zu", workspace_bytes, L.total);
        return PTGNN_E_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    auto b16 = [&](size_t off) { return reinterpret_cast<__nv_bfloat16 *>(ws + off); };
    const __nv_bfloat16 *h = reinterpret_cast<const __nv_bfloat16 *>(node_states);
    const __nv_bfloat16 *hsrc = gather_states ? reinterpret_cast<const __nv_bfloat16 *>(gather_states) : h;
    __nv_bfloat16 *out = reinterpret_cast<__nv_bfloat16 *>(out_states);
    __nv_bfloat16 *msg = b16(L.msg), *wb = b16(L.w), *wd = b16(L.wd);
    __nv_bfloat16 *y = dense_weight ? b16(L.y) : out;

    // 0. weights -> bf16
    int rc = PTGNN_OK;
    if (fused_path) {
        PTGNN_CHECK_ARG(bp->group_off && edge_weights && num_types > 0, "mlp_forward_fused (bf16): null block plan arrays");
        rc = fused::pack_weights(1, num_types, H, ut, edge_weights, wb, bp->status, st);
        if (rc) return rc;
    } else if (num_types > 0) {
        ConvSrc cs{};
        cs.num = num_types; cs.elems = D * Kw;
        for (int t = 0; t < num_types; ++t) cs.w[t] = edge_weights[t];
        {
            TimedScope timed__(PTGNN_KERNEL_PACK, st);
            convert_weights_kernel<<<148, 256, 0, st>>>(cs, wb);
        }
        PTGNN_LAUNCHED();
    }
    if (dense_weight) {
        ConvSrc cs{};
        cs.num = 1; cs.elems = out_dim * D; cs.w[0] = dense_weight;
        {
            TimedScope timed__(PTGNN_KERNEL_PACK, st);
            convert_weights_kernel<<<148, 256, 0, st>>>(cs, wd);
        }
        PTGNN_LAUNCHED();
    }

    if (fused_path) {
        // 1+2. gather -> W_t -> segmented reduce (+ activation + LayerNorm at write-out) in one kernel
        fused::AggregateArgs a{};
        a.nprod = 1; a.src_rows = hsrc; a.tgt_rows = h; a.num_nodes = num_nodes; a.K = H; a.num_types = num_types;
        a.use_target = ut; a.reduce = reduce; a.block_targets = bp->block_targets; a.group_off = bp->group_off; a.src_f = bp->src_f;
        a.tl_f = bp->tl_f; a.row_ptr = row_ptr; a.packed_weights = wb;
        a.epi = fused::Epilogue{message_activation, ln_weight, ln_bias, ln_eps};
        a.out = y; a.out_mode = 1; a.status = bp->status;
        rc = fused::aggregate(a, st);
        if (rc || !dense_weight) return rc;
    } else {
    // 1. messages  m_e = W_t [h_src ; h_tgt]
    if (E > 0) {
        MsgPolicyB::Params mp{};
        rc = make_map_bf16(&mp.map_w, wb, (uint64_t)num_types * D, Kw, D < 128 ? D : 128);
        if (rc) return rc;
        mp.h = hsrc; mp.h_tgt = h; mp.src32 = src32; mp.tgt32 = tgt32; mp.use_target = ut; mp.pos = pos; mp.msg = msg; mp.H = H; mp.D = D;
        mp.num_types = num_types;
        mp.n_blocks = (D + 127) / 128;
        mp.dbg = debug_bits(); mp.trace = tc::trace_buffer(PTGNN_KERNEL_MESSAGE + 10);
        int tiles = 0;
        for (int t = 0; t < num_types; ++t) {
            mp.edge_off[t] = (int32_t)type_off[t];
            mp.tile_off[t] = tiles;
            tiles += (int)ceil_div(type_off[t + 1] - type_off[t], TILE_M);
        }
        for (int t = num_types; t <= PTGNN_MAX_EDGE_TYPES; ++t) { mp.edge_off[t] = (int32_t)type_off[num_types]; mp.tile_off[t] = tiles; }
        rc = launch_pipeline<MsgPolicyB>(mp, tiles * mp.n_blocks, PTGNN_KERNEL_MESSAGE, st);
        if (rc) return rc;
    }

    // 2. aggregate + activation + LayerNorm (fp32) -> y (bf16)
    const ReduceEpilogueB epi{message_activation, ln_weight, ln_bias, ln_eps};
    rc = reduce_bf16(reduce, msg, row_ptr, num_nodes, D, y, &epi, st);
    if (rc || !dense_weight) return rc;
    }

    // 3. dense update
    DensePolicyB::Params dp{};
    rc = make_map_bf16(&dp.map_y, y, num_nodes, D, 128);
    if (!rc) rc = make_map_bf16(&dp.map_w, wd, out_dim, D, out_dim < 128 ? out_dim : 128);
    if (rc) return rc;
    dp.bias = dense_bias; dp.out = out; dp.num_nodes = (int)num_nodes; dp.D = D; dp.Hout = out_dim; dp.act = dense_activation;
    dp.n_blocks = (out_dim + 127) / 128; dp.dbg = debug_bits(); dp.trace = tc::trace_buffer(PTGNN_KERNEL_DENSE + 10);
    return launch_pipeline<DensePolicyB>(dp, (int)ceil_div(num_nodes, TILE_M) * dp.n_blocks, PTGNN_KERNEL_DENSE, st);
}

extern "C" int ptgnn_b200_mlp_forward_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes,
                                           int32_t in_dim, int32_t message_dim, int32_t out_dim, int32_t num_types,
                                           const int64_t *type_off, const int32_t *row_ptr, const int32_t *pos,
                                           const int32_t *src32, const int32_t *tgt32, const float *const *edge_weights,
                                           int32_t use_target_state, int32_t reduce, int32_t message_activation,
                                           const float *ln_weight, const float *ln_bias, float ln_eps,
                                           const float *dense_weight, const float *dense_bias, int32_t dense_activation,
                                           uint16_t *out_states, void *workspace, size_t workspace_bytes, void *stream) {
    return mlp_forward_bf16_impl(node_states, gather_states, num_nodes, in_dim, message_dim, out_dim, num_types, type_off, row_ptr,
                                 pos, src32, tgt32, edge_weights, use_target_state, reduce, message_activation, ln_weight, ln_bias,
                                 ln_eps, dense_weight, dense_bias, dense_activation, out_states, workspace, workspace_bytes, stream,
                                 nullptr);
}

// ---- fused variants (called from the dtype-dispatching entry points in layers.cu) ---------------------------------------------
namespace ptgnn {
namespace tcb {
size_t gated_fused_workspace_bytes_bf16(int64_t N, int T, int H, int D) { return ws_layout(N, 0, T, H, D).total; }
size_t gated_fused_cache_bytes_bf16(int T, int H, int D) { return gated_cache_bytes_bf16(T, H, D); }
int gated_forward_fused_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes, int32_t state_dim,
                             int32_t message_dim, int32_t num_types, const ptgnn_b200_block_plan *bp, const int32_t *row_ptr,
                             const float *const *edge_weights, const float *gru_w_ih, const float *gru_w_hh, const float *gru_b_ih,
                             const float *gru_b_hh, int32_t reduce, uint16_t *out_states, void *workspace, size_t workspace_bytes,
                             void *weight_cache, size_t weight_cache_bytes, int32_t cache_valid, void *stream) {
    return gated_forward_bf16_impl(node_states, gather_states, num_nodes, state_dim, message_dim, num_types, nullptr, row_ptr,
                                   nullptr, nullptr, edge_weights, gru_w_ih, gru_w_hh, gru_b_ih, gru_b_hh, reduce, out_states,
                                   workspace, workspace_bytes, weight_cache, weight_cache_bytes, cache_valid, stream, bp);
}
size_t mlp_fused_workspace_bytes_bf16(int64_t N, int T, int H, int D, int Hout, int use_target) {
    return mlp_ws_layout(N, 0, T, H, D, Hout > 0 ? Hout : D, use_target).total;
}
int mlp_forward_fused_bf16(const uint16_t *node_states, const uint16_t *gather_states, int64_t num_nodes, int32_t in_dim,
                           int32_t message_dim, int32_t out_dim, int32_t num_types, const ptgnn_b200_block_plan *bp,
                           const int32_t *row_ptr, const float *const *edge_weights, int32_t use_target_state, int32_t reduce,
                           int32_t message_activation, const float *ln_weight, const float *ln_bias, float ln_eps,
                           const float *dense_weight, const float *dense_bias, int32_t dense_activation, uint16_t *out_states,
                           void *workspace, size_t workspace_bytes, void *stream) {
    return mlp_forward_bf16_impl(node_states, gather_states, num_nodes, in_dim, message_dim, out_dim, num_types, nullptr, row_ptr,
                                 nullptr, nullptr, nullptr, edge_weights, use_target_state, reduce, message_activation, ln_weight,
                                 ln_bias, ln_eps, dense_weight, dense_bias, dense_activation, out_states, workspace, workspace_bytes,
                                 stream, bp);
}
}  // namespace tcb
}  // namespace ptgnn
