// Persistent, warp-specialised tcgen05 pipeline shared by the three GEMM-bearing kernels of the hot path
// (per-edge messages, GRUCell update, Mlp dense update).  fp32-exact via 3xTF32 (see tc_common.cuh).
//
//   warps 0-3  PRODUCERS  stage the raw fp32 A tile (128 rows x 32 k) in shared memory -- gathered node-state rows via
//                         cp.async (per-edge messages) or one TMA tile (contiguous node rows) -- and the pre-split
//                         weight tile (hi, lo) via TMA.  Each producer thread then owns ONE row: it reads its 32
//                         floats, splits them into TF32 hi/lo and writes both into TENSOR MEMORY (tcgen05.st), so
//                         the MMAs take A from TMEM ("TS" form) and shared memory only feeds B.
//   warp  4    MMA        one thread: wait full[slot]; per K=8 step issue hi*hi into the MAIN accumulator and
//                         hi*lo + lo*hi into the CORRECTION accumulator (tensor-core accumulation truncates, so the
//                         small terms must not perturb the main sum); tcgen05.commit -> empty[slot]; per tile
//                         commit -> tmem_full.
//   warps 5-12 EPILOGUE   wait tmem_full; drain main + correction into registers (row per thread, two warps per TMEM
//                         lane quarter, half of the columns each); release the accumulator (tmem_empty) BEFORE the
//                         policy's store phase (smem-transposed coalesced rows), so the next tile's MMAs overlap it.
//
// Why TS: an SS-mode tf32 MMA (128x128x8) reads 8 KB of operands from shared memory = the 64 cycles its math takes,
// and the producers' traffic then starves it (measured: tensor pipe 17 % active).  With A in TMEM an MMA reads 4 KB.
//
// One CTA per SM (grid = #SMs), static round-robin over tiles.
//   shared memory: 4 slots x 48 KB (raw A | B_hi | B_lo, 128-byte SWIZZLE_128B rows) + 16 KB epilogue transpose
//   tensor memory (512 columns): [0,128) main acc | [128,256) correction acc | [256,512) 4 x (A_hi 32 | A_lo 32)
// Every mbarrier wait is bounded (tc_common.cuh): a protocol bug traps instead of hanging the GPU.
//
// A Policy supplies:
//   struct Params;   struct Tile;
//   __device__ static int  num_tiles(const Params&);
//   __device__ static void tile_setup(const Params&, int tile, Tile&);
//   __device__ static int  num_segments(const Params&, const Tile&);
//   __device__ static Segment segment(const Params&, const Tile&, int seg);
//   __device__ static int  gather_row(const Params&, const Tile&, int seg, int r);   only when segment.a_map == nullptr
//   __device__ static int  mma_groups(const Params&, const Tile&, int seg, MmaGroup (&g)[2]);
//   __device__ static void drain(const Params&, const Tile&, uint32_t tmem_lane, int half, float (&acc)[64]);
//                          (read this warp's share of main + correction accumulators; tmem_ld_sum16/32 below)
//   __device__ static void store(const Params&, const Tile&, float (&acc)[64], int quarter, int half, int lane, float* stage);
#pragma once
#include <cuda.h>

#include "tc_common.cuh"

namespace ptgnn {
namespace tc {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 32;                       // fp32 per k-chunk = one 128-byte swizzled row
constexpr int NUM_SLOTS = 4;
constexpr int LOOKAHEAD = 2;                      // chunks of loads in flight ahead of the chunk being converted; with 4 slots the
                                                  // producer then waits on the MMA of chunk c-2, not c-1: one iteration of slack
                                                  // takes the arrive->commit->wake handshake (~0.5 us) off the critical path
constexpr int PRODUCER_THREADS = 128;
constexpr int MMA_WARP = 4;
constexpr int NUM_EPI_WARPS = 8;                  // two per TMEM lane quarter, each draining half of the columns
constexpr int NUM_THREADS = (5 + NUM_EPI_WARPS) * 32;
constexpr int OPERAND_BYTES = TILE_M * CHUNK_K * 4;   // 16 KB: one 128 x 32 fp32 operand tile
constexpr int SLOT_BYTES = 3 * OPERAND_BYTES;         // raw A | B_hi | B_lo
constexpr int RING_BYTES = NUM_SLOTS * SLOT_BYTES;
constexpr int STAGE_BYTES_PER_WARP = 32 * 32 * 4;     // epilogue transpose buffer: 32 rows x 32 fp32
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/ + NUM_EPI_WARPS * STAGE_BYTES_PER_WARP;
constexpr int CORR_OFF = 128;                         // correction accumulator columns
constexpr int A_TMEM_OFF = 256;                       // A operand ring: slot s at columns A_TMEM_OFF + 64 s (hi | lo)

struct Segment {        // one K-range of the tile's GEMM
    const float *a;     // gathered A rows (row pitch lda) -- used when a_map == nullptr
    int lda;
    const CUtensorMap *a_map;    // contiguous A rows: TMA box {32 cols, 128 rows} at (k, a_row0)
    int a_row0;
    const CUtensorMap *b_hi_map, *b_lo_map;   // TMA box {32 cols, b_box_rows} at (b_col0 + k, b_row0)
    int b_row0, b_col0, b_box_rows;
    int K;              // columns of this segment (multiple of 4)
};
struct MmaGroup {       // per K-step: B rows [row_off, row_off + n) -> accumulator columns [col_off, col_off + n)
    int n, row_off, col_off;
    bool fresh;         // true: the first K-step of this segment overwrites the accumulator columns
};

__device__ __forceinline__ uint32_t swz(int row, int q) { return (uint32_t)(row * 128 + ((q ^ (row & 7)) << 4)); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// One lane polls the mbarrier, the warp re-converges on __syncwarp (32x fewer try_wait instructions on the XU pipe).
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// ---- epilogue transposes through shared memory ---------------------------------------------------------------
// The accumulator comes out of TMEM one ROW per thread; storing that directly makes every warp-wide store touch 32
// different 128-byte lines.  Staging 32 rows x NCOLS through a (chunk-XOR-swizzled) buffer lets each store
// instruction write whole rows: 4 (NCOLS = 32) or 8 (NCOLS = 16) lines per instruction instead of 32.
// `row_off` is this lane's destination element offset from `dst_base` (negative = row not stored).
template <int NCOLS>
__device__ __forceinline__ void warp_store_rows(float *stage, const float *v, float *dst_base, long long row_off, int lane) {
    constexpr int CPR = NCOLS / 4;   // 16-byte chunks per row
#pragma unroll
    for (int j = 0; j < CPR; ++j)
        *reinterpret_cast<float4 *>(stage + (lane * CPR + (j ^ (lane & (CPR - 1)))) * 4) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
        const int idx = it * 32 + lane, row = idx / CPR, ch = idx % CPR;
        const float4 val = *reinterpret_cast<const float4 *>(stage + (row * CPR + (ch ^ (row & (CPR - 1)))) * 4);
        const long long off = __shfl_sync(0xffffffffu, row_off, row);
        if (off >= 0) *reinterpret_cast<float4 *>(dst_base + off + ch * 4) = val;
    }
    __syncwarp();
}
// The mirror image for reads: every lane ends up with NCOLS consecutive floats of ITS row.
template <int NCOLS>
__device__ __forceinline__ void warp_load_rows(float *stage, float *v, const float *src_base, long long row_off, int lane) {
    constexpr int CPR = NCOLS / 4;
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
        const int idx = it * 32 + lane, row = idx / CPR, ch = idx % CPR;
        const long long off = __shfl_sync(0xffffffffu, row_off, row);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (off >= 0) val = *reinterpret_cast<const float4 *>(src_base + off + ch * 4);
        *reinterpret_cast<float4 *>(stage + (row * CPR + (ch ^ (row & (CPR - 1)))) * 4) = val;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CPR; ++j) {
        const float4 val = *reinterpret_cast<const float4 *>(stage + (lane * CPR + (j ^ (lane & (CPR - 1)))) * 4);
        v[4 * j] = val.x; v[4 * j + 1] = val.y; v[4 * j + 2] = val.z; v[4 * j + 3] = val.w;
    }
    __syncwarp();
}

// accumulator value = main + correction
__device__ __forceinline__ void tmem_ld_sum32(uint32_t taddr, float *v) {
    float m[32], c[32];
    tmem_ld_32cols(taddr, m);
    tmem_ld_32cols(taddr + CORR_OFF, c);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = m[i] + c[i];
}
__device__ __forceinline__ void tmem_ld_sum16(uint32_t taddr, float *v) {
    float m[16], c[16];
    tmem_ld_16cols(taddr, m);
    tmem_ld_16cols(taddr + CORR_OFF, c);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = m[i] + c[i];
}

template <class Policy>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_pipeline_kernel(const __grid_constant__ typename Policy::Params p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned ring (SWIZZLE_128B descriptors / TMA swizzle assume base_offset = 0)
    unsigned char *ring = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + RING_BYTES);
    uint64_t *full = bars, *empty = bars + NUM_SLOTS, *landed = bars + 2 * NUM_SLOTS;
    uint64_t *tmem_full = bars + 3 * NUM_SLOTS, *tmem_empty = bars + 3 * NUM_SLOTS + 1;
    uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(bars + 3 * NUM_SLOTS + 2);
    float *stage_base = reinterpret_cast<float *>(ring + RING_BYTES + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_SLOTS; ++s) {
            mbar_init(&full[s], PRODUCER_THREADS);   // producers: A converted into TMEM (and B landed)
            mbar_init(&empty[s], 1);                 // MMA commit: smem slot + TMEM A buffer may be overwritten
            mbar_init(&landed[s], 1);                // TMA bytes of this slot have landed
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, NUM_EPI_WARPS);
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_smem;
    const int total_tiles = Policy::num_tiles(p);
    // timing experiments only (PTGNN_TC_DEBUG; results are wrong when set): 1 = no MMAs, 2 = no loads, 4 = no stores,
    // 8 = no A conversion
    const int dbg = p.dbg;

    if (warp < 4) {
        // =========================================== PRODUCERS ===========================================
        // copy mapping: lane l of warp w stages 16-byte chunk q = l & 7 of rows 32 w + (l >> 3) + 4 i (i < 8), i.e. every
        // warp stages exactly the 32 rows it converts afterwards (only a __syncwarp between the two steps).
        const int q = lane & 7, rsub = warp * 32 + (lane >> 3);
        const int my_row = warp * 32 + lane;     // the row this thread converts (TMEM lane 32 * warp + lane)
        struct Cursor { int tile, seg, kc; };
        typename Policy::Tile t_load, t_proc;
        int rows_load[8];
        Cursor cl{(int)blockIdx.x, 0, 0}, cp{(int)blockIdx.x, 0, 0};
        bool load_valid = cl.tile < total_tiles, proc_valid = load_valid;
        uint32_t c_load = 0, c_proc = 0;

        auto load_rows = [&]() {
            if (Policy::segment(p, t_load, cl.seg).a_map != nullptr) return;
#pragma unroll
            for (int i = 0; i < 8; ++i) rows_load[i] = Policy::gather_row(p, t_load, cl.seg, rsub + 4 * i);
        };
        if (load_valid) { Policy::tile_setup(p, cl.tile, t_load); load_rows(); }
        if (proc_valid) Policy::tile_setup(p, cp.tile, t_proc);

        auto issue = [&]() {   // stage chunk (cl) into slot c_load % NUM_SLOTS
            const uint32_t slot = c_load % NUM_SLOTS, use = c_load / NUM_SLOTS;
            mbar_wait_warp(&empty[slot], (use & 1) ^ 1, lane);
            unsigned char *base = ring + slot * SLOT_BYTES;
            const Segment sg = Policy::segment(p, t_load, cl.seg);
            const int kchunk = cl.kc * CHUNK_K;
            if (threadIdx.x == 0) {   // bulk tensor copies: weights (hi, lo) and, for contiguous rows, the raw A tile
                if (dbg & 2) {
                    mbar_arrive(&landed[slot]);
                } else {
                    const uint32_t bytes = 2u * (uint32_t)sg.b_box_rows * 128u + (sg.a_map ? (uint32_t)OPERAND_BYTES : 0u);
                    mbar_expect_tx(&landed[slot], bytes);
                    if (sg.a_map) tma_load_2d(base, sg.a_map, kchunk, sg.a_row0, &landed[slot]);
                    tma_load_2d(base + OPERAND_BYTES, sg.b_hi_map, sg.b_col0 + kchunk, sg.b_row0, &landed[slot]);
                    tma_load_2d(base + 2 * OPERAND_BYTES, sg.b_lo_map, sg.b_col0 + kchunk, sg.b_row0, &landed[slot]);
                }
            }
            if (sg.a_map == nullptr && !(dbg & 2)) {   // gathered rows
                const int k0 = kchunk + q * 4;
                const bool k_ok = k0 < sg.K;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int g = rows_load[i];
                    const bool ok = k_ok && g >= 0;
                    cp_async16(smem_u32(base + swz(rsub + 4 * i, q)),
                               ok ? (const void *)(sg.a + (size_t)g * sg.lda + k0) : (const void *)sg.a, ok ? 16 : 0);
                }
            }
            ++c_load;
            ++cl.kc;
            if (cl.kc * CHUNK_K >= sg.K) {
                cl.kc = 0;
                ++cl.seg;
                if (cl.seg >= Policy::num_segments(p, t_load)) {
                    cl.seg = 0;
                    cl.tile += gridDim.x;
                    load_valid = cl.tile < total_tiles;
                    if (load_valid) Policy::tile_setup(p, cl.tile, t_load);
                }
                if (load_valid) load_rows();
            }
        };

#pragma unroll
        for (int i = 0; i < LOOKAHEAD; ++i) {
            if (load_valid) issue();
            cp_async_commit();
        }
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        while (proc_valid) {
            const uint32_t slot = c_proc % NUM_SLOTS, use = c_proc / NUM_SLOTS;
            cp_async_wait<LOOKAHEAD - 1>();          // this thread's gathered pieces of chunk c_proc
            __syncwarp();                            // ... and those of the other lanes of this warp (same 32 rows)
            mbar_wait_warp(&landed[slot], use & 1, lane);   // TMA tiles of chunk c_proc
            // convert this thread's row: 32 raw floats -> TF32 hi / lo -> TMEM columns of slot's A buffer
            const unsigned char *base = ring + slot * SLOT_BYTES;
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v = *reinterpret_cast<const float4 *>(base + swz(my_row, j));
                hi[4 * j] = tf32_hi(v.x); hi[4 * j + 1] = tf32_hi(v.y); hi[4 * j + 2] = tf32_hi(v.z); hi[4 * j + 3] = tf32_hi(v.w);
                lo[4 * j] = v.x - hi[4 * j]; lo[4 * j + 1] = v.y - hi[4 * j + 1];
                lo[4 * j + 2] = v.z - hi[4 * j + 2]; lo[4 * j + 3] = v.w - hi[4 * j + 3];
            }
            const uint32_t a_buf = tmem_lane + A_TMEM_OFF + slot * 64;
            if (!(dbg & 8)) {
                tmem_st_32cols(a_buf, hi);
                tmem_st_32cols(a_buf + 32, lo);
                tmem_st_wait();
            }
            tc_fence_before_sync();
            mbar_arrive(&full[slot]);
            ++c_proc;
            if (load_valid) issue();
            cp_async_commit();
            const Segment sg = Policy::segment(p, t_proc, cp.seg);
            ++cp.kc;
            if (cp.kc * CHUNK_K >= sg.K) {
                cp.kc = 0;
                ++cp.seg;
                if (cp.seg >= Policy::num_segments(p, t_proc)) {
                    cp.seg = 0;
                    cp.tile += gridDim.x;
                    proc_valid = cp.tile < total_tiles;
                    if (proc_valid) Policy::tile_setup(p, cp.tile, t_proc);
                }
            }
        }
        cp_async_wait<0>();
    } else if (warp == MMA_WARP) {
        // =========================================== MMA ISSUER ===========================================
        if (lane == 0) {
            uint32_t c = 0, tcount = 0;
            typename Policy::Tile t;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
                Policy::tile_setup(p, tile, t);
                mbar_wait(tmem_empty, (tcount & 1) ^ 1);     // epilogue has drained the previous tile's accumulators
                tc_fence_after_sync();
                const int nseg = Policy::num_segments(p, t);
                for (int seg = 0; seg < nseg; ++seg) {
                    const Segment sg = Policy::segment(p, t, seg);
                    MmaGroup g[2];
                    const int ng = Policy::mma_groups(p, t, seg, g);
                    const int nkc = (sg.K + CHUNK_K - 1) / CHUNK_K;
                    for (int kc = 0; kc < nkc; ++kc, ++c) {
                        const uint32_t slot = c % NUM_SLOTS, use = c / NUM_SLOTS;
                        mbar_wait(&full[slot], use & 1);
                        tc_fence_after_sync();
                        const uint32_t base = smem_u32(ring + slot * SLOT_BYTES);
                        const uint32_t a_buf = tmem_base + A_TMEM_OFF + slot * 64;
                        const int kvalid = min(CHUNK_K, sg.K - kc * CHUNK_K);
                        const int ksteps = (kvalid + 7) / 8;
                        for (int ks = 0; ks < ksteps && !(dbg & 1); ++ks) {
                            const uint32_t a_hi = a_buf + ks * 8, a_lo = a_buf + 32 + ks * 8;
                            for (int gi = 0; gi < ng; ++gi) {
                                const uint64_t b_hi = make_smem_desc_sw128(base + OPERAND_BYTES + g[gi].row_off * 128 + ks * 32);
                                const uint64_t b_lo = make_smem_desc_sw128(base + 2 * OPERAND_BYTES + g[gi].row_off * 128 + ks * 32);
                                const uint32_t idesc = make_instr_desc(FMT_TF32, TILE_M, (uint32_t)g[gi].n);
                                const uint32_t d_main = tmem_base + g[gi].col_off;
                                const uint32_t d_corr = d_main + CORR_OFF;
                                const bool first = g[gi].fresh && kc == 0 && ks == 0;
                                mma_tf32_ts(d_main, a_hi, b_hi, idesc, first ? 0u : 1u);
                                mma_tf32_ts(d_corr, a_hi, b_lo, idesc, first ? 0u : 1u);
                                mma_tf32_ts(d_corr, a_lo, b_hi, idesc, 1u);
                            }
                        }
                        mma_commit(&empty[slot]);
                    }
                }
                mma_commit(tmem_full);
            }
        }
        __syncwarp();
    } else {
        // =========================================== EPILOGUE ===========================================
        const int ew = warp - 5;                 // 0..7
        const int quarter = warp & 3;            // TMEM lanes 32*quarter .. +31 are the ones this warp may read
        const int half = ew >> 2;                // which half of the accumulator columns this warp drains
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float *stage = stage_base + ew * (STAGE_BYTES_PER_WARP / 4);
        uint32_t tcount = 0;
        typename Policy::Tile t;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            Policy::tile_setup(p, tile, t);
            mbar_wait_warp(tmem_full, tcount & 1, lane);
            tc_fence_after_sync();
            float acc[64];
            Policy::drain(p, t, tmem_lane, half, acc);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);          // the next tile's MMAs may start while we store
            if (!(dbg & 4)) Policy::store(p, t, acc, quarter, half, lane, stage);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace tc
}  // namespace ptgnn
