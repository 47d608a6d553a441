// Persistent, warp-specialised tcgen05 pipeline shared by the three GEMM-bearing kernels of the hot path
// (per-edge messages, GRUCell update, Mlp dense update).
//
//   warps 0-3  PRODUCERS  A tile: either gathered fp32 rows via cp.async (per-edge messages) or a TMA tile load
//                         (contiguous node rows); B tile: the pre-split weight tile (hi, lo) via TMA.  Then the raw A
//                         tile is split into TF32 hi/lo IN PLACE, fence.proxy.async, arrive on full[slot].
//   warp  4    MMA        one thread: wait full[slot]; per K=8 step issue hi*hi into the MAIN accumulator and
//                         hi*lo + lo*hi into the CORRECTION accumulator (tensor-core accumulation truncates, so the
//                         tiny terms must not perturb the main sum); tcgen05.commit -> empty[slot]; per tile
//                         commit -> tmem_full[acc].
//   warps 5-8  EPILOGUE   wait tmem_full[acc]; tcgen05.ld main + correction (row per thread), policy epilogue
//                         (scatter message rows / GRU gate math / bias+activation); arrive tmem_empty[acc].
//
// One CTA per SM (grid = #SMs), static round-robin over tiles.  Shared memory: 3-slot ring x 64 KB
// (A_hi | A_lo | B_hi | B_lo, SWIZZLE_128B K-major).  TMEM: 512 columns = 2 accumulator sets x (128 main + 128 corr),
// so loads, MMAs and the epilogue of neighbouring tiles overlap.  Every mbarrier wait is bounded (tc_common.cuh).
//
// A Policy supplies:
//   struct Params;   struct Tile;
//   __device__ static int  num_tiles(const Params&);
//   __device__ static void tile_setup(const Params&, int tile, Tile&);
//   __device__ static int  num_segments(const Params&, const Tile&);
//   __device__ static Segment segment(const Params&, const Tile&, int seg);
//   __device__ static int  gather_row(const Params&, const Tile&, int seg, int r);   only when segment.a_map == nullptr
//   __device__ static int  mma_groups(const Params&, const Tile&, int seg, MmaGroup (&g)[2]);
//   __device__ static void epilogue(const Params&, const Tile&, uint32_t tmem_acc, int quarter, int lane, float* stage);
//                          (stage = this warp's 4 KB transpose buffer: TMEM hands every thread one ROW, global memory
//                           wants warps to touch one row at a time -- see warp_store_rows / warp_load_rows)
#pragma once
#include <cuda.h>

#include "tc_common.cuh"

namespace ptgnn {
namespace tc {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 32;                       // fp32 per k-chunk = one 128-byte swizzled row
constexpr int NUM_SLOTS = 3;
constexpr int LOOKAHEAD = 2;                      // chunks of loads in flight ahead of the chunk being split
constexpr int PRODUCER_THREADS = 128;
constexpr int MMA_WARP = 4;
constexpr int NUM_THREADS = 9 * 32;
constexpr int OPERAND_BYTES = TILE_M * CHUNK_K * 4;   // 16 KB: one 128 x 32 fp32 operand tile
constexpr int SLOT_BYTES = 4 * OPERAND_BYTES;         // A_hi | A_lo | B_hi | B_lo
constexpr int RING_BYTES = NUM_SLOTS * SLOT_BYTES;
constexpr int STAGE_BYTES_PER_WARP = 32 * 32 * 4;           // epilogue transpose buffer: 32 rows x 32 fp32
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/ + 4 * STAGE_BYTES_PER_WARP;
constexpr int ACC_SET_COLS = 256;                     // 128 main + 128 correction
constexpr int CORR_OFF = 128;

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

template <class Policy>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_pipeline_kernel(const __grid_constant__ typename Policy::Params p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned ring (SWIZZLE_128B descriptors / TMA swizzle assume base_offset = 0)
    unsigned char *ring = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + RING_BYTES);
    uint64_t *full = bars, *empty = bars + NUM_SLOTS, *landed = bars + 2 * NUM_SLOTS;
    uint64_t *tmem_full = bars + 3 * NUM_SLOTS, *tmem_empty = bars + 3 * NUM_SLOTS + 2;
    uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(bars + 3 * NUM_SLOTS + 4);
    float *stage_base = reinterpret_cast<float *>(ring + RING_BYTES + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_SLOTS; ++s) {
            mbar_init(&full[s], PRODUCER_THREADS);   // producers: tile split + fenced
            mbar_init(&empty[s], 1);                 // MMA commit: slot may be overwritten
            mbar_init(&landed[s], 1);                // TMA bytes of this slot have landed
        }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc<512>(tmem_base_smem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_smem;
    const int total_tiles = Policy::num_tiles(p);

    if (warp < 4) {
        // =========================================== PRODUCERS ===========================================
        const int pt = threadIdx.x;              // 0..127
        const int q = pt & 7, rbase = pt >> 3;   // this thread owns 16-byte chunk q of rows rbase + 16*i
        struct Cursor { int tile, seg, kc; };
        typename Policy::Tile t_load, t_proc;
        int rows_load[8];
        Cursor cl{(int)blockIdx.x, 0, 0}, cp{(int)blockIdx.x, 0, 0};
        bool load_valid = cl.tile < total_tiles, proc_valid = load_valid;
        uint32_t c_load = 0, c_proc = 0;

        auto load_rows = [&]() {
            if (Policy::segment(p, t_load, cl.seg).a_map != nullptr) return;
#pragma unroll
            for (int i = 0; i < 8; ++i) rows_load[i] = Policy::gather_row(p, t_load, cl.seg, rbase + 16 * i);
        };
        if (load_valid) { Policy::tile_setup(p, cl.tile, t_load); load_rows(); }
        if (proc_valid) Policy::tile_setup(p, cp.tile, t_proc);

        auto issue = [&]() {   // stage chunk (cl) into slot c_load % NUM_SLOTS
            const uint32_t slot = c_load % NUM_SLOTS, use = c_load / NUM_SLOTS;
            mbar_wait(&empty[slot], (use & 1) ^ 1);
            unsigned char *base = ring + slot * SLOT_BYTES;
            const Segment sg = Policy::segment(p, t_load, cl.seg);
            const int kchunk = cl.kc * CHUNK_K;
            if (pt == 0) {   // bulk tensor copies: weights (hi, lo) and, for contiguous rows, the raw A tile
                const uint32_t bytes = 2u * (uint32_t)sg.b_box_rows * 128u + (sg.a_map ? (uint32_t)OPERAND_BYTES : 0u);
                mbar_expect_tx(&landed[slot], bytes);
                if (sg.a_map) tma_load_2d(base, sg.a_map, kchunk, sg.a_row0, &landed[slot]);
                tma_load_2d(base + 2 * OPERAND_BYTES, sg.b_hi_map, sg.b_col0 + kchunk, sg.b_row0, &landed[slot]);
                tma_load_2d(base + 3 * OPERAND_BYTES, sg.b_lo_map, sg.b_col0 + kchunk, sg.b_row0, &landed[slot]);
            }
            if (sg.a_map == nullptr) {   // gathered rows
                const int k0 = kchunk + q * 4;
                const bool k_ok = k0 < sg.K;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int g = rows_load[i];
                    const bool ok = k_ok && g >= 0;
                    cp_async16(smem_u32(base + swz(rbase + 16 * i, q)),
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
        while (proc_valid) {
            const uint32_t slot = c_proc % NUM_SLOTS, use = c_proc / NUM_SLOTS;
            cp_async_wait<LOOKAHEAD - 1>();          // this thread's gathered pieces of chunk c_proc
            mbar_wait(&landed[slot], use & 1);       // TMA tiles of chunk c_proc
            // split this thread's A pieces in place: raw -> hi (same spot), lo (A_lo tile)
            unsigned char *base = ring + slot * SLOT_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 *ph = reinterpret_cast<float4 *>(base + swz(rbase + 16 * i, q));
                float4 *pl = reinterpret_cast<float4 *>(base + OPERAND_BYTES + swz(rbase + 16 * i, q));
                const float4 v = *ph;
                float4 hi, lo;
                hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
                lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
                *ph = hi;
                *pl = lo;
            }
            fence_proxy_async_smem();
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
                const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
                mbar_wait(&tmem_empty[acc], (acc_use & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_acc = tmem_base + acc * ACC_SET_COLS;
                const int nseg = Policy::num_segments(p, t);
                bool corr_written[2] = {false, false};   // per group slot: has the correction accumulator been initialised?
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
                        const int kvalid = min(CHUNK_K, sg.K - kc * CHUNK_K);
                        const int ksteps = (kvalid + 7) / 8;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t a_hi = make_smem_desc_sw128(base + ks * 32);
                            const uint64_t a_lo = make_smem_desc_sw128(base + OPERAND_BYTES + ks * 32);
                            for (int gi = 0; gi < ng; ++gi) {
                                const uint64_t b_hi = make_smem_desc_sw128(base + 2 * OPERAND_BYTES + g[gi].row_off * 128 + ks * 32);
                                const uint64_t b_lo = make_smem_desc_sw128(base + 3 * OPERAND_BYTES + g[gi].row_off * 128 + ks * 32);
                                const uint32_t idesc = make_instr_desc(FMT_TF32, TILE_M, (uint32_t)g[gi].n);
                                const uint32_t d_main = tmem_acc + g[gi].col_off;
                                const uint32_t d_corr = d_main + CORR_OFF;
                                const bool first = g[gi].fresh && kc == 0 && ks == 0;
                                mma_tf32_ss(d_main, a_hi, b_hi, idesc, first ? 0u : 1u);
                                mma_tf32_ss(d_corr, a_hi, b_lo, idesc, first ? 0u : 1u);
                                mma_tf32_ss(d_corr, a_lo, b_hi, idesc, 1u);
                            }
                        }
                        mma_commit(&empty[slot]);
                    }
                }
                (void)corr_written;
                mma_commit(&tmem_full[acc]);
            }
        }
        __syncwarp();
    } else {
        // =========================================== EPILOGUE ===========================================
        const int quarter = warp & 3;   // TMEM lanes 32*quarter .. +31 are the ones this warp may read
        uint32_t tcount = 0;
        typename Policy::Tile t;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            Policy::tile_setup(p, tile, t);
            const uint32_t acc = tcount & 1, acc_use = tcount >> 1;
            mbar_wait(&tmem_full[acc], acc_use & 1);
            tc_fence_after_sync();
            Policy::epilogue(p, t, tmem_base + acc * ACC_SET_COLS + ((uint32_t)(quarter * 32) << 16), quarter, lane,
                             stage_base + quarter * (STAGE_BYTES_PER_WARP / 4));
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ---- epilogue transposes through shared memory ---------------------------------------------------------------
// The accumulator comes out of TMEM one ROW per thread; storing that directly makes every warp-wide store touch 32
// different 128-byte lines.  Staging 32 rows x NCOLS through a (chunk-XOR-swizzled) buffer lets each store
// instruction write whole rows: 4 (NCOLS = 32) or 8 (NCOLS = 16) lines per instruction instead of 32.
// `row_off` is this lane's destination element offset from `dst_base` (negative = row not stored).
template <int NCOLS>
__device__ __forceinline__ void warp_store_rows(float *stage, const float (&v)[NCOLS], float *dst_base, long long row_off,
                                                int lane) {
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
__device__ __forceinline__ void warp_load_rows(float *stage, float (&v)[NCOLS], const float *src_base, long long row_off,
                                               int lane) {
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

// accumulator value = main + correction (two TMEM loads)
__device__ __forceinline__ void tmem_ld_acc16(uint32_t taddr, float (&v)[16]) {
    float c[16];
    tmem_ld_16cols(taddr, v);
    tmem_ld_16cols(taddr + CORR_OFF, c);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += c[i];
}
__device__ __forceinline__ void tmem_ld_acc32(uint32_t taddr, float (&v)[32]) {
    float c[32];
    tmem_ld_32cols(taddr, v);
    tmem_ld_32cols(taddr + CORR_OFF, c);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += c[i];
}

}  // namespace tc
}  // namespace ptgnn
