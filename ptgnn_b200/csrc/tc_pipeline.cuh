// Persistent, warp-specialised tcgen05 pipeline shared by the three GEMM-bearing kernels of the hot path
// (per-edge messages, GRUCell update, Mlp dense update).
//
//   warps 0-3  PRODUCERS  gather fp32 rows (node states / aggregates) with cp.async straight into the swizzled
//                         A tile, stream the pre-split weight tile (hi, lo), split A into TF32 hi/lo IN PLACE,
//                         fence.proxy.async, arrive on full[slot]
//   warp  4    MMA        one thread: wait full[slot], issue 3 tcgen05.mma (hi*hi, hi*lo, lo*hi) per K=8 step into
//                         the TMEM accumulator, tcgen05.commit -> empty[slot]; per tile commit -> tmem_full[acc]
//   warps 5-8  EPILOGUE   wait tmem_full[acc], tcgen05.ld the 128 x N fp32 tile (row per thread), apply the policy's
//                         epilogue (scatter message rows / GRU gate math / bias+activation), arrive tmem_empty[acc]
//
// One CTA per SM (grid = #SMs), static round-robin over tiles; the shared-memory ring (3 slots x 64 KB) and the
// double-buffered TMEM accumulator let the loads, the MMAs and the epilogue of neighbouring tiles overlap.  Every
// mbarrier wait is bounded (tc_common.cuh) so a protocol bug traps instead of hanging the GPU.
//
// A Policy supplies:
//   struct Params;                                              (passed as __grid_constant__)
//   static constexpr int ACC_COLS;                              TMEM columns per accumulator (<= 128)
//   struct Tile { ... };                                        per-tile registers
//   __device__ static int  num_tiles(const Params&);
//   __device__ static void tile_setup(const Params&, int tile, Tile&);
//   __device__ static int  num_segments(const Params&, const Tile&);
//   __device__ static Segment segment(const Params&, const Tile&, int seg);
//   __device__ static int  gather_row(const Params&, const Tile&, int seg, int r);   r in [0,128) -> A row or -1
//   __device__ static int  mma_groups(const Params&, const Tile&, int seg, MmaGroup (&g)[2]);
//   __device__ static void epilogue(const Params&, const Tile&, uint32_t tmem_acc, int quarter, int lane);
#pragma once
#include "tc_common.cuh"

namespace ptgnn {
namespace tc {

constexpr int TILE_M = 128;
constexpr int CHUNK_K = 32;                       // fp32 per k-chunk = one 128-byte swizzled row
constexpr int NUM_SLOTS = 3;
constexpr int LOOKAHEAD = 2;                      // chunks of loads in flight ahead of the chunk being split
constexpr int PRODUCER_THREADS = 128;
constexpr int MMA_WARP = 4;
constexpr int FIRST_EPI_WARP = 5;
constexpr int NUM_THREADS = 9 * 32;
constexpr int OPERAND_BYTES = TILE_M * CHUNK_K * 4;   // 16 KB: one 128 x 32 fp32 operand tile
constexpr int SLOT_BYTES = 4 * OPERAND_BYTES;         // A_hi | A_lo | B_hi | B_lo
constexpr int RING_BYTES = NUM_SLOTS * SLOT_BYTES;
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/;

struct Segment {        // one K-range of the tile's GEMM: A rows from `a` (row pitch lda), B rows from b_hi/b_lo
    const float *a;
    const float *b_hi;
    const float *b_lo;
    int lda, ldb;
    int K;              // columns of this segment (multiple of 4)
    int b_rows;         // rows of B to stage (<= 128, multiple of 8)
};
struct MmaGroup {       // one tcgen05.mma per K-step: B rows [row_off, row_off + n) -> accumulator columns [col_off, +n)
    int n, row_off, col_off;
    bool fresh;         // true: the first K-step of this segment overwrites the accumulator columns
};

__device__ __forceinline__ uint32_t swz(int row, int q) { return (uint32_t)(row * 128 + ((q ^ (row & 7)) << 4)); }

template <class Policy>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_pipeline_kernel(const __grid_constant__ typename Policy::Params p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned ring (SWIZZLE_128B descriptors assume base_offset = 0)
    unsigned char *ring = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + RING_BYTES);
    uint64_t *full = bars, *empty = bars + NUM_SLOTS, *tmem_full = bars + 2 * NUM_SLOTS, *tmem_empty = bars + 2 * NUM_SLOTS + 2;
    uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(bars + 2 * NUM_SLOTS + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NUM_SLOTS; ++s) { mbar_init(&full[s], PRODUCER_THREADS); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        mbar_init_fence();
    }
    if (warp == 0) tmem_alloc<256>(tmem_base_smem);
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
            const int k0 = cl.kc * CHUNK_K + q * 4;
            const bool k_ok = k0 < sg.K;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                const int g = rows_load[i];
                const bool ok = k_ok && g >= 0;
                cp_async16(smem_u32(base + swz(r, q)), ok ? (const void *)(sg.a + (size_t)g * sg.lda + k0) : (const void *)sg.a, ok ? 16 : 0);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                if (r < sg.b_rows) {
                    const size_t off = (size_t)r * sg.ldb + k0;
                    cp_async16(smem_u32(base + 2 * OPERAND_BYTES + swz(r, q)), k_ok ? (const void *)(sg.b_hi + off) : (const void *)sg.b_hi, k_ok ? 16 : 0);
                    cp_async16(smem_u32(base + 3 * OPERAND_BYTES + swz(r, q)), k_ok ? (const void *)(sg.b_lo + off) : (const void *)sg.b_lo, k_ok ? 16 : 0);
                }
            }
            ++c_load;
            // advance the load cursor
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
            cp_async_wait<LOOKAHEAD - 1>();
            // split this thread's A pieces of chunk c_proc in place: raw -> hi (same spot), lo (A_lo tile)
            unsigned char *base = ring + (c_proc % NUM_SLOTS) * SLOT_BYTES;
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
            mbar_arrive(&full[c_proc % NUM_SLOTS]);
            ++c_proc;
            if (load_valid) issue();
            cp_async_commit();
            // advance the processing cursor
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
                const uint32_t tmem_acc = tmem_base + acc * 128;
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
                        const int ksteps = min(CHUNK_K, sg.K - kc * CHUNK_K) / 8 + ((min(CHUNK_K, sg.K - kc * CHUNK_K) % 8) ? 1 : 0);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t a_hi = make_smem_desc_sw128(base + ks * 32);
                            const uint64_t a_lo = make_smem_desc_sw128(base + OPERAND_BYTES + ks * 32);
                            for (int gi = 0; gi < ng; ++gi) {
                                const uint64_t b_hi = make_smem_desc_sw128(base + 2 * OPERAND_BYTES + g[gi].row_off * 128 + ks * 32);
                                const uint64_t b_lo = make_smem_desc_sw128(base + 3 * OPERAND_BYTES + g[gi].row_off * 128 + ks * 32);
                                const uint32_t idesc = make_instr_desc(FMT_TF32, TILE_M, (uint32_t)g[gi].n);
                                const uint32_t d = tmem_acc + g[gi].col_off;
                                const uint32_t first = (g[gi].fresh && kc == 0 && ks == 0) ? 0u : 1u;
                                mma_tf32_ss(d, a_hi, b_hi, idesc, first);
                                mma_tf32_ss(d, a_hi, b_lo, idesc, 1u);
                                mma_tf32_ss(d, a_lo, b_hi, idesc, 1u);
                            }
                        }
                        mma_commit(&empty[slot]);
                    }
                }
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
            Policy::epilogue(p, t, tmem_base + acc * 128 + ((uint32_t)(quarter * 32) << 16), quarter, lane);
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace tc
}  // namespace ptgnn
