// Fused gather -> per-edge-type Linear -> segmented reduce on tcgen05 (see fused_mp.cuh for the design).
#include "fused_mp.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace ptgnn {
namespace fused {

using tc::mbar_arrive;
using tc::mbar_init;
using tc::mbar_wait;

// ---- geometry ------------------------------------------------------------------------------------------------------
constexpr int NUM_THREADS = 16 * 32;
constexpr int MMA_WARP = 4;
constexpr int GATHER_WARP0 = 5, GATHER_THREADS = 64;
constexpr int SCHED_WARP = 7;
constexpr int EPI_THREADS = 256;                       // two epilogue warpgroups (warps 0-3, 8-11): lower / upper half of a block's targets
constexpr int NUM_CONSUMER_WARPS = 4 + 1 + 2 + 8;      // weight loaders, MMA, gatherers, epilogue (scheduler table readers)
constexpr int NUM_SLOTS = 3, LOOKAHEAD = 2;
constexpr int SLOT_BYTES = 32768;
constexpr int RING_BYTES = NUM_SLOTS * SLOT_BYTES;
constexpr int META_RING = 8;
constexpr int SCHED_RING = 4;
// 512 threads launch with 128 registers each; warps 4-7 and the weight loaders give registers back, the two epilogue warpgroups take them:
// (2 * 160 + 72 + 120) * 128 = 65536.  ptxas only extends a role's budget beyond the launch cap when that role is the LAST branch.
constexpr int EPI_REGS = 160, W_REGS = 120, MID_REGS = 72;
constexpr int ACC_TMEM_OFF = 256;                               // weight buffers below, accumulators above
constexpr int EPI_BAR_ID = 2;

struct Meta {                     // per sub-group: what the epilogue needs to know about the accumulator columns
    int32_t tloff[128];           // byte offset of the column's target row inside agg_s (0 for columns >= n)
    uint32_t endmask[4];          // bit c: column c is the last edge of its (target, type) segment
    uint32_t lowmask[4];          // bit c: column c's target lies in the lower half of the block (epilogue warpgroup 0's columns)
    int32_t n, pad[3];
};
struct Sched {                    // one target block: its id and the T+1 sorted-edge offsets of its (block, type) groups
    int32_t blk, pad[3];
    int32_t off[PTGNN_MAX_EDGE_TYPES + 4];
};
constexpr int AGG_OFF = RING_BYTES;
__host__ __device__ constexpr int agg_bytes(int B) { return B * kD * 4; }
constexpr int SMEM_TAIL = META_RING * (int)sizeof(Meta) + SCHED_RING * (int)sizeof(Sched) + 256 /*barriers*/;
constexpr int smem_bytes(int B) { return 1024 + RING_BYTES + agg_bytes(B) + SMEM_TAIL; }
static_assert(smem_bytes(kMaxBlockTargets) <= 232448, "shared memory budget");

struct Params {
    const unsigned char *src_rows, *tgt_rows;
    const uint4 *wpack;
    const int32_t *group_off, *src_f;
    const uint8_t *tl_f;
    const int32_t *row_ptr;
    void *out;
    int num_nodes, num_blocks, B, T, reduce, out_mode;
    int32_t *status;
    unsigned long long *trace;     // debug (PTGNN_FUSED_TRACE=1): per-role event timeline of CTA 0, else nullptr
    Epilogue epi;
};

// ---- small PTX helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32cols_u32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ uint32_t swz(int row, int q) { return (uint32_t)(row * 128 + ((q ^ (row & 7)) << 4)); }

// Up to 32 mbarriers polled by ONE try_wait instruction per round: lane i checks (addr, parity) of its own barrier (inactive lanes
// report done).  A phase check costs ~200 cycles even when the barrier is already complete (measured, tools/fused_trace.py), so
// the MMA warp's three per-step waits (accumulator drained, rows landed, weights landed) are folded into one.
__device__ __forceinline__ void mbar_wait_lanes(uint32_t addr, uint32_t parity, bool active) {
    uint64_t t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done = 1;
        if (active)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (__all_sync(0xffffffffu, done != 0)) return;
        if ((spin & 0xFF) == 0xFF) {
            const uint64_t now = tc::global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}

// Debug timeline (PTGNN_FUSED_TRACE=1): CTA 0, one thread per role, records (clock64, step, tag) at the pipeline hand-offs into its
// 2048-entry region of the trace buffer; read back with ptgnn_b200_debug_fused_trace (tools/fused_trace.py).
struct Trace {
    unsigned long long *buf;
    int n;
    __device__ __forceinline__ void mark(int tag, uint32_t step) {
        if (buf != nullptr && n < 2048) buf[n++] = ((unsigned long long)clock64() << 24) | ((unsigned long long)(step & 0xFFFFu) << 8) | (unsigned)(tag & 0xFF);
    }
};
__device__ __forceinline__ Trace make_trace(unsigned long long *base, int role, bool on) {
    return Trace{(base != nullptr && blockIdx.x == 0 && on) ? base + role * 2048 : nullptr, 0};
}

// ---- the (block, group, sub-group, segment) walk every role performs in the same order -----------------------------------
struct Step { int blk, t, e, n, seg; bool first_sub, last_sub; };
template <int NSEG>
struct StepGen {
    const Sched *sched;
    uint64_t *sfull, *sempty;
    int T, nmax, lane;
    uint32_t it = 0;
    const Sched *tab = nullptr;
    int blk = -1, t = 0, e0 = 0, e = 0, e1 = 0, seg = 0;
    bool in_block = false;
    // 0: `s` is the next step | 1: the block s.blk has no more steps | 2: no more blocks
    __device__ __forceinline__ int next(Step &s) {
        for (;;) {
            if (!in_block) {
                const uint32_t r = it % SCHED_RING;
                mbar_wait(&sfull[r], (it / SCHED_RING) & 1);
                tab = &sched[r];
                blk = tab->blk;
                if (blk < 0) return 2;
                in_block = true;
                t = -1; e = e1 = 0; seg = 0;
            }
            if (e < e1) {
                s.blk = blk; s.t = t; s.e = e; s.n = min(nmax, e1 - e); s.seg = seg;
                s.first_sub = e == e0; s.last_sub = e + nmax >= e1;
                if (++seg == NSEG) { seg = 0; e += nmax; }
                return 0;
            }
            ++t;
            while (t < T && tab->off[t + 1] == tab->off[t]) ++t;
            if (t >= T) {
                in_block = false;
                s.blk = blk;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sempty[it % SCHED_RING]);     // this warp no longer reads the table
                ++it;
                return 1;
            }
            e0 = e = tab->off[t]; e1 = tab->off[t + 1]; seg = 0;
        }
    }
};

// t = op(prev, v) unless the column starts a segment: ONE predicated instruction (a select would put two ALU latencies per
// column on the serial chain)
template <int RED>
__device__ __forceinline__ void continue_segment(float &t, float prev, float v, uint32_t start_bit) {
    if (RED == PTGNN_REDUCE_MAX)
        asm("{\n\t.reg .pred pc;\n\tsetp.eq.u32 pc, %3, 0;\n\t@pc max.f32 %0, %1, %2;\n\t}" : "+f"(t) : "f"(prev), "f"(v), "r"(start_bit));
    else if (RED == PTGNN_REDUCE_MIN)
        asm("{\n\t.reg .pred pc;\n\tsetp.eq.u32 pc, %3, 0;\n\t@pc min.f32 %0, %1, %2;\n\t}" : "+f"(t) : "f"(prev), "f"(v), "r"(start_bit));
    else
        asm("{\n\t.reg .pred pc;\n\tsetp.eq.u32 pc, %3, 0;\n\t@pc add.f32 %0, %1, %2;\n\t}" : "+f"(t) : "f"(prev), "f"(v), "r"(start_bit));
}
__device__ __forceinline__ void sts_f32_if(uint32_t addr, float v, uint32_t bit) {
    asm volatile("{\n\t.reg .pred pe;\n\tsetp.ne.u32 pe, %2, 0;\n\t@pe st.shared.f32 [%0], %1;\n\t}" ::"r"(addr), "f"(v), "r"(bit) : "memory");
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// x -> (hi, lo') fp16 pair; returns false when |x| is not representable (>= 65504, inf, NaN)
__device__ __forceinline__ bool split_f16(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
    return fabsf(x) < 65504.0f;
}

template <int RED> __device__ __forceinline__ float red_identity() {
    return RED == PTGNN_REDUCE_MAX ? -FLT_MAX : (RED == PTGNN_REDUCE_MIN ? FLT_MAX : 0.0f);
}
template <int RED> __device__ __forceinline__ float red_op(float a, float m) {
    if (RED == PTGNN_REDUCE_MAX) return fmaxf(a, m);        // one FMNMX; a NaN message never wins (torch_scatter's strict compare)
    if (RED == PTGNN_REDUCE_MIN) return fminf(a, m);
    return a + m;
}

// =====================================================================================================================
// Write-out of a finished block: rows [row_lo, min(row_hi, rows of the block)) of agg_s, taken by the calling warp (ew of the four
// of its group) 2 rows at a time; a row is reset to the identity as soon as it has been read.  The row loop is specialised at compile
// time on the output format and on "plain sum" (no mean / max fix-up / activation / LayerNorm): the generic version executed ~180
// instructions per row.  Register pressure matters here: the epilogue's column loop sits at the 128-register limit, and a 4-row
// unroll (or a non-inlined call: ABI-constrained allocation) made ptxas spill around every LDTM of the column loop -- drain time
// per sub-group went from 1,200 to 2,300 cycles (sessions r02g/r02h) -- so check `-Xptxas -v` for 0 spills after touching this.
template <int RED>
__device__ __forceinline__ void write_out_block(const Params *p, uint32_t agg_saddr, int row0, int row_lo, int row_hi, int ew, int lane) {
    const float IDENT = red_identity<RED>();
    const int rows = min(p->B, p->num_nodes - row0);
    const int my_hi = min(row_hi, rows);
    const float4 ident4 = make_float4(IDENT, IDENT, IDENT, IDENT);
    auto finish_row = [&](auto mode_tag, auto plain_tag, int r, float4 a) {
        constexpr int MODE = decltype(mode_tag)::value;
        constexpr bool PLAIN = decltype(plain_tag)::value;
        const int v = row0 + r;
        if (!PLAIN) {
            if (RED == PTGNN_REDUCE_MEAN) {
                const int cnt = __ldg(p->row_ptr + v + 1) - __ldg(p->row_ptr + v);
                const float c = (float)(cnt < 1 ? 1 : cnt);
                a.x /= c; a.y /= c; a.z /= c; a.w /= c;
            }
            if (RED == PTGNN_REDUCE_MAX || RED == PTGNN_REDUCE_MIN) {   // never updated -> 0 (torch_scatter)
                if (a.x == IDENT) a.x = 0.0f;
                if (a.y == IDENT) a.y = 0.0f;
                if (a.z == IDENT) a.z = 0.0f;
                if (a.w == IDENT) a.w = 0.0f;
            }
            if (p->epi.act != PTGNN_ACT_NONE) {
                a.x = apply_act(a.x, p->epi.act); a.y = apply_act(a.y, p->epi.act);
                a.z = apply_act(a.z, p->epi.act); a.w = apply_act(a.w, p->epi.act);
            }
            if (p->epi.ln_w != nullptr) {       // LayerNorm over the 128 features of the row (same order as reduce.cuh)
                float sum = (a.x + a.y) + (a.z + a.w);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float mean = sum / (float)kD;
                const float dx = a.x - mean, dy = a.y - mean, dz = a.z - mean, dw = a.w - mean;
                float qq = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
                const float rstd = rsqrtf(qq / (float)kD + p->epi.ln_eps);
                const float4 w = *reinterpret_cast<const float4 *>(p->epi.ln_w + lane * 4);
                const float4 b = *reinterpret_cast<const float4 *>(p->epi.ln_b + lane * 4);
                a.x = dx * rstd * w.x + b.x; a.y = dy * rstd * w.y + b.y;
                a.z = dz * rstd * w.z + b.z; a.w = dw * rstd * w.w + b.w;
            }
        }
        if (MODE == 1) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y), hi = __floats2bfloat162_rn(a.z, a.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t *>(&lo); pk.y = *reinterpret_cast<uint32_t *>(&hi);
            reinterpret_cast<uint2 *>(p->out)[(size_t)v * (kD / 4) + lane] = pk;
        } else if (MODE == 2) {      // fp16 (hi | lo') row: hi halfs at [0, 128), lo' halfs at [128, 256); packed conversions
            const __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn((a.x - f01.x) * 2048.0f, (a.y - f01.y) * 2048.0f);
            const __half2 l23 = __floats2half2_rn((a.z - f23.x) * 2048.0f, (a.w - f23.y) * 2048.0f);
            uint2 *row = reinterpret_cast<uint2 *>(p->out) + (size_t)v * (2 * kD / 4);
            row[lane] = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
            row[kD / 4 + lane] = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
            const float big = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
            if (!(big < 65504.0f) && p->status != nullptr) *reinterpret_cast<volatile int32_t *>(p->status) = 1;
        } else {
            reinterpret_cast<float4 *>(p->out)[(size_t)v * (kD / 4) + lane] = a;
        }
    };
    // 2 rows per iteration (independent loads in flight); rows are reset as they are read: the next block needs no initialisation pass
    auto write_rows = [&](auto mode_tag, auto plain_tag) {
        for (int r0 = row_lo + ew; r0 < my_hi; r0 += 4 * 2) {
            float4 v4[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = r0 + 4 * u;
                if (r < my_hi) {
                    const uint32_t rowa = agg_saddr + (uint32_t)(r * kD + lane * 4) * 4u;
                    v4[u] = lds_f32x4(rowa);
                    sts_f32x4(rowa, ident4);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (r0 + 4 * u < my_hi) finish_row(mode_tag, plain_tag, r0 + 4 * u, v4[u]);
        }
    };
    const bool plain = RED == PTGNN_REDUCE_SUM && p->epi.act == PTGNN_ACT_NONE && p->epi.ln_w == nullptr;
    using std::integral_constant;
    if (p->out_mode == 2) { if (plain) write_rows(integral_constant<int, 2>{}, integral_constant<bool, true>{}); else write_rows(integral_constant<int, 2>{}, integral_constant<bool, false>{}); }
    else if (p->out_mode == 1) { if (plain) write_rows(integral_constant<int, 1>{}, integral_constant<bool, true>{}); else write_rows(integral_constant<int, 1>{}, integral_constant<bool, false>{}); }
    else { if (plain) write_rows(integral_constant<int, 0>{}, integral_constant<bool, true>{}); else write_rows(integral_constant<int, 0>{}, integral_constant<bool, false>{}); }
}

template <int NPROD, int K, int NSEG, int RED>
__global__ void __launch_bounds__(NUM_THREADS, 1) fused_aggregate_kernel(const __grid_constant__ Params p) {
    constexpr int NPART = NPROD == 3 ? 2 : 1;                      // hi | lo' parts of a row / of the weights
    constexpr int NMAX = NPROD == 3 ? 64 : (K <= 128 ? 128 : 64);  // edges per MMA (accumulator columns)
    constexpr int ROW_BYTES = K * 2 * NPART;                       // one packed state row
    constexpr int KCH = K / 64;                                    // 128-byte swizzled chunks per part
    constexpr int NT = NPART * KCH;                                // operand tiles per slot
    constexpr int TILE_BYTES = NMAX * 128;
    static_assert(NT * TILE_BYTES <= SLOT_BYTES, "slot size");
    constexpr int WPART_COLS = K / 2;                              // TMEM columns of one weight part
    constexpr int WBUF_COLS = NPART * WPART_COLS;
    static_assert(2 * WBUF_COLS <= ACC_TMEM_OFF, "weight buffers");
    constexpr int ACC_COLS = 128;                                  // per accumulator set: main [0, NMAX) | correction [64, 128)
    static_assert(NPROD == 1 || NMAX == 64, "correction accumulator offset");
    constexpr uint32_t FMT = NPROD == 3 ? 0u /*F16*/ : tc::FMT_BF16;

    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned ring; the pad is added as an OFFSET so that the pointers keep the shared address space (an integer
    // round trip makes every access a generic LD/ST with 64-bit address arithmetic)
    unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float *agg_s = reinterpret_cast<float *>(ring + AGG_OFF);
    unsigned char *tail = ring + AGG_OFF + agg_bytes(p.B);
    Meta *meta_ring = reinterpret_cast<Meta *>(tail);
    Sched *sched = reinterpret_cast<Sched *>(tail + META_RING * sizeof(Meta));
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail + META_RING * sizeof(Meta) + SCHED_RING * sizeof(Sched));
    uint64_t *x_full = bars, *x_empty = bars + 3, *w_full = bars + 6, *w_empty = bars + 8, *acc_full = bars + 10,
             *acc_empty = bars + 12, *sched_full = bars + 14, *sched_empty = bars + 18;
    uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(bars + 22);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        // x_full: per gather thread one asynchronous arrival when its copies have landed (cp.async.mbarrier.arrive.noinc) and one
        // ordinary arrival that publishes the step's column metadata
        for (int s = 0; s < NUM_SLOTS; ++s) { mbar_init(&x_full[s], 2 * GATHER_THREADS); mbar_init(&x_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&w_full[b], 128); mbar_init(&w_empty[b], 1);
            mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8);
        }
        for (int r = 0; r < SCHED_RING; ++r) { mbar_init(&sched_full[r], 1); mbar_init(&sched_empty[r], NUM_CONSUMER_WARPS); }
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc<512>(tmem_base_smem);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_smem, 0);
    const int T = p.T;

    if (warp >= 4 && warp < 8) {
        tc::reg_dealloc<MID_REGS>();
        if (warp == MMA_WARP) {
            // ============================================ MMA ISSUER ============================================
            const bool leader = tc::elect_one();
            StepGen<NSEG> gen{sched, sched_full, sched_empty, T, NMAX, lane};
            uint32_t xs = 0, sg = 0, wl = 0, wl0 = 0;
            Trace tr = make_trace(p.trace, 1, leader);
            Step s;
            for (;;) {
                const int ev = gen.next(s);
                if (ev == 2) break;
                if (ev == 1) continue;
                if (s.first_sub && s.seg == 0) { wl0 = wl; wl += NSEG; }
                const uint32_t ab = sg & 1;
                tr.mark(10, xs);
                const uint32_t slot = xs % NUM_SLOTS;
                const uint32_t wli = wl0 + s.seg, wb = wli & 1;
                {   // lane 0: the epilogue has drained this accumulator set | lane 1: the rows have landed | lane 2: the weights have
                    const uint32_t addr = lane == 0 ? smem_u32(&acc_empty[ab]) : (lane == 1 ? smem_u32(&x_full[slot]) : smem_u32(&w_full[wb]));
                    const uint32_t parity = lane == 0 ? (((sg >> 1) & 1) ^ 1) : (lane == 1 ? ((xs / NUM_SLOTS) & 1) : ((wli >> 1) & 1));
                    const bool active = lane == 0 ? s.seg == 0 : (lane == 1 ? true : (lane == 2 && s.first_sub));
                    mbar_wait_lanes(addr, parity, active);
                }
                tc::fence_proxy_async_smem();          // rows written by cp.async (generic proxy) -> read by the MMA (async proxy)
                tr.mark(13, xs);
                tc::tc_fence_after_sync();
                const uint32_t n16 = (uint32_t)(s.n + 15) & ~15u;
                const uint32_t idesc = tc::make_instr_desc(FMT, 128, n16);
                const uint32_t slot_addr = smem_u32(ring + slot * SLOT_BYTES);
                const uint32_t d_main = tmem_base + ACC_TMEM_OFF + ab * ACC_COLS, d_corr = d_main + 64;
                const uint32_t a_base = tmem_base + wb * WBUF_COLS;
#pragma unroll
                for (int ks = 0; ks < K / 16; ++ks) {
                    const int kc = ks >> 2, kk = ks & 3;
                    const uint32_t acc = (s.seg == 0 && ks == 0) ? 0u : 1u;
                    const uint64_t x_hi = tc::make_smem_desc_sw128(slot_addr + kc * TILE_BYTES) + kk * 2;
                    if (NPROD == 3) {
                        const uint64_t x_lo = tc::make_smem_desc_sw128(slot_addr + (KCH + kc) * TILE_BYTES) + kk * 2;
                        if (leader) {
                            // x * w ~= hi*hi (main) + 2^-11 (hi*lo' + lo'*hi) (correction accumulator, scaled by 2^11)
                            mma_f16_ts(d_main, a_base + ks * 8, x_hi, idesc, acc);
                            mma_f16_ts(d_corr, a_base + ks * 8, x_lo, idesc, acc);
                            mma_f16_ts(d_corr, a_base + WPART_COLS + ks * 8, x_hi, idesc, 1u);
                        }
                    } else {
                        if (leader) mma_f16_ts(d_main, a_base + ks * 8, x_hi, idesc, acc);
                    }
                }
                if (leader) tc::mma_commit(&x_empty[slot]);
                if (leader && s.last_sub) tc::mma_commit(&w_empty[wb]);
                if (leader && s.seg == NSEG - 1) tc::mma_commit(&acc_full[ab]);
                __syncwarp();
                tr.mark(14, xs);
                ++xs;
                if (s.seg == NSEG - 1) ++sg;
            }
        } else if (warp == SCHED_WARP) {
            // ============================================ SCHEDULER ============================================
            // Publishes, a few blocks ahead, the group-offset row of each target block this CTA owns (static round robin).
            for (uint32_t i = 0;; ++i) {
                const uint32_t r = i % SCHED_RING;
                mbar_wait(&sched_empty[r], ((i / SCHED_RING) & 1) ^ 1);
                const long long blk = (long long)blockIdx.x + (long long)i * gridDim.x;
                const bool valid = blk < p.num_blocks;
                Sched *e = &sched[r];
                if (valid) {
                    for (int t = lane; t <= T; t += 32) e->off[t] = __ldg(p.group_off + blk * T + t);
                }
                if (lane == 0) e->blk = valid ? (int)blk : -1;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sched_full[r]);
                if (!valid) break;
            }
        } else {
            // ============================================ ROW GATHERERS ============================================
            // 64 threads copy the packed state rows of every (sub-group, segment) into a ring slot with 16-byte cp.async:
            // thread g moves piece q = g & 7 of every 128-byte chunk of rows (g >> 3) + 8 i.  Tile (part, kc) of a slot holds
            // columns [64 kc, 64 kc + 64) of that part, SWIZZLE_128B K-major -- the MMA's B operand.  Rows >= n are left as
            // they are: an accumulator column depends on its own B row only, and the epilogue never reads those columns.
            const int g = (int)threadIdx.x - GATHER_WARP0 * 32;
            const int q = g & 7, rsub = g >> 3;
            StepGen<NSEG> gen{sched, sched_full, sched_empty, T, NMAX, lane};
            uint32_t c_issue = 0, c_done = 0, sgc = 0;
            Trace tr = make_trace(p.trace, 0, g == 0);
            // Everything a step needs from global memory (row indices, the targets of its columns) is loaded ONE STEP AHEAD:
            // `fetch` only issues the loads, the copies of the current step are issued while they are in flight, `finish_meta`
            // consumes them afterwards.
            Step cur, nxt;
            bool has_nxt = false;
            int idx_nxt[NMAX / 8];
            int tl_a[NMAX / 64], tl_b[NMAX / 64];
            auto fetch = [&]() {                       // -> nxt, idx_nxt, tl_a / tl_b (raw loads only)
                int ev;
                do { ev = gen.next(nxt); } while (ev == 1);
                has_nxt = ev == 0;
                if (!has_nxt) return;
                if (nxt.seg == 0) {
#pragma unroll
                    for (int i = 0; i < NMAX / 8; ++i) {
                        const int r = rsub + 8 * i;
                        idx_nxt[i] = r < nxt.n ? __ldg(p.src_f + nxt.e + r) : -1;
                    }
#pragma unroll
                    for (int half = 0; half < NMAX / 64; ++half) {
                        const int c = g + 64 * half;
                        tl_a[half] = c < nxt.n ? (int)__ldg(p.tl_f + nxt.e + c) : -1;
                        tl_b[half] = c + 1 < nxt.n ? (int)__ldg(p.tl_f + nxt.e + c + 1) : -2;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NMAX / 8; ++i) {
                        const int r = rsub + 8 * i;
                        idx_nxt[i] = r < nxt.n ? nxt.blk * p.B + (int)__ldg(p.tl_f + nxt.e + r) : -1;
                    }
                }
            };
            auto finish_meta = [&](const Step &st) {   // column metadata of step `st` for the epilogue (columns g and g + 64),
                if (st.seg != 0) return;               // from the tl_a / tl_b loads issued one whole step earlier
                Meta *m = &meta_ring[sgc % META_RING];
#pragma unroll
                for (int half = 0; half < NMAX / 64; ++half) {
                    const int c = g + 64 * half;
                    const bool valid = tl_a[half] >= 0;
                    const bool end = valid && tl_b[half] != tl_a[half];      // last column (tl_b = -2) or the target changes
                    m->tloff[c] = valid ? tl_a[half] * (kD * 4) : 0;
                    const uint32_t word = __ballot_sync(0xffffffffu, end);
                    const uint32_t low = __ballot_sync(0xffffffffu, valid && tl_a[half] < (p.B >> 1));
                    if (lane == 0) { m->endmask[c >> 5] = word; m->lowmask[c >> 5] = low; }
                }
                if (g == 0) m->n = st.n;
                ++sgc;
            };
            auto issue_next = [&]() {                  // issue the copies of `nxt`, prefetch the step after it
                cur = nxt;
                int idx[NMAX / 8];
#pragma unroll
                for (int i = 0; i < NMAX / 8; ++i) idx[i] = idx_nxt[i];
                finish_meta(cur);                      // consumes the loads of the PREVIOUS call
                fetch();                               // loads for the following step: in flight during the copies below
                const uint32_t slot = c_issue % NUM_SLOTS;
                tr.mark(1, c_issue);
                mbar_wait(&x_empty[slot], ((c_issue / NUM_SLOTS) & 1) ^ 1);
                tr.mark(2, c_issue);
                const unsigned char *rows = cur.seg == 0 ? p.src_rows : p.tgt_rows;
                const uint32_t sbase = smem_u32(ring + slot * SLOT_BYTES) + swz(rsub, q);
#pragma unroll
                for (int i = 0; i < NMAX / 8; ++i) {
                    if (idx[i] >= 0) {
                        const unsigned char *src = rows + (size_t)idx[i] * ROW_BYTES + q * 16;
#pragma unroll
                        for (int ti = 0; ti < NT; ++ti)
                            cp_async16(sbase + ti * TILE_BYTES + i * 1024, src + (ti / KCH) * (K * 2) + (ti % KCH) * 128, 16);
                    }
                }
                tr.mark(3, c_issue);
                ++c_issue;
            };
            fetch();
            bool more = has_nxt;
            // Completion is signalled by the copies themselves (cp.async.mbarrier.arrive.noinc): the gatherers never wait for data, only
            // for a free slot, so a step's arrival is not held back by the issue of the next one (with cp.async.wait_group it was:
            // measured 2800 cycles from slot grant to arrival, most of it the next step's slot wait).  The MMA warp makes the landed
            // bytes visible to the tensor core's async proxy with fence.proxy.async after its wait.
            while (more) {
                const uint32_t slot = c_issue % NUM_SLOTS;
                issue_next();
                more = has_nxt;
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&x_full[slot])) : "memory");
                tr.mark(4, c_done);
                mbar_arrive(&x_full[slot]);            // release: this thread's metadata stores of the step
                ++c_done;
            }
            cp_async_wait<0>();
        }
    } else if (warp >= 12) {
        // ============================================ WEIGHT LOADERS ============================================
        // Thread d owns TMEM lane d = row d of W_t.  The packed weights are laid out so that a warp-wide 16-byte load is one
        // contiguous 512-byte burst: wpack[(((t * NSEG + seg) * NPART + part) * (K / 8) + c4) * 128 + d] = columns 4 c4 .. 4 c4 + 3.
        // All loads of a (type, segment) are in flight before the buffer's release is awaited.
        tc::reg_dealloc<W_REGS>();
        const int d = (warp & 3) * 32 + lane;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        StepGen<NSEG> gen{sched, sched_full, sched_empty, T, NMAX, lane};
        uint32_t wl = 0;
        Trace tr = make_trace(p.trace, 4, (warp & 3) == 0 && lane == 0);
        Step s;
        for (;;) {
            const int ev = gen.next(s);
            if (ev == 2) break;
            if (ev == 1 || !s.first_sub) continue;
            tr.mark(30, wl);
            // 64 TMEM columns (16 16-byte loads, 64 registers) per round: 512 threads leave 128 registers per thread, so the 128
            // columns of an fp32 (hi | lo') weight buffer go in two rounds; the first round's loads are issued before the
            // buffer's release is awaited (the loaders run up to two groups ahead of the MMAs)
            constexpr int ROUND = WBUF_COLS < 64 ? WBUF_COLS : 64, NROUNDS = WBUF_COLS / ROUND;
            const uint4 *src = p.wpack + ((size_t)(s.t * NSEG + s.seg) * NPART * (K / 8)) * 128 + d;
            uint32_t w[ROUND];
            auto load_round = [&](int r) {
#pragma unroll
                for (int j = 0; j < ROUND / 4; ++j) {
                    const uint4 v = ldg_nc_u4(src + (size_t)(r * (ROUND / 4) + j) * 128);
                    w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
                }
            };
            load_round(0);
            const uint32_t wb = wl & 1;
            tr.mark(31, wl);
            mbar_wait(&w_empty[wb], ((wl >> 1) & 1) ^ 1);        // the MMAs that read this buffer two loads ago are done
            tr.mark(32, wl);
            tc::tc_fence_after_sync();
#pragma unroll
            for (int r = 0; r < NROUNDS; ++r) {
#pragma unroll
                for (int j = 0; j < ROUND / 32; ++j) {
                    uint32_t chunk[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) chunk[i] = w[32 * j + i];
                    tmem_st_32cols_u32(tmem_lane + wb * WBUF_COLS + r * ROUND + 32 * j, chunk);
                }
                if (r + 1 < NROUNDS) {
                    tc::tmem_st_wait();                            // the stores have read their registers
                    load_round(r + 1);
                }
            }
            tc::tmem_st_wait();
            tc::tc_fence_before_sync();
            mbar_arrive(&w_full[wb]);
            tr.mark(33, wl);
            ++wl;
        }
    } else {
        // ============================================ EPILOGUE ============================================
        // Thread d owns message feature d = TMEM lane d and column d of agg_s.  For every accumulator column (edge) in plan
        // order: value = main (+ 2^-11 correction); at the first edge of a (target, type) segment the running value is
        // (re)loaded from agg_s[target][d], at the last one it is stored back -- a target's messages are accumulated one by one
        // in the reference's order, across types and sub-groups.
        // TWO warpgroups (a single warp per scheduler is latency-bound: measured IPC 0.17): group 0 takes the columns whose
        // target lies in the lower half of the block, group 1 the upper half.  Edges are sorted by target, so each group owns a
        // contiguous column range of every sub-group (split = number of lower-half columns) and the two never touch the same
        // agg_s row -- no synchronisation between them except at the block's write-out.
        tc::reg_alloc<EPI_REGS>();                 // granted once warps 4-7 and 12-15 have released theirs
        const int eg = warp >> 3, ew = warp & 3;             // warps 0-3: group 0, warps 8-11: group 1
        const int d = ew * 32 + lane;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(ew * 32) << 16) + ACC_TMEM_OFF;
        const uint32_t aggcol_s = smem_u32(agg_s + d);      // shared-space address of agg_s[0][d]
        const float IDENT = red_identity<RED>();
        StepGen<NSEG> gen{sched, sched_full, sched_empty, T, NMAX, lane};
        const int row_lo = eg == 0 ? 0 : (p.B >> 1), row_hi = eg == 0 ? (p.B >> 1) : p.B;     // rows this group initialises
        for (int r = row_lo; r < row_hi; ++r) agg_s[r * kD + d] = IDENT;
        uint32_t sg = 0;
        float acc = IDENT;
        Trace tr = make_trace(p.trace, 2 + eg, ew == 0 && lane == 0);
        Step s;
        for (;;) {
            const int ev = gen.next(s);
            if (ev == 2) break;
            if (ev == 0) {
                if (s.seg != NSEG - 1) continue;
                const uint32_t ab = sg & 1;
                tr.mark(20, sg);
                mbar_wait(&acc_full[ab], (sg >> 1) & 1);
                tr.mark(21, sg);
                tc::tc_fence_after_sync();
                const Meta *m = &meta_ring[sg % META_RING];
                const int n = s.n;
                int split = __popc(m->lowmask[0]) + __popc(m->lowmask[1]);
                if (NMAX > 64) split += __popc(m->lowmask[2]) + __popc(m->lowmask[3]);
                const int c_lo = eg == 0 ? 0 : split, c_hi = eg == 0 ? split : n;       // this group's columns
                constexpr int W = 16;
                for (int c0 = c_lo & ~15; c0 < c_hi; c0 += 16) {
                    uint32_t vm[W], vc[W];
                    tc::tmem_ld_16cols_async(tmem_lane + ab * ACC_COLS + c0, vm);
                    if (NPROD == 3) tc::tmem_ld_16cols_async(tmem_lane + ab * ACC_COLS + 64 + c0, vc);
                    uint32_t addr[W];
#pragma unroll
                    for (int j = 0; j < W / 4; ++j) {
                        const int4 o = *reinterpret_cast<const int4 *>(&m->tloff[c0 + 4 * j]);
                        addr[4 * j] = aggcol_s + o.x; addr[4 * j + 1] = aggcol_s + o.y;
                        addr[4 * j + 2] = aggcol_s + o.z; addr[4 * j + 3] = aggcol_s + o.w;
                    }
                    const uint32_t endw = (m->endmask[c0 >> 5] >> (c0 & 31)) & 0xFFFFu;
                    // the column before this batch ended a segment (or the batch opens the sub-group: always reload)
                    const uint32_t prev_end = c0 == 0 ? 1u : (m->endmask[(c0 - 1) >> 5] >> ((c0 - 1) & 31)) & 1u;
                    const uint32_t startw = (endw << 1) | prev_end;
                    // columns of the batch that belong to this group: [max(c_lo, c0), min(c_hi, c0 + 16))
                    const int first = c_lo > c0 ? c_lo - c0 : 0, last = c_hi - c0 < W ? c_hi - c0 : W;
                    const uint32_t storew = endw & (0xFFFFu << first) & (0xFFFFu >> (W - last));
                    float pre[W];
#pragma unroll
                    for (int c = 0; c < W; ++c) pre[c] = lds_f32(addr[c]);
                    tc::tmem_ld_wait();
                    // t[c] = op(pre[c], v[c]) for every column (independent); a column that CONTINUES a segment (rare: most
                    // (target, type) segments hold one edge) then overwrites it with op(t[c-1], v[c]) -- a predicated op, in
                    // column order, so a target's messages are still combined one by one in plan order.  Columns of the other
                    // group are computed but never stored; this group's first column always starts a segment.
                    float t[W];
                    constexpr bool ADD = RED == PTGNN_REDUCE_SUM || RED == PTGNN_REDUCE_MEAN;
#pragma unroll
                    for (int c = 0; c < W; c += 2) {       // two columns per instruction: Blackwell's packed fp32 FMA / ADD (same rounding)
                        float2 v = make_float2(__uint_as_float(vm[c]), __uint_as_float(vm[c + 1]));
                        if (NPROD == 3) {
                            v = __ffma2_rn(make_float2(__uint_as_float(vc[c]), __uint_as_float(vc[c + 1])), make_float2(1.0f / 2048.0f, 1.0f / 2048.0f), v);
                        } else {                                   // the autocast Linear's bf16 output
                            v.x = __bfloat162float(__float2bfloat16_rn(v.x));
                            v.y = __bfloat162float(__float2bfloat16_rn(v.y));
                        }
                        vm[c] = __float_as_uint(v.x); vm[c + 1] = __float_as_uint(v.y);
                        if (ADD) {
                            const float2 s2 = __fadd2_rn(make_float2(pre[c], pre[c + 1]), v);
                            t[c] = s2.x; t[c + 1] = s2.y;
                        } else {
                            t[c] = red_op<RED>(pre[c], v.x); t[c + 1] = red_op<RED>(pre[c + 1], v.y);
                        }
                    }
                    continue_segment<RED>(t[0], acc, __uint_as_float(vm[0]), startw & 1u);
#pragma unroll
                    for (int c = 1; c < W; ++c) continue_segment<RED>(t[c], t[c - 1], __uint_as_float(vm[c]), startw & (1u << c));
                    acc = t[W - 1];
#pragma unroll
                    for (int c = 0; c < W; ++c) sts_f32_if(addr[c], t[c], storew & (1u << c));
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ab]);
                tr.mark(22, sg);
                ++sg;
                continue;
            }
            // ---- block finished.  Each group writes out ITS half of the rows as soon as its own four warps are done (no waiting for
            // the other group).  The row loop is specialised at compile time on the output format and on "plain sum" (no mean /
            // max fix-up / activation / LayerNorm): the generic version executed ~180 instructions per row, 13,600 cycles per block.
            tr.mark(23, sg);
            if (eg == 0) named_bar_sync(EPI_BAR_ID, 128); else named_bar_sync(EPI_BAR_ID + 1, 128);
            write_out_block<RED>(&p, smem_u32(agg_s), s.blk * p.B, row_lo, row_hi, ew, lane);
            if (eg == 0) named_bar_sync(EPI_BAR_ID, 128); else named_bar_sync(EPI_BAR_ID + 1, 128);
            tr.mark(24, sg);
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================================
// packing kernels
// =====================================================================================================================
// fp32 [rows, K] -> rows of 2K fp16: hi[K] | lo'[K].  One thread per 8 consecutive elements (32 bytes in, 2 x 16 bytes out).
__global__ void __launch_bounds__(256) pack_states_kernel(const float *__restrict__ h, long long rows, int K,
                                                          uint4 *__restrict__ out, int32_t *__restrict__ status) {
    const long long total = rows * (K / 8);
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (K / 8);
        const int c8 = (int)(i - row * (K / 8));
        const float4 a = ld_stream_f4(reinterpret_cast<const float4 *>(h + row * K + c8 * 8));
        const float4 b = ld_stream_f4(reinterpret_cast<const float4 *>(h + row * K + c8 * 8 + 4));
        const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bad |= !split_f16(x[j], hi[j], lo[j]);
        uint4 vh, vl;
        vh.x = pack_h2(hi[0], hi[1]); vh.y = pack_h2(hi[2], hi[3]); vh.z = pack_h2(hi[4], hi[5]); vh.w = pack_h2(hi[6], hi[7]);
        vl.x = pack_h2(lo[0], lo[1]); vl.y = pack_h2(lo[2], lo[3]); vl.z = pack_h2(lo[4], lo[5]); vl.w = pack_h2(lo[6], lo[7]);
        uint4 *dst = out + row * (K / 4);       // row = 4K bytes = K/4 uint4: hi part first (K/8 uint4), then lo'
        dst[c8] = vh;
        dst[K / 8 + c8] = vl;
    }
    if (bad && status != nullptr) *reinterpret_cast<volatile int32_t *>(status) = 1;
}

struct WeightSrc {
    const float *w[PTGNN_MAX_EDGE_TYPES];
};
// out[(((t * nseg + seg) * npart + part) * (K / 8) + c4) * 128 + d] = 8 elements k = seg K + 8 c4 .. + 7 of row d
template <int NPROD>
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ WeightSrc src, int num_types, int K, int nseg,
                                                           uint4 *__restrict__ out, int32_t *__restrict__ status) {
    constexpr int NPART = NPROD == 3 ? 2 : 1;
    const int per_mat = nseg * (K / 8) * 128;             // (seg, c4, d) triples per type
    const long long total = (long long)num_types * per_mat;
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / per_mat);
        int r = (int)(i - (long long)t * per_mat);
        const int seg = r / ((K / 8) * 128);
        r -= seg * (K / 8) * 128;
        const int c4 = r / 128, d = r % 128;
        const float *row = src.w[t] + (size_t)d * (nseg * K) + seg * K + c4 * 8;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = row[j];
        const size_t base = ((size_t)(t * nseg + seg) * NPART) * (K / 8) * 128;
        if (NPROD == 3) {
            __half hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bad |= !split_f16(x[j], hi[j], lo[j]);
            uint4 vh, vl;
            vh.x = pack_h2(hi[0], hi[1]); vh.y = pack_h2(hi[2], hi[3]); vh.z = pack_h2(hi[4], hi[5]); vh.w = pack_h2(hi[6], hi[7]);
            vl.x = pack_h2(lo[0], lo[1]); vl.y = pack_h2(lo[2], lo[3]); vl.z = pack_h2(lo[4], lo[5]); vl.w = pack_h2(lo[6], lo[7]);
            out[base + (size_t)c4 * 128 + d] = vh;
            out[base + (size_t)(K / 8 + c4) * 128 + d] = vl;
        } else {
            uint4 v;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(x[0], x[1]), p1 = __floats2bfloat162_rn(x[2], x[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(x[4], x[5]), p3 = __floats2bfloat162_rn(x[6], x[7]);
            v.x = *reinterpret_cast<uint32_t *>(&p0); v.y = *reinterpret_cast<uint32_t *>(&p1);
            v.z = *reinterpret_cast<uint32_t *>(&p2); v.w = *reinterpret_cast<uint32_t *>(&p3);
            out[base + (size_t)c4 * 128 + d] = v;
        }
    }
    if (bad && status != nullptr) *reinterpret_cast<volatile int32_t *>(status + 1) = 1;
}

// =====================================================================================================================
// host side
// =====================================================================================================================
bool supported(int nprod, int K, int D, int use_target) {
    (void)use_target;
    if (D != kD) return false;
    if (nprod == 3) return K == 64 || K == 128;
    if (nprod == 1) return K == 64 || K == 128 || K == 256;
    return false;
}
size_t packed_weight_bytes(int nprod, int num_types, int K, int use_target) {
    const int nseg = use_target ? 2 : 1, npart = nprod == 3 ? 2 : 1;
    return ws_slice((size_t)num_types * nseg * npart * (K / 8) * 128, 16);
}
size_t packed_state_bytes(int nprod, int64_t rows, int K) {
    return nprod == 3 ? ws_slice((size_t)rows * K * 4 + 16, 1) : 0;
}
int recommended_block_targets(int64_t num_nodes) {
    // the largest B <= kMaxBlockTargets (multiple of 8) for which the block count is a whole number of 148-CTA waves or less
    if (num_nodes <= 0) return kMaxBlockTargets;
    const int64_t waves = ceil_div(num_nodes, (int64_t)kMaxBlockTargets * 148);
    int64_t B = ceil_div(num_nodes, waves * 148);
    B = (B + 7) / 8 * 8;
    if (B < 8) B = 8;
    if (B > kMaxBlockTargets) B = kMaxBlockTargets;
    return (int)B;
}

int pack_weights(int nprod, int num_types, int K, int use_target, const float *const *weights, void *packed, int32_t *status,
                 cudaStream_t st) {
    WeightSrc src{};
    for (int t = 0; t < num_types; ++t) src.w[t] = weights[t];
    const int nseg = use_target ? 2 : 1;
    {
        TimedScope timed__(PTGNN_KERNEL_PACK, st);
        if (nprod == 3) pack_weights_kernel<3><<<148, 256, 0, st>>>(src, num_types, K, nseg, static_cast<uint4 *>(packed), status);
        else pack_weights_kernel<1><<<148, 256, 0, st>>>(src, num_types, K, nseg, static_cast<uint4 *>(packed), status);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

int pack_states(const float *h, int64_t rows, int K, void *packed, int32_t *status, cudaStream_t st) {
    if (rows <= 0) return PTGNN_OK;
    const int64_t items = rows * (K / 8);
    const unsigned grid = (unsigned)(ceil_div(items, 256) < 148 * 8 ? ceil_div(items, 256) : 148 * 8);
    {
        TimedScope timed__(PTGNN_KERNEL_PACK, st);
        pack_states_kernel<<<grid, 256, 0, st>>>(h, (long long)rows, K, static_cast<uint4 *>(packed), status);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

template <int NPROD, int K, int NSEG, int RED>
static int launch_one(const Params &p, cudaStream_t st) {
    auto kernel = fused_aggregate_kernel<NPROD, K, NSEG, RED>;
    const int smem = smem_bytes(p.B);
    // per launch, not once per process: the attribute belongs to the current device's context
    PTGNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    const int grid = p.num_blocks < sms ? p.num_blocks : sms;
    {
        TimedScope timed__(PTGNN_KERNEL_MESSAGE, st);
        kernel<<<grid, NUM_THREADS, smem, st>>>(p);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}
template <int NPROD, int K, int NSEG>
static int launch_red(const Params &p, cudaStream_t st) {
    switch (p.reduce) {
        case PTGNN_REDUCE_SUM: return launch_one<NPROD, K, NSEG, PTGNN_REDUCE_SUM>(p, st);
        case PTGNN_REDUCE_MEAN: return launch_one<NPROD, K, NSEG, PTGNN_REDUCE_MEAN>(p, st);
        case PTGNN_REDUCE_MAX: return launch_one<NPROD, K, NSEG, PTGNN_REDUCE_MAX>(p, st);
        default: return launch_one<NPROD, K, NSEG, PTGNN_REDUCE_MIN>(p, st);
    }
}
template <int NPROD, int K>
static int launch_seg(const Params &p, int use_target, cudaStream_t st) {
    return use_target ? launch_red<NPROD, K, 2>(p, st) : launch_red<NPROD, K, 1>(p, st);
}

static unsigned long long *g_trace_dev = nullptr;
static unsigned long long *trace_buffer() {
    static int want = -1;
    if (want < 0) { const char *e = getenv("PTGNN_FUSED_TRACE"); want = (e && e[0] == '1') ? 1 : 0; }
    if (!want) return nullptr;
    if (!g_trace_dev && cudaMalloc(&g_trace_dev, 5 * 2048 * 8) != cudaSuccess) return nullptr;
    cudaMemset(g_trace_dev, 0, 5 * 2048 * 8);
    return g_trace_dev;
}

int aggregate(const AggregateArgs &a, cudaStream_t st) {
    PTGNN_CHECK_ARG(supported(a.nprod, a.K, kD, a.use_target), "fused aggregate: unsupported nprod=%d K=%d", a.nprod, a.K);
    PTGNN_CHECK_ARG(a.block_targets >= 8 && a.block_targets <= kMaxBlockTargets, "fused aggregate: block_targets=%d out of [8, %d]",
                    a.block_targets, kMaxBlockTargets);
    PTGNN_CHECK_ARG(a.num_types > 0 && a.num_types <= PTGNN_MAX_EDGE_TYPES, "fused aggregate: bad num_types=%d", a.num_types);
    PTGNN_CHECK_ARG(a.reduce >= PTGNN_REDUCE_SUM && a.reduce <= PTGNN_REDUCE_MIN, "fused aggregate: bad reduce %d", a.reduce);
    if (a.num_nodes <= 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(a.src_rows && a.group_off && a.packed_weights && a.out, "fused aggregate: null pointer");
    PTGNN_CHECK_ARG(!a.use_target || a.tgt_rows, "fused aggregate: null target rows");
    PTGNN_CHECK_ARG(a.reduce != PTGNN_REDUCE_MEAN || a.row_ptr, "fused aggregate: mean needs row_ptr");
    Params p{};
    p.src_rows = static_cast<const unsigned char *>(a.src_rows);
    p.tgt_rows = static_cast<const unsigned char *>(a.tgt_rows);
    p.wpack = static_cast<const uint4 *>(a.packed_weights);
    p.group_off = a.group_off; p.src_f = a.src_f; p.tl_f = a.tl_f; p.row_ptr = a.row_ptr;
    p.out = a.out; p.num_nodes = (int)a.num_nodes; p.B = a.block_targets;
    p.num_blocks = (int)ceil_div(a.num_nodes, a.block_targets);
    p.T = a.num_types; p.reduce = a.reduce; p.out_mode = a.out_mode; p.status = a.status; p.epi = a.epi;
    p.trace = trace_buffer();
#ifdef PTGNN_FUSED_QUICK   // compile-time experiments (ptxas -v / SASS of the benchmarked instances only); never defined in the build
    return a.nprod == 3 ? launch_one<3, 128, 1, PTGNN_REDUCE_SUM>(p, st) : launch_one<1, 128, 1, PTGNN_REDUCE_SUM>(p, st);
#else
    if (a.nprod == 3) {
        if (a.K == 64) return launch_seg<3, 64>(p, a.use_target, st);
        return launch_seg<3, 128>(p, a.use_target, st);
    }
    if (a.K == 64) return launch_seg<1, 64>(p, a.use_target, st);
    if (a.K == 128) return launch_seg<1, 128>(p, a.use_target, st);
    return launch_seg<1, 256>(p, a.use_target, st);
#endif
}

}  // namespace fused
}  // namespace ptgnn

// debug only (not part of the public header): copies the last fused-kernel timeline (5 roles x 2048 entries) to `out`; 0 if tracing is off
extern "C" int ptgnn_b200_debug_fused_trace(unsigned long long *out) {
    if (!ptgnn::fused::g_trace_dev) return 0;
    cudaDeviceSynchronize();
    cudaMemcpy(out, ptgnn::fused::g_trace_dev, 5 * 2048 * 8, cudaMemcpyDeviceToHost);
    return 1;
}
