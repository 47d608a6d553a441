// nn.GRUCell update, weights-stationary:  h' = GRUCell(agg, h)   (reference gatedmessagepassing.py:69)
//
// Why a second GRU kernel.  The round-1 pipeline (tc_pipeline.cuh, GruPolicy) walks (row tile, 32-hidden-unit block) tiles in
// row-major order: every tile streams its gate-weight block again and every row tile is fetched once per block -- ~2.4 GB of
// L2 -> SM traffic per launch at config 2, which is what bounds it (the L2 fabric delivers ~12 TB/s chip-wide), not the tensor
// pipe.  Here a CTA keeps ONE hidden-unit block for its whole life: its gate weights are loaded into shared memory once and
// stay (B operand of every MMA), and only the node rows stream through a TMA ring (A operand).  L2 traffic drops to the row
// tiles alone.
//
// Arithmetic (same two modes as the fused aggregation kernel, fused_mp.cuh):
//   NPROD = 3  fp32-exact "3xFP16": operands are (hi, lo') fp16 pairs -- the aggregate arrives in that form from the fused
//              kernel's write-out, the states from pack_states, the weights are packed once per parameter version;
//              hi*hi -> main accumulator, hi*lo' + lo'*hi -> correction accumulator (scaled by 2^11), fp32 h for the blend.
//   NPROD = 1  bf16 operands (the reference under torch.autocast), fp32 accumulation and gate math.
//
// Accumulator columns of a tile (128 rows x 32 hidden units): [0,32) i_n | [32,64) r | [64,96) z | [96,128) h_n.
//   state segment first: P2 = [0; W_hr; W_hz; W_hn] (128 rows) -- its first K-step runs N = 128 and overwrites all columns
//   (the zero block initialises i_n), later K-steps N = 96 over rows 32..127; then the aggregate segment:
//   P1 = [W_in; W_ir; W_iz] (96 rows) accumulates into columns 0..95.
// Roles: warp 0 TMA producer | warp 1 MMA issuer | warps 4-11 epilogue, two sets of four on alternate tiles (TMEM: two
// accumulator sets x (main 128 | correction 128) columns).
#include "gru_ws.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_pipeline.cuh"

namespace ptgnn {
namespace gruws {

using tc::mbar_arrive;
using tc::mbar_init;
using tc::mbar_wait;

constexpr int NUM_THREADS = 12 * 32;
constexpr int TILE_M = 128;
constexpr int IO_REGS = 64, EPI_REGS = 216;          // (64 + 216 + 216) * 128 = 63488

struct Params {
    CUtensorMap map_agg, map_h;          // A: [N, NPART * K] 16-bit, box {64, 128}
    CUtensorMap map_p1, map_p2;          // B: [NPART * n_jb * 96, D] and [NPART * n_jb * 128, H] 16-bit, boxes {64, 96} / {64, 128}
    const float *h32;                    // NPROD 3: fp32 states for the blend
    const __nv_bfloat16 *h16;            // NPROD 1
    const float4 *bias4;                 // (b_ir + b_hr, b_iz + b_hz, b_in, b_hn) per hidden unit
    void *out;                           // fp32 [N, H] (NPROD 3) or bf16 (NPROD 1)
    __half *out_packed;                  // optional (NPROD 3): the new states also as fp16 (hi | lo') rows of 2H halfs -- what the next
                                         // layer's fused aggregation and GRU take as MMA operands (saves its pack_states pass)
    int32_t *status;                     // optional: status[0] = 1 if a new state is outside the fp16 range (out_packed only)
    int num_nodes, H, D, n_jb, n_rb;
};

template <int NPROD> struct Geometry {
    static constexpr int NPART = NPROD == 3 ? 2 : 1;
    static constexpr int A_TILE = TILE_M * 128;                 // 16 KB: 128 rows x 64 16-bit elements
    static constexpr int SLOT_BYTES = NPART * A_TILE;
    static constexpr int NUM_SLOTS = NPROD == 3 ? 3 : 6;
    static constexpr int RING_BYTES = NUM_SLOTS * SLOT_BYTES;   // 96 KB
    static constexpr int P1_TILE = 96 * 128, P2_TILE = 128 * 128;
    __host__ __device__ static constexpr int b_bytes(int H, int D) { return NPART * ((D / 64) * P1_TILE + (H / 64) * P2_TILE); }
    __host__ __device__ static constexpr int smem_bytes(int H, int D) { return 1024 + RING_BYTES + b_bytes(H, D) + H * 16 + 256; }
};

template <int NPROD>
__global__ void __launch_bounds__(NUM_THREADS, 1) gru_ws_kernel(const __grid_constant__ Params p) {
    using G = Geometry<NPROD>;
    constexpr int NPART = G::NPART;
    constexpr uint32_t FMT = NPROD == 3 ? 0u /*F16*/ : tc::FMT_BF16;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *bres = ring + G::RING_BYTES;                                   // resident weights: P2 tiles, then P1 tiles
    const int kch_h = p.H / 64, kch_d = p.D / 64;
    unsigned char *p1res = bres + NPART * kch_h * G::P2_TILE;
    float4 *bias_s = reinterpret_cast<float4 *>(bres + G::b_bytes(p.H, p.D));
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(bias_s) + p.H * 16);
    uint64_t *a_full = bars, *a_empty = bars + G::NUM_SLOTS, *acc_full = bars + 2 * G::NUM_SLOTS, *acc_empty = bars + 2 * G::NUM_SLOTS + 2;
    uint64_t *b_full = bars + 2 * G::NUM_SLOTS + 4;
    uint32_t *tmem_base_smem = reinterpret_cast<uint32_t *>(bars + 2 * G::NUM_SLOTS + 5);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < G::NUM_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        mbar_init(b_full, 1);
        tc::mbar_init_fence();
    }
    if (warp == 2) tc::tmem_alloc<512>(tmem_base_smem);
    for (int j = threadIdx.x; j < p.H; j += NUM_THREADS) bias_s[j] = p.bias4[j];
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_smem, 0);

    // this CTA's hidden-unit block and its row tiles
    const int jb = blockIdx.x % p.n_jb;
    const int rb0 = blockIdx.x / p.n_jb, rb_stride = gridDim.x / p.n_jb;
    const int chunks_per_tile = kch_h + kch_d;

    if (warp < 4) {
        tc::reg_dealloc<IO_REGS>();
        if (warp == 0) {
            // ============================================ TMA PRODUCER ============================================
            const bool leader = tc::elect_one();
            if (leader) {   // the resident weights, once
                tc::mbar_expect_tx(b_full, (uint32_t)G::b_bytes(p.H, p.D));
                for (int part = 0; part < NPART; ++part) {
                    for (int kc = 0; kc < kch_h; ++kc)
                        tc::tma_load_2d(bres + (part * kch_h + kc) * G::P2_TILE, &p.map_p2, kc * 64, (part * p.n_jb + jb) * 128, b_full);
                    for (int kc = 0; kc < kch_d; ++kc)
                        tc::tma_load_2d(p1res + (part * kch_d + kc) * G::P1_TILE, &p.map_p1, kc * 64, (part * p.n_jb + jb) * 96, b_full);
                }
            }
            __syncwarp();
            uint32_t c = 0;
            for (int rb = rb0; rb < p.n_rb; rb += rb_stride) {
                for (int ch = 0; ch < chunks_per_tile; ++ch, ++c) {
                    const uint32_t slot = c % G::NUM_SLOTS;
                    mbar_wait(&a_empty[slot], ((c / G::NUM_SLOTS) & 1) ^ 1);
                    const bool state_seg = ch < kch_h;
                    const CUtensorMap *map = state_seg ? &p.map_h : &p.map_agg;
                    const int kc = state_seg ? ch : ch - kch_h, K = state_seg ? p.H : p.D;
                    unsigned char *dst = ring + slot * G::SLOT_BYTES;
                    if (leader) {
                        tc::mbar_expect_tx(&a_full[slot], (uint32_t)G::SLOT_BYTES);
                        for (int part = 0; part < NPART; ++part)
                            tc::tma_load_2d(dst + part * G::A_TILE, map, part * K + kc * 64, rb * TILE_M, &a_full[slot]);
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 1) {
            // ============================================ MMA ISSUER ============================================
            const bool leader = tc::elect_one();
            mbar_wait(b_full, 0);
            tc::tc_fence_after_sync();
            const uint32_t idesc128 = tc::make_instr_desc(FMT, TILE_M, 128), idesc96 = tc::make_instr_desc(FMT, TILE_M, 96);
            const uint32_t b2_addr = smem_u32(bres), b1_addr = smem_u32(p1res);
            uint32_t c = 0, tcount = 0;
            for (int rb = rb0; rb < p.n_rb; rb += rb_stride, ++tcount) {
                const uint32_t set = tcount & 1;
                mbar_wait(&acc_empty[set], ((tcount >> 1) & 1) ^ 1);
                tc::tc_fence_after_sync();
                const uint32_t d_main = tmem_base + set * 256, d_corr = d_main + 128;
                for (int ch = 0; ch < chunks_per_tile; ++ch, ++c) {
                    const uint32_t slot = c % G::NUM_SLOTS;
                    mbar_wait(&a_full[slot], (c / G::NUM_SLOTS) & 1);
                    tc::tc_fence_after_sync();
                    const bool state_seg = ch < kch_h;
                    const int kc = state_seg ? ch : ch - kch_h;
                    const uint32_t a_addr = smem_u32(ring + slot * G::SLOT_BYTES);
                    const uint64_t a_hi = tc::make_smem_desc_sw128(a_addr), a_lo = tc::make_smem_desc_sw128(a_addr + G::A_TILE);
                    // B tile of this chunk: state segment = P2 (128 rows), aggregate segment = P1 (96 rows)
                    const uint32_t b_hi_addr = state_seg ? b2_addr + kc * G::P2_TILE : b1_addr + kc * G::P1_TILE;
                    const uint32_t b_lo_addr = state_seg ? b2_addr + (kch_h + kc) * G::P2_TILE : b1_addr + (kch_d + kc) * G::P1_TILE;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const bool first = ch == 0 && ks == 0;            // overwrites every accumulator column (zero block -> i_n)
                        const uint32_t row_off = (state_seg && !first) ? 32u * 128u : 0u;   // skip P2's zero block after the first step
                        const uint32_t col_off = (state_seg && !first) ? 32u : 0u;
                        const uint32_t idesc = first ? idesc128 : idesc96;
                        const uint64_t b_hi = tc::make_smem_desc_sw128(b_hi_addr + row_off) + ks * 2;
                        const uint32_t acc = first ? 0u : 1u;
                        if (NPROD == 3) {
                            const uint64_t b_lo = tc::make_smem_desc_sw128(b_lo_addr + row_off) + ks * 2;
                            if (leader) {
                                tc::mma_bf16_ss(d_main + col_off, a_hi + ks * 2, b_hi, idesc, acc);
                                tc::mma_bf16_ss(d_corr + col_off, a_hi + ks * 2, b_lo, idesc, acc);
                                tc::mma_bf16_ss(d_corr + col_off, a_lo + ks * 2, b_hi, idesc, 1u);
                            }
                        } else {
                            if (leader) tc::mma_bf16_ss(d_main + col_off, a_hi + ks * 2, b_hi, idesc, acc);
                        }
                    }
                    if (leader) tc::mma_commit(&a_empty[slot]);
                    __syncwarp();
                }
                if (leader) tc::mma_commit(&acc_full[set]);
                __syncwarp();
            }
        }
    } else {
        // ============================================ EPILOGUE ============================================
        // Set s (four warps, one per TMEM lane quarter) takes tiles s, s + 2, ...: a lane owns one node row and, in two passes,
        // 16 hidden units each: drain main (+ 2^-11 correction), gate math, blend with h, store 64 (32) contiguous bytes.
        tc::reg_alloc<EPI_REGS>();
        const int ew = warp - 4, quarter = warp & 3, set = ew >> 2;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + set * 256;
        const int rb_step = 2 * rb_stride;
        struct Pre { long long off; float4 f[8]; };      // fp32: 2 passes x 16 floats; bf16: 2 passes x 32 bytes in f[0..3]
        auto prefetch = [&](int rb, Pre &pre) {
            const int row = rb * TILE_M + quarter * 32 + lane;
            pre.off = row < p.num_nodes ? (long long)row * p.H + jb * 32 : -1;
#pragma unroll
            for (int i = 0; i < 8; ++i) pre.f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pre.off >= 0) {
                if (NPROD == 3) {
                    const float4 *src = reinterpret_cast<const float4 *>(p.h32 + pre.off);
#pragma unroll
                    for (int i = 0; i < 8; ++i) pre.f[i] = __ldg(src + i);
                } else {
                    const float4 *src = reinterpret_cast<const float4 *>(p.h16 + pre.off);
#pragma unroll
                    for (int i = 0; i < 4; ++i) pre.f[i] = __ldg(src + i);
                }
            }
        };
        Pre pre, pre_next;
        int rb = rb0 + set * rb_stride;
        if (rb < p.n_rb) prefetch(rb, pre);
        for (uint32_t use = 0; rb < p.n_rb; rb += rb_step, ++use) {
            const int nxt = rb + rb_step;
            if (nxt < p.n_rb) prefetch(nxt, pre_next);
            mbar_wait(&acc_full[set], use & 1);
            tc::tc_fence_after_sync();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float acc[64];
                {
                    uint32_t m[4][16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) tc::tmem_ld_16cols_async(tmem_lane + 32 * g + 16 * half, m[g]);
                    if (NPROD == 3) {
                        uint32_t cc[4][16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) tc::tmem_ld_16cols_async(tmem_lane + 128 + 32 * g + 16 * half, cc[g]);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                acc[16 * g + i] = fmaf(__uint_as_float(cc[g][i]), 1.0f / 2048.0f, __uint_as_float(m[g][i]));
                    } else {
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int i = 0; i < 16; ++i) acc[16 * g + i] = __uint_as_float(m[g][i]);
                    }
                }
                if (half == 1) {
                    tc::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[set]);      // the next tile's MMAs may overwrite this set
                }
                const int j0 = jb * 32 + 16 * half;
                float hv[16];
                if (NPROD == 3) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 v = pre.f[4 * half + i];
                        hv[4 * i] = v.x; hv[4 * i + 1] = v.y; hv[4 * i + 2] = v.z; hv[4 * i + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float4 v = pre.f[2 * half + i];
                        const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const __nv_bfloat162 hp = *reinterpret_cast<const __nv_bfloat162 *>(&w[u]);
                            hv[8 * i + 2 * u] = __low2float(hp); hv[8 * i + 2 * u + 1] = __high2float(hp);
                        }
                    }
                }
                float o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 b = bias_s[j0 + i];
                    if (NPROD == 3) {
                        const float rr = sigmoid_fast(acc[16 + i] + b.x);
                        const float zz = sigmoid_fast(acc[32 + i] + b.y);
                        const float nn = tanh_fast(acc[i] + b.z + rr * (acc[48 + i] + b.w));
                        o[i] = (1.0f - zz) * nn + zz * hv[i];
                    } else {
                        const float rr = sigmoid_mufu(acc[16 + i] + b.x);
                        const float zz = sigmoid_mufu(acc[32 + i] + b.y);
                        const float nn = tanh_mufu(fmaf(rr, acc[48 + i] + b.w, acc[i] + b.z));
                        o[i] = fmaf(zz, hv[i] - nn, nn);
                    }
                }
                if (pre.off >= 0) {
                    if (NPROD == 3) {
                        float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(p.out) + pre.off + 16 * half);
#pragma unroll
                        for (int i = 0; i < 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                        if (p.out_packed != nullptr) {      // same split as pack_states: hi = rn16(x), lo' = rn16((x - hi) * 2^11)
                            uint32_t hw[8], lw[8];
                            float big = 0.0f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const __half2 h2 = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
                                const float2 f2 = __half22float2(h2);
                                const __half2 l2 = __floats2half2_rn((o[2 * i] - f2.x) * 2048.0f, (o[2 * i + 1] - f2.y) * 2048.0f);
                                hw[i] = *reinterpret_cast<const uint32_t *>(&h2);
                                lw[i] = *reinterpret_cast<const uint32_t *>(&l2);
                                big = fmaxf(big, fmaxf(fabsf(o[2 * i]), fabsf(o[2 * i + 1])));
                            }
                            // row r = (off - 32 jb) / H holds 2H halfs: hi at [0, H), lo' at [H, 2H)
                            __half *rowp = p.out_packed + 2 * (pre.off - jb * 32) + jb * 32 + 16 * half;
                            uint4 *dh = reinterpret_cast<uint4 *>(rowp), *dl = reinterpret_cast<uint4 *>(rowp + p.H);
                            dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]); dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
                            dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]); dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                            if (!(big < 65504.0f) && p.status != nullptr) *reinterpret_cast<volatile int32_t *>(p.status) = 1;
                        }
                    } else {
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat162 v = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
                            w[i] = *reinterpret_cast<uint32_t *>(&v);
                        }
                        uint4 *dst = reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(p.out) + pre.off + 16 * half);
                        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            }
            pre = pre_next;
        }
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (warp == 2) tc::tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================================
// gate-blocked weight packing:  P1[part][jb][96][D] = [W_in; W_ir; W_iz],  P2[part][jb][128][H] = [0; W_hr; W_hz; W_hn]
// (weight_ih / weight_hh gate order is r, z, n);  NPROD 3: part 0 = fp16 hi, part 1 = fp16 lo';  NPROD 1: bf16
// =====================================================================================================================
__device__ __forceinline__ void split16(float x, uint16_t &hi, uint16_t &lo) {
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn((x - __half2float(h)) * 2048.0f));
}
template <int NPROD>
__global__ void __launch_bounds__(256) pack_gru_ws_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                                                          const float *__restrict__ b_ih, const float *__restrict__ b_hh, int H, int D,
                                                          uint16_t *__restrict__ p1, uint16_t *__restrict__ p2, float4 *__restrict__ bias4) {
    const int n_jb = H / 32;
    const long long n1 = (long long)n_jb * 96 * D, n2 = (long long)n_jb * 128 * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + H; i += (long long)gridDim.x * blockDim.x) {
        if (i >= n1 + n2) {
            const int j = (int)(i - n1 - n2);
            bias4[j] = make_float4(b_ih[j] + b_hh[j], b_ih[H + j] + b_hh[H + j], b_ih[2 * H + j], b_hh[2 * H + j]);
            continue;
        }
        float x;
        uint16_t *dst;
        long long idx, part_stride;
        if (i < n1) {
            const int k = (int)(i % D), n = (int)((i / D) % 96), jb = (int)(i / ((long long)96 * D));
            const int blk = n / 32, gate = blk == 0 ? 2 : blk - 1;       // rows: i_n (gate n), r, z
            x = w_ih[(size_t)(gate * H + jb * 32 + n % 32) * D + k];
            dst = p1; idx = i; part_stride = n1;
        } else {
            const long long r = i - n1;
            const int k = (int)(r % H), n = (int)((r / H) % 128), jb = (int)(r / ((long long)128 * H));
            const int blk = n / 32;                                       // rows: zero, r, z, h_n
            x = blk == 0 ? 0.0f : w_hh[(size_t)((blk - 1) * H + jb * 32 + n % 32) * H + k];
            dst = p2; idx = r; part_stride = n2;
        }
        if (NPROD == 3) {
            uint16_t hi, lo;
            split16(x, hi, lo);
            dst[idx] = hi;
            dst[part_stride + idx] = lo;
        } else {
            const __nv_bfloat16 v = __float2bfloat16_rn(x);
            dst[idx] = *reinterpret_cast<const uint16_t *>(&v);
        }
    }
}

// =====================================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
// 16-bit row-major [rows, cols], box {64 columns = 128 bytes, box_rows}, SWIZZLE_128B, out-of-bounds -> 0
static int make_map16(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool bf16) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return PTGNN_E_CUDA; }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (gru_ws) failed with CUresult %d", (int)r); return PTGNN_E_CUDA; }
    return PTGNN_OK;
}

bool supported(int nprod, int H, int D) {
    if (H % 64 != 0 || D % 64 != 0 || H < 64 || D < 64) return false;
    const int smem = nprod == 3 ? Geometry<3>::smem_bytes(H, D) : Geometry<1>::smem_bytes(H, D);
    return smem <= 232448 && 148 / (H / 32) >= 1;
}
size_t pack_bytes(int nprod, int H, int D) {
    const int npart = nprod == 3 ? 2 : 1;
    const size_t n_jb = H / 32;
    return ws_slice(npart * n_jb * 96 * D, 2) + ws_slice(npart * n_jb * 128 * H, 2) + ws_slice((size_t)H, 16);
}
static void pack_layout(int nprod, int H, int D, char *base, uint16_t *&p1, uint16_t *&p2, float4 *&bias4) {
    const int npart = nprod == 3 ? 2 : 1;
    const size_t n_jb = H / 32;
    p1 = reinterpret_cast<uint16_t *>(base);
    p2 = reinterpret_cast<uint16_t *>(base + ws_slice(npart * n_jb * 96 * D, 2));
    bias4 = reinterpret_cast<float4 *>(base + ws_slice(npart * n_jb * 96 * D, 2) + ws_slice(npart * n_jb * 128 * H, 2));
}
int pack(int nprod, int H, int D, const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh, void *packed, cudaStream_t st) {
    uint16_t *p1, *p2;
    float4 *bias4;
    pack_layout(nprod, H, D, static_cast<char *>(packed), p1, p2, bias4);
    {
        TimedScope timed__(PTGNN_KERNEL_PACK, st);
        if (nprod == 3) pack_gru_ws_kernel<3><<<148, 256, 0, st>>>(w_ih, w_hh, b_ih, b_hh, H, D, p1, p2, bias4);
        else pack_gru_ws_kernel<1><<<148, 256, 0, st>>>(w_ih, w_hh, b_ih, b_hh, H, D, p1, p2, bias4);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

int update(int nprod, const void *agg_rows, const void *h_rows, const void *h_plain, int64_t num_nodes, int H, int D, const void *packed,
           void *out, void *out_packed, int32_t *status, cudaStream_t st) {
    PTGNN_CHECK_ARG(supported(nprod, H, D), "gru_ws: unsupported dims H=%d D=%d", H, D);
    PTGNN_CHECK_ARG(out_packed == nullptr || nprod == 3, "gru_ws: packed output is an fp32-path (3xFP16) feature");
    if (num_nodes <= 0) return PTGNN_OK;
    const int npart = nprod == 3 ? 2 : 1;
    uint16_t *p1, *p2;
    float4 *bias4;
    pack_layout(nprod, H, D, const_cast<char *>(static_cast<const char *>(packed)), p1, p2, bias4);
    Params p{};
    const uint64_t n_jb = H / 32;
    const bool bf = nprod == 1;
    int rc = make_map16(&p.map_agg, agg_rows, num_nodes, (uint64_t)npart * D, 128, bf);
    if (!rc) rc = make_map16(&p.map_h, h_rows, num_nodes, (uint64_t)npart * H, 128, bf);
    if (!rc) rc = make_map16(&p.map_p1, p1, npart * n_jb * 96, D, 96, bf);
    if (!rc) rc = make_map16(&p.map_p2, p2, npart * n_jb * 128, H, 128, bf);
    if (rc) return rc;
    p.h32 = static_cast<const float *>(h_plain); p.h16 = static_cast<const __nv_bfloat16 *>(h_plain);
    p.bias4 = bias4; p.out = out; p.out_packed = static_cast<__half *>(out_packed); p.status = status; p.num_nodes = (int)num_nodes; p.H = H; p.D = D; p.n_jb = (int)n_jb;
    p.n_rb = (int)ceil_div(num_nodes, TILE_M);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    int groups = sms / (int)n_jb;                       // CTAs per hidden-unit block
    if (groups > p.n_rb) groups = p.n_rb;
    if (groups < 1) groups = 1;
    const int grid = groups * (int)n_jb;
    if (nprod == 3) {
        const int smem = Geometry<3>::smem_bytes(H, D);
        PTGNN_CUDA(cudaFuncSetAttribute(gru_ws_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        {
            TimedScope timed__(PTGNN_KERNEL_GRU, st);
            gru_ws_kernel<3><<<grid, NUM_THREADS, smem, st>>>(p);
        }
    } else {
        const int smem = Geometry<1>::smem_bytes(H, D);
        PTGNN_CUDA(cudaFuncSetAttribute(gru_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        {
            TimedScope timed__(PTGNN_KERNEL_GRU, st);
            gru_ws_kernel<1><<<grid, NUM_THREADS, smem, st>>>(p);
        }
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}

}  // namespace gruws
}  // namespace ptgnn
