// tcgen05.mma issue-rate micro-benchmark: the tensor-pipe ceiling the roofline of the message / GRU kernels is quoted against.
//
// One CTA per SM; one elected thread issues `iters` x 4 back-to-back tcgen05.mma (M = 128, N given, K = 16 halfs / 8 tf32 per
// instruction) on operands that already sit in shared memory (SS) or tensor memory + shared memory (TS, the fused kernel's
// form); nothing is loaded or stored in the timed loop, so the time is the tensor pipe's and nothing else's.  Small N is part
// of the sweep on purpose: the fused aggregation kernel issues N = 16..64 (one (target block, edge type) group at a time), and
// the pipe does not run those at the N = 256 rate.
//
// Debug entry point (not in include/ptgnn_b200.h): tools/tc_peak.py drives it and writes profiles/r02_tcgen05_peaks.json,
// which bench.py reads for the `tensor` roofline denominators.
#include "tc_common.cuh"

namespace ptgnn {
namespace tcpeak {

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int A_BYTES = 128 * 128;          // 128 rows x 128 bytes (SWIZZLE_128B K-major): 4 K-steps
constexpr int B_BYTES = 256 * 128;          // up to N = 256 rows

// KIND: 0 = kind::f16 (fp16 operands), 2 = kind::tf32.  TS: A operand in tensor memory.
template <int KIND, bool TS>
__global__ void __launch_bounds__(128, 1) tc_peak_kernel(int N, int iters) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    unsigned char *base = smem_raw + pad;
    uint32_t *a_s = reinterpret_cast<uint32_t *>(base);
    uint32_t *b_s = reinterpret_cast<uint32_t *>(base + A_BYTES);
    uint64_t *done = reinterpret_cast<uint64_t *>(base + A_BYTES + B_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(base + A_BYTES + B_BYTES + 16);
    const int warp = threadIdx.x >> 5;
    // operand bits: finite values in [0.5, 1) with varying mantissas (data-dependent power is part of a realistic ceiling)
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15;
        a_s[i] = KIND == 0 ? (((h & 0x03FFu) | 0x3800u) | ((((h >> 10) & 0x03FFu) | 0x3800u) << 16)) : ((h & 0x007FE000u) | 0x3F000000u);
    }
    if (threadIdx.x == 0) { tc::mbar_init(done, 1); tc::mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (TS) {       // A: lane = row, 32 columns = the four K-steps (8 columns of packed halfs / tf32 values each)
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(a_s[(threadIdx.x * 32 + i) & (A_BYTES / 4 - 1)]);
        tc::tmem_st_32cols(tmem_base + ((uint32_t)(warp * 32) << 16) + 256, v);
        tc::tmem_st_wait();
        tc::tc_fence_before_sync();
        __syncthreads();
        tc::tc_fence_after_sync();
    }
    if (warp == 1) {
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::make_instr_desc(KIND == 0 ? 0u : tc::FMT_TF32, 128, (uint32_t)N);
        const uint64_t da = tc::make_smem_desc_sw128(smem_u32(a_s)), db = tc::make_smem_desc_sw128(smem_u32(b_s));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (leader) {
                    // two accumulators, alternated, when both fit below the A operand's columns (N <= 128)
                    const uint32_t d = tmem_base + ((N <= 128 && (ks & 1)) ? 128u : 0u);
                    if (TS) {
                        if (KIND == 0) mma_f16_ts(d, tmem_base + 256 + ks * 8, db + ks * 2, idesc, 1u);
                        else tc::mma_tf32_ts(d, tmem_base + 256 + ks * 8, db + ks * 2, idesc, 1u);
                    } else {
                        if (KIND == 0) tc::mma_bf16_ss(d, da + ks * 2, db + ks * 2, idesc, 1u);
                        else tc::mma_tf32_ss(d, da + ks * 2, db + ks * 2, idesc, 1u);
                    }
                }
            }
        }
        if (leader) tc::mma_commit(done);
        __syncwarp();
        tc::mbar_wait(done, 0);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

template <int KIND, bool TS>
static int run(int N, int iters, int grid, float *ms) {
    const int smem = A_BYTES + B_BYTES + 1024 + 64;
    PTGNN_CUDA(cudaFuncSetAttribute(tc_peak_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    PTGNN_CUDA(cudaEventCreate(&e0));
    PTGNN_CUDA(cudaEventCreate(&e1));
    tc_peak_kernel<KIND, TS><<<grid, 128, smem>>>(N, iters / 8 + 1);       // warm-up
    PTGNN_CUDA(cudaEventRecord(e0));
    tc_peak_kernel<KIND, TS><<<grid, 128, smem>>>(N, iters);
    PTGNN_CUDA(cudaEventRecord(e1));
    PTGNN_CUDA(cudaEventSynchronize(e1));
    PTGNN_CUDA(cudaGetLastError());
    PTGNN_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return PTGNN_OK;
}

}  // namespace tcpeak
}  // namespace ptgnn

// kind: 0 = f16, 2 = tf32; ts: A operand in tensor memory; n: MMA N (multiple of 16, 16..256); returns the kernel time of
// grid x iters x 4 MMAs of 128 x n x (16 | 8) in *ms; *grid_out = CTAs launched (one per SM)
extern "C" int ptgnn_b200_debug_tcgen05_peak(int kind, int ts, int n, int iters, float *ms, int *grid_out) {
    using namespace ptgnn;
    PTGNN_CHECK_ARG((kind == 0 || kind == 2) && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0 && ms, "tcgen05_peak: bad arguments");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    if (grid_out) *grid_out = sms;
    if (kind == 0) return ts ? tcpeak::run<0, true>(n, iters, sms, ms) : tcpeak::run<0, false>(n, iters, sms, ms);
    return ts ? tcpeak::run<2, true>(n, iters, sms, ms) : tcpeak::run<2, false>(n, iters, sms, ms);
}
