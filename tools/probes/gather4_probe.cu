// Probe (debug tool, not part of the library): does cp.async.bulk.tensor.2d...tile::gather4 work on this part, which box
// shape does the tensor map need, what is the shared-memory layout under SWIZZLE_128B, and how fast does one SM gather
// 128-byte row pieces with it?   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *b, uint32_t parity) {
    for (long long i = 0; i < 20000000ll; ++i) if (mbar_try(b, parity)) return true;
    return false;
}
__device__ __forceinline__ void gather4(void *dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

__global__ void correctness_kernel(const __grid_constant__ CUtensorMap map, uint16_t *out, int *status) {
    __shared__ __align__(1024) unsigned char buf[1024];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 1024; ++i) buf[i] = 0xEE;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(&bar, 512);
        gather4(buf, &map, 64, 5, 900, 17, 333, &bar);
        *status = mbar_wait(&bar, 0) ? 1 : -1;
        for (int i = 0; i < 512; ++i) out[i] = reinterpret_cast<uint16_t *>(buf)[i];
    }
}

// throughput: every CTA gathers `chunks` chunks of 128 random rows x 128 bytes through a SLOTS-deep ring.
// MODE 0: TMA gather4 (warp 0, lane l gathers rows 4l..4l+3); MODE 1: cp.async 16-byte pieces by 128 threads;
// MODE 2: rows 0..63 by gather4, rows 64..127 by cp.async.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
template <int SLOTS, int MODE>
__global__ void __launch_bounds__(128) throughput_kernel(const __grid_constant__ CUtensorMap map, const uint16_t *g, int C, const int *rows, int chunks, unsigned long long *sink) {
    extern __shared__ __align__(1024) unsigned char ring[];
    __shared__ uint64_t bars[SLOTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { for (int s = 0; s < SLOTS; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int *my = rows + (size_t)blockIdx.x * chunks * 128;
    unsigned long long acc = 0;
    constexpr int AHEAD = SLOTS - 1;
    const int tma_rows = MODE == 0 ? 128 : (MODE == 2 ? 64 : 0);
    for (int c = 0; c < chunks + AHEAD; ++c) {
        if (c < chunks) {
            const int s = c % SLOTS;
            if (tma_rows > 0 && warp == 0) {
                if (lane == 0) mbar_expect(&bars[s], tma_rows * 128);
                __syncwarp();
                if (lane * 4 < tma_rows) {
                    const int4 r = *reinterpret_cast<const int4 *>(my + c * 128 + lane * 4);
                    gather4(ring + s * 16384 + lane * 512, &map, 0, r.x, r.y, r.z, r.w, &bars[s]);
                }
            }
            if (MODE != 0) {
                const int q = tid & 7, rsub = tid >> 3;      // 16 rows per pass, 8 passes
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rsub + 16 * i;
                    if (r >= tma_rows) {
                        const int row = my[c * 128 + r];
                        cp_async16(smem_u32(ring + s * 16384 + r * 128 + ((q ^ (r & 7)) * 16)), g + (size_t)row * C + q * 8);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int d = c - AHEAD;
        if (d >= 0) {
            const int s = d % SLOTS;
            asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD) : "memory");
            if (tma_rows > 0 && !mbar_wait(&bars[s], (d / SLOTS) & 1)) { if (tid == 0) sink[1] = 0xDEAD; return; }
            __syncthreads();
            acc += *reinterpret_cast<const unsigned long long *>(ring + s * 16384 + tid * 8);
            __syncthreads();
        }
    }
    if (acc == 0x1234567) sink[0] = acc;
}
template <int SLOTS, int MODE>
static void run_tp(const CUtensorMap &map, const uint16_t *g, int C, const int *dr, int chunks, int grid, unsigned long long *sink) {
    const int smem = SLOTS * 16384 + 1024;
    cudaFuncSetAttribute(throughput_kernel<SLOTS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(a);
        throughput_kernel<SLOTS, MODE><<<grid, 128, smem>>>(map, g, C, dr, chunks, sink);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  throughput kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    unsigned long long sk[2]; cudaMemcpy(sk, sink, 16, cudaMemcpyDeviceToHost);
    const double bytes = (double)grid * chunks * 128 * 128;
    printf("  slots=%2d mode=%d (%s): %.3f ms, %7.1f GB/s (%.1f GB/s per SM)%s\n", SLOTS, MODE,
           MODE == 0 ? "gather4" : (MODE == 1 ? "cp.async" : "half/half"), best, bytes / best * 1e-6, bytes / best * 1e-6 / grid, sk[1] ? "  TIMEOUT" : "");
}

int main() {
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int R = 262144, C = 128;   // bf16 [R, C] = 64 MB
    std::vector<uint16_t> h((size_t)R * C);
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (uint16_t)((r * 131 + c) & 0xFFFF);
    uint16_t *g; cudaMalloc(&g, h.size() * 2); cudaMemcpy(g, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    uint16_t *out; cudaMalloc(&out, 1024); int *status; cudaMalloc(&status, 4);
    for (int box_rows = 1; box_rows <= 1; ++box_rows) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}; cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
        CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box_rows=%d encode=%d\n", box_rows, (int)cr);
        if (cr != CUDA_SUCCESS) continue;
        cudaMemset(status, 0, 4); cudaMemset(out, 0, 1024);
        correctness_kernel<<<1, 32>>>(map, out, status);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; uint16_t o[512];
        if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 1024, cudaMemcpyDeviceToHost);
        printf("  status=%d\n", st);
        if (st != 1) continue;
        const int rows[4] = {5, 900, 17, 333};
        int ok_swz = 0, ok_lin = 0;
        for (int i = 0; i < 4; ++i) for (int c = 0; c < 64; ++c) {
            const uint16_t want = (uint16_t)((rows[i] * 131 + 64 + c) & 0xFFFF);
            const int chunk = c / 8, within = c % 8;
            ok_swz += o[i * 64 + ((chunk ^ (i & 7)) * 8) + within] == want;
            ok_lin += o[i * 64 + c] == want;
        }
        printf("  match swizzled=%d/256 linear=%d/256  first words: %04x %04x %04x %04x\n", ok_swz, ok_lin, o[0], o[1], o[64], o[65]);
        if (ok_swz != 256) continue;
        // throughput
        const int chunks = 256, grid = 148;
        std::vector<int> hr((size_t)grid * chunks * 128);
        uint64_t s = 88172645463325252ull;
        for (auto &v : hr) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int)(s % R); }
        int *dr; cudaMalloc(&dr, hr.size() * 4); cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
        unsigned long long *sink; cudaMalloc(&sink, 16); cudaMemset(sink, 0, 16);
        run_tp<4, 0>(map, g, C, dr, chunks, grid, sink);  run_tp<8, 0>(map, g, C, dr, chunks, grid, sink);  run_tp<12, 0>(map, g, C, dr, chunks, grid, sink);
        run_tp<4, 1>(map, g, C, dr, chunks, grid, sink);  run_tp<8, 1>(map, g, C, dr, chunks, grid, sink);  run_tp<12, 1>(map, g, C, dr, chunks, grid, sink);
        run_tp<4, 2>(map, g, C, dr, chunks, grid, sink);  run_tp<8, 2>(map, g, C, dr, chunks, grid, sink);  run_tp<12, 2>(map, g, C, dr, chunks, grid, sink);
    }
    return 0;
}
