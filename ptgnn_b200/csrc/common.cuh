// Shared helpers for the ptgnn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/ptgnn_b200.h"

namespace ptgnn {

// ---- error reporting ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define PTGNN_CHECK_ARG(cond, ...)      \
    do {                                \
        if (!(cond)) {                  \
            ::ptgnn::set_error(__VA_ARGS__); \
            return PTGNN_E_INVALID;     \
        }                               \
    } while (0)

#define PTGNN_CUDA(call)                                                                          \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            ::ptgnn::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,               \
                               cudaGetErrorString(err__));                                        \
            return PTGNN_E_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

// RAII bracket around one kernel launch: when per-kernel timing is enabled (bench.py's roofline leg) it records a
// CUDA event on the launch stream before and after; otherwise it costs one relaxed atomic load.
struct TimedScope {
    int cat;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    TimedScope(int category, cudaStream_t stream);
    ~TimedScope();
};

// Call right after a kernel launch: counts it and surfaces launch-configuration errors.
#define PTGNN_LAUNCHED()                                   \
    do {                                                   \
        ::ptgnn::g_launch_count.fetch_add(1);              \
        PTGNN_CUDA(cudaGetLastError());                    \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

static inline size_t ws_slice(size_t count, size_t elt) { return align_up(count * elt, 256); }

// ---- per-type edge tables passed by value as a kernel parameter (< 4 KB) -------------------------
struct EdgeTables {
    const int64_t *src[PTGNN_MAX_EDGE_TYPES];
    const int64_t *tgt[PTGNN_MAX_EDGE_TYPES];
    int64_t off[PTGNN_MAX_EDGE_TYPES + 1];
    int num_types;
};

struct TypeOffsets {  // edge-id prefix offsets only (for kernels that need edge id -> type)
    int32_t off[PTGNN_MAX_EDGE_TYPES + 1];
    int num_types;
};

template <typename Off>
__device__ __forceinline__ int type_of_edge(const Off *off, int num_types, int64_t e) {
    // largest t with off[t] <= e  (empty types are skipped because off[t+1] == off[t] <= e moves on)
    int lo = 0, hi = num_types - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int64_t)off[mid] <= e) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes == 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- L2 residency control ---------------------------------------------------------------------------------------
// Optional (PTGNN_L2_HINTS=4): read the message rows in the reduce with an evict-first L2 policy.  Measured on B200:
// no gain (and `cp.async ... L2::cache_hint` for the gathers faults), so the default is plain streaming loads.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_stream_f4_hint(const float4 *p, uint64_t policy) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
    return r;
}

__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {  // read-once data: do not allocate in L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float apply_act(float x, int kind) {
    switch (kind) {
        case PTGNN_ACT_GELU: return gelu_erf(x);
        case PTGNN_ACT_TANH: return tanhf(x);
        case PTGNN_ACT_RELU: return x > 0.0f ? x : 0.0f;
        default: return x;
    }
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
// Gate math of the tensor-core GRU epilogues (the epilogue warps are few, so instruction count per output matters).
// fp32 path: ex2.approx exp (<= 2 ulp + |x| 2^-22 relative) and rcp.approx (1 ulp): sigmoid / tanh absolute error
// <= ~3e-7, an order below the layer tolerance.  bf16 path: tanh.approx (one MUFU, relative error 2^-11), well inside
// the bf16 rounding of the result.
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - 2.0f * rcp_approx(__expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ float tanh_mufu(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_mufu(float x) { return fmaf(0.5f, tanh_mufu(0.5f * x), 0.5f); }

}  // namespace ptgnn
