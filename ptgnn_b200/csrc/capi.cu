// C-ABI glue: error string, launch counter, and the host-buffer entry point used for end-to-end measurement.
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace ptgnn {

static thread_local char g_error[512] = "";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---- optional per-kernel timing ---------------------------------------------------------------------
static std::atomic<int> g_timing_on{0};
struct TimedRecord { int cat; cudaEvent_t a, b; };
static std::mutex g_timing_mu;
static std::vector<TimedRecord> g_timing_records;

TimedScope::TimedScope(int category, cudaStream_t stream) : cat(category), st(stream) {
    if (g_timing_on.load(std::memory_order_relaxed)) {
        if (cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, st);
    }
}
TimedScope::~TimedScope() {
    if (a && b) {
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_timing_mu);
        g_timing_records.push_back({cat, a, b});
    }
}

namespace {
// RAII device buffer for the host-buffer entry point (the per-layer entry points never allocate).
struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
}  // namespace

}  // namespace ptgnn

using namespace ptgnn;

extern "C" int ptgnn_b200_abi_version(void) { return PTGNN_B200_ABI_VERSION; }
extern "C" const char *ptgnn_b200_last_error(void) { return g_error; }
extern "C" int64_t ptgnn_b200_launch_count(void) { return g_launch_count.load(); }

extern "C" int ptgnn_b200_kernel_timing_enable(int32_t enable) {
    g_timing_on.store(enable ? 1 : 0);
    return PTGNN_OK;
}

extern "C" int ptgnn_b200_kernel_timing_read(double *ms, int64_t *launches, int32_t ncat) {
    PTGNN_CHECK_ARG(ms && launches && ncat > 0, "kernel_timing_read: bad arguments");
    PTGNN_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (const TimedRecord &r : g_timing_records) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cat < ncat) {
            ms[r.cat] += t;
            launches[r.cat] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_timing_records.clear();
    return PTGNN_OK;
}

// Mirrors GraphNeuralNetwork.gnn's layer loop (reference ptgnn/neuralmodels/gnn/graphneuralnetwork.py:121-131) for
// a homogeneous stack of GatedMessagePassingLayers, from HOST buffers to HOST buffers.
extern "C" int ptgnn_b200_gated_gnn_forward_host_f32(const float *node_states, int64_t num_nodes, int32_t state_dim,
                                                     int32_t num_types, const int64_t *const *src_ptrs,
                                                     const int64_t *const *tgt_ptrs, const int64_t *counts,
                                                     int32_t num_layers, const float *const *edge_weights,
                                                     const float *const *gru_w_ih, const float *const *gru_w_hh,
                                                     const float *const *gru_b_ih, const float *const *gru_b_hh,
                                                     int32_t reduce, float *out_states) {
    PTGNN_CHECK_ARG(num_types >= 0 && num_types <= PTGNN_MAX_EDGE_TYPES, "gnn_forward_host: bad num_types=%d", num_types);
    PTGNN_CHECK_ARG(num_layers >= 1 && num_nodes >= 0 && state_dim > 0, "gnn_forward_host: bad sizes");
    const int H = state_dim;
    std::vector<int64_t> type_off(num_types + 1, 0);
    for (int t = 0; t < num_types; ++t) type_off[t + 1] = type_off[t] + counts[t];
    const int64_t E = type_off[num_types];
    cudaStream_t st = nullptr;

    DevBuf d_src, d_tgt, d_state[2], d_plan32, d_etype, d_status, d_ws, d_w;
    PTGNN_CUDA(d_src.alloc(sizeof(int64_t) * (size_t)E));
    PTGNN_CUDA(d_tgt.alloc(sizeof(int64_t) * (size_t)E));
    std::vector<const int64_t *> dsrc(num_types), dtgt(num_types);
    for (int t = 0; t < num_types; ++t) {
        dsrc[t] = d_src.as<int64_t>() + type_off[t];
        dtgt[t] = d_tgt.as<int64_t>() + type_off[t];
        if (counts[t]) {
            PTGNN_CUDA(cudaMemcpyAsync((void *)dsrc[t], src_ptrs[t], sizeof(int64_t) * (size_t)counts[t], cudaMemcpyHostToDevice, st));
            PTGNN_CUDA(cudaMemcpyAsync((void *)dtgt[t], tgt_ptrs[t], sizeof(int64_t) * (size_t)counts[t], cudaMemcpyHostToDevice, st));
        }
    }
    const size_t state_bytes = sizeof(float) * (size_t)num_nodes * H;
    PTGNN_CUDA(d_state[0].alloc(state_bytes));
    PTGNN_CUDA(d_state[1].alloc(state_bytes));
    PTGNN_CUDA(cudaMemcpyAsync(d_state[0].p, node_states, state_bytes, cudaMemcpyHostToDevice, st));

    // plan arrays: row_ptr[N+1] | perm | pos | src_sorted | src32 | tgt32 (int32 each, 256-byte aligned slices)
    const size_t sN = ws_slice((size_t)num_nodes + 1, 4), sE = ws_slice((size_t)E + 1, 4);
    PTGNN_CUDA(d_plan32.alloc(sN + 5 * sE));
    PTGNN_CUDA(d_etype.alloc((size_t)E + 1));
    PTGNN_CUDA(d_status.alloc(4));
    char *pb = d_plan32.as<char>();
    int32_t *row_ptr = reinterpret_cast<int32_t *>(pb), *perm = reinterpret_cast<int32_t *>(pb + sN);
    int32_t *pos = reinterpret_cast<int32_t *>(pb + sN + sE), *src_sorted = reinterpret_cast<int32_t *>(pb + sN + 2 * sE);
    int32_t *src32 = reinterpret_cast<int32_t *>(pb + sN + 3 * sE), *tgt32 = reinterpret_cast<int32_t *>(pb + sN + 4 * sE);

    const size_t ws_plan = ptgnn_b200_plan_workspace_bytes(num_nodes, E);
    const size_t ws_layer = ptgnn_b200_gated_workspace_bytes(num_nodes, E, num_types, H, H);
    const size_t ws_bytes = ws_plan > ws_layer ? ws_plan : ws_layer;
    PTGNN_CUDA(d_ws.alloc(ws_bytes));
    int rc = ptgnn_b200_plan_build(num_nodes, num_nodes, num_types, dsrc.data(), dtgt.data(), counts, row_ptr, perm, pos, src_sorted,
                                   d_etype.as<uint8_t>(), src32, tgt32, d_status.as<int32_t>(), d_ws.p, ws_bytes, st);
    if (rc) return rc;

    // weights: per layer T x [H,H] + w_ih [3H,H] + w_hh [3H,H] + b_ih [3H] + b_hh [3H]
    const size_t per_layer = (size_t)num_types * H * H + 6 * (size_t)H * H + 6 * (size_t)H;
    PTGNN_CUDA(d_w.alloc(sizeof(float) * per_layer * num_layers));
    std::vector<const float *> dev_w((size_t)num_types);
    for (int l = 0; l < num_layers; ++l) {
        float *base = d_w.as<float>() + per_layer * l;
        for (int t = 0; t < num_types; ++t)
            PTGNN_CUDA(cudaMemcpyAsync(base + (size_t)t * H * H, edge_weights[(size_t)l * num_types + t],
                                       sizeof(float) * H * H, cudaMemcpyHostToDevice, st));
        float *wih = base + (size_t)num_types * H * H, *whh = wih + 3 * (size_t)H * H;
        float *bih = whh + 3 * (size_t)H * H, *bhh = bih + 3 * H;
        PTGNN_CUDA(cudaMemcpyAsync(wih, gru_w_ih[l], sizeof(float) * 3 * H * H, cudaMemcpyHostToDevice, st));
        PTGNN_CUDA(cudaMemcpyAsync(whh, gru_w_hh[l], sizeof(float) * 3 * H * H, cudaMemcpyHostToDevice, st));
        PTGNN_CUDA(cudaMemcpyAsync(bih, gru_b_ih[l], sizeof(float) * 3 * H, cudaMemcpyHostToDevice, st));
        PTGNN_CUDA(cudaMemcpyAsync(bhh, gru_b_hh[l], sizeof(float) * 3 * H, cudaMemcpyHostToDevice, st));
    }
    int cur = 0;
    for (int l = 0; l < num_layers; ++l) {
        float *base = d_w.as<float>() + per_layer * l;
        for (int t = 0; t < num_types; ++t) dev_w[t] = base + (size_t)t * H * H;
        float *wih = base + (size_t)num_types * H * H, *whh = wih + 3 * (size_t)H * H;
        float *bih = whh + 3 * (size_t)H * H, *bhh = bih + 3 * H;
        rc = ptgnn_b200_gated_forward_f32(d_state[cur].as<float>(), nullptr, num_nodes, H, H, num_types, type_off.data(), row_ptr,
                                          pos, src32, dev_w.data(), wih, whh, bih, bhh, reduce,
                                          d_state[cur ^ 1].as<float>(), d_ws.p, ws_bytes, st);
        if (rc) return rc;
        cur ^= 1;
    }
    int32_t status = 0;
    PTGNN_CUDA(cudaMemcpyAsync(&status, d_status.p, 4, cudaMemcpyDeviceToHost, st));
    PTGNN_CUDA(cudaMemcpyAsync(out_states, d_state[cur].p, state_bytes, cudaMemcpyDeviceToHost, st));
    PTGNN_CUDA(cudaStreamSynchronize(st));
    if (status != 0) {
        set_error("gnn_forward_host: %d edge indices outside [0, %lld)", status, (long long)num_nodes);
        return PTGNN_E_INDEX;
    }
    return PTGNN_OK;
}
