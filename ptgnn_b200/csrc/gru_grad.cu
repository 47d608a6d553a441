// GRUCell backward, pointwise half: from the gate pre-activations gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh
// (gate order r, z, n; torch.nn.GRUCell as used at ptgnn/neuralmodels/gnn/messagepassing/gatedmessagepassing.py:69), the previous
// state h and the upstream gradient g of h' = (1 - z) n + z h, computes in ONE pass
//   d_gi = [d_r, d_z, d_n],   d_gh = [d_r, d_z, d_n * r],   d_h_direct = g z
// with d_n = g (1 - z)(1 - n^2), d_z = g (h - n) z (1 - z), d_r = d_n h_n r (1 - r).  The four GEMM-shaped products around it
// (gi, gh, d_gi W_ih, d_gh W_hh) run on the dense kernels; as separate torch pointwise ops this was 17 % of a training step.
#include "common.cuh"

namespace ptgnn {

__global__ void __launch_bounds__(256) gru_gate_grads_kernel(const float *__restrict__ gi, const float *__restrict__ gh,
                                                             const float *__restrict__ h, const float *__restrict__ g, long long rows, int H,
                                                             float *__restrict__ d_gi, float *__restrict__ d_gh, float *__restrict__ d_h) {
    const long long total = rows * (H / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (H / 4);
        const int j = (int)(i - row * (H / 4)) * 4;
        const float *gi_r = gi + row * 3 * H + j, *gh_r = gh + row * 3 * H + j;
        const float4 ir = *reinterpret_cast<const float4 *>(gi_r), iz = *reinterpret_cast<const float4 *>(gi_r + H),
                     in_ = *reinterpret_cast<const float4 *>(gi_r + 2 * H);
        const float4 hr = *reinterpret_cast<const float4 *>(gh_r), hz = *reinterpret_cast<const float4 *>(gh_r + H),
                     hn = *reinterpret_cast<const float4 *>(gh_r + 2 * H);
        const float4 hv = *reinterpret_cast<const float4 *>(h + row * H + j), gv = *reinterpret_cast<const float4 *>(g + row * H + j);
        const float a_ir[4] = {ir.x, ir.y, ir.z, ir.w}, a_iz[4] = {iz.x, iz.y, iz.z, iz.w}, a_in[4] = {in_.x, in_.y, in_.z, in_.w};
        const float a_hr[4] = {hr.x, hr.y, hr.z, hr.w}, a_hz[4] = {hz.x, hz.y, hz.z, hz.w}, a_hn[4] = {hn.x, hn.y, hn.z, hn.w};
        const float a_h[4] = {hv.x, hv.y, hv.z, hv.w}, a_g[4] = {gv.x, gv.y, gv.z, gv.w};
        float dr[4], dz[4], dn[4], dnr[4], dh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float r = 1.0f / (1.0f + expf(-(a_ir[k] + a_hr[k])));
            const float z = 1.0f / (1.0f + expf(-(a_iz[k] + a_hz[k])));
            const float n = tanhf(a_in[k] + r * a_hn[k]);
            dn[k] = a_g[k] * (1.0f - z) * (1.0f - n * n);
            dz[k] = a_g[k] * (a_h[k] - n) * z * (1.0f - z);
            dr[k] = dn[k] * a_hn[k] * r * (1.0f - r);
            dnr[k] = dn[k] * r;
            dh[k] = a_g[k] * z;
        }
        float *o_gi = d_gi + row * 3 * H + j, *o_gh = d_gh + row * 3 * H + j;
        *reinterpret_cast<float4 *>(o_gi) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        *reinterpret_cast<float4 *>(o_gi + H) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        *reinterpret_cast<float4 *>(o_gi + 2 * H) = make_float4(dn[0], dn[1], dn[2], dn[3]);
        *reinterpret_cast<float4 *>(o_gh) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        *reinterpret_cast<float4 *>(o_gh + H) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        *reinterpret_cast<float4 *>(o_gh + 2 * H) = make_float4(dnr[0], dnr[1], dnr[2], dnr[3]);
        *reinterpret_cast<float4 *>(d_h + row * H + j) = make_float4(dh[0], dh[1], dh[2], dh[3]);
    }
}

}  // namespace ptgnn

extern "C" int ptgnn_b200_gru_gate_grads_f32(const float *gi, const float *gh, const float *h, const float *grad_out, int64_t num_nodes,
                                             int32_t state_dim, float *d_gi, float *d_gh, float *d_h_direct, void *stream) {
    using namespace ptgnn;
    PTGNN_CHECK_ARG(num_nodes >= 0 && state_dim > 0 && state_dim % 4 == 0, "gru_gate_grads: state_dim=%d must be a positive multiple of 4", state_dim);
    if (num_nodes == 0) return PTGNN_OK;
    PTGNN_CHECK_ARG(gi && gh && h && grad_out && d_gi && d_gh && d_h_direct, "gru_gate_grads: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = (long long)num_nodes * (state_dim / 4);
    const long long blocks = (total + 255) / 256;
    {
        TimedScope timed__(PTGNN_KERNEL_GRU, st);
        gru_gate_grads_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(gi, gh, h, grad_out, num_nodes, state_dim, d_gi, d_gh,
                                                                                                 d_h_direct);
    }
    PTGNN_LAUNCHED();
    return PTGNN_OK;
}
