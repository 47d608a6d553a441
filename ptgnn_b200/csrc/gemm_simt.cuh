// fp32-exact (FFMA) row-gather GEMM core:  C[m][n] = sum_k A[row(m)][k] * B[n][k]
//
// Both operands are "K-major" (A rows = node-state rows gathered by index, B rows = nn.Linear.weight rows),
// which is exactly how the reference stores them (F.embedding rows, nn.Linear.weight [out, in]):
//   reference ptgnn/neuralmodels/gnn/messagepassing/gatedmessagepassing.py:54-60  (gather + Linear)
//   reference ptgnn/neuralmodels/gnn/messagepassing/mlpmessagepassing.py:88-98    (two gathers + cat + MLP)
// so no transposition or [E, .] intermediate is ever materialised: the gather IS the A-tile load.
//
// CTA tile 128 x (16*TN) x 32, 256 threads (16 x 16), 8 x TN accumulators per thread, 3-stage cp.async (LDGSTS)
// pipeline, XOR-swizzled 128-byte shared-memory rows (conflict-free 16-byte fragment loads for both operands).
// The A operand may be the concatenation of two gathered rows ([h_src ; h_tgt], [agg ; h]): columns < K0 come
// from (a0, idx0), the rest from (a1, idx1).
#pragma once
#include "common.cuh"

namespace ptgnn {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 32;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGES = 3;

template <int TN>
struct GemmTile {
    static constexpr int BN = 16 * TN;
    static constexpr int STAGE_FLOATS = (GEMM_BM + BN) * GEMM_BK;
    static constexpr int PIPE_BYTES = GEMM_STAGES * STAGE_FLOATS * 4;
    static constexpr int CS_STRIDE = BN + 16;                       // staging row pitch (floats)
    static constexpr int CS_BYTES = GEMM_BM * CS_STRIDE * 4;
    static constexpr int IDX_BYTES = 3 * GEMM_BM * 4;               // idx0, idx1, out-row
    static constexpr int SMEM_BYTES = (PIPE_BYTES > CS_BYTES ? PIPE_BYTES : CS_BYTES) + IDX_BYTES;
};

struct AOperand {
    const float *a0;  // rows for k in [0, K0)
    const float *a1;  // rows for k in [K0, K)   (may be nullptr when K0 == K)
    int ld0, ld1;     // row pitches (floats)
    int K0, K;        // K0 % 4 == 0, K % 4 == 0
};

template <int TN>
__device__ __forceinline__ void gemm_load_stage(float *sA, float *sB, const AOperand &A, const int *s_idx0,
                                                const int *s_idx1, const float *__restrict__ Bw, int ldb, int n0,
                                                int Nb, int k0, int tid) {
    constexpr int BN = GemmTile<TN>::BN;
    // A: 128 rows x 8 sixteen-byte chunks
#pragma unroll
    for (int i = 0; i < (GEMM_BM * 8) / GEMM_THREADS; ++i) {
        const int c = tid + i * GEMM_THREADS;
        const int row = c >> 3, q = c & 7;
        const int kk = k0 + q * 4;
        const float *src = A.a0;
        int bytes = 0;
        if (kk < A.K0) {
            const int r = s_idx0[row];
            if (r >= 0) { src = A.a0 + (size_t)r * A.ld0 + kk; bytes = 16; }
        } else if (kk < A.K) {
            const int r = s_idx1[row];
            if (r >= 0) { src = A.a1 + (size_t)r * A.ld1 + (kk - A.K0); bytes = 16; }
        }
        cp_async16(smem_u32(sA + row * GEMM_BK + ((q ^ (row & 7)) << 2)), src, bytes);
    }
    // B: BN rows x 8 chunks
#pragma unroll
    for (int i = 0; i < (BN * 8 + GEMM_THREADS - 1) / GEMM_THREADS; ++i) {
        const int c = tid + i * GEMM_THREADS;
        if ((BN * 8) % GEMM_THREADS != 0 && c >= BN * 8) break;
        const int row = c >> 3, q = c & 7;
        const int kk = k0 + q * 4;
        const int n = n0 + row;
        const float *src = Bw;
        int bytes = 0;
        if (n < Nb && kk < A.K) { src = Bw + (size_t)n * ldb + kk; bytes = 16; }
        cp_async16(smem_u32(sB + row * GEMM_BK + ((q ^ (row & 7)) << 2)), src, bytes);
    }
}

// acc column used by fragment column j:  j < 4 -> j,  j >= 4 -> j + ACC_SHIFT   (GRU phase 2 uses shift 2)
template <int TN, int ACC_SHIFT>
__device__ __forceinline__ void gemm_compute_stage(const float *sA, const float *sB, float (&acc)[8][8], int tx,
                                                   int ty) {
#pragma unroll
    for (int kq = 0; kq < GEMM_BK / 4; ++kq) {
        float4 b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = tx + 16 * j;
            b[j] = *reinterpret_cast<const float4 *>(sB + n * GEMM_BK + ((kq ^ (n & 7)) << 2));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = ty + 16 * i;
            const float4 a = *reinterpret_cast<const float4 *>(sA + m * GEMM_BK + ((kq ^ (m & 7)) << 2));
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                float &c = acc[i][j < 4 ? j : j + ACC_SHIFT];
                c = fmaf(a.x, b[j].x, c);
                c = fmaf(a.y, b[j].y, c);
                c = fmaf(a.z, b[j].z, c);
                c = fmaf(a.w, b[j].w, c);
            }
        }
    }
}

// Runs the whole K loop for one CTA tile.  On return all cp.async groups are drained and the CTA is synchronised,
// so the pipeline buffers may be reused (second phase / C staging).
template <int TN, int ACC_SHIFT>
__device__ __forceinline__ void gemm_mainloop(float (&acc)[8][8], float *pipe, const AOperand &A, const int *s_idx0,
                                              const int *s_idx1, const float *__restrict__ Bw, int ldb, int n0, int Nb) {
    constexpr int STAGE = GemmTile<TN>::STAGE_FLOATS;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int num_kt = (A.K + GEMM_BK - 1) / GEMM_BK;
#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < num_kt)
            gemm_load_stage<TN>(pipe + s * STAGE, pipe + s * STAGE + GEMM_BM * GEMM_BK, A, s_idx0, s_idx1, Bw, ldb, n0,
                                Nb, s * GEMM_BK, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < num_kt; ++kt) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        const int pf = kt + GEMM_STAGES - 1;
        if (pf < num_kt) {
            float *st = pipe + (pf % GEMM_STAGES) * STAGE;
            gemm_load_stage<TN>(st, st + GEMM_BM * GEMM_BK, A, s_idx0, s_idx1, Bw, ldb, n0, Nb, pf * GEMM_BK, tid);
        }
        cp_async_commit();
        const float *cur = pipe + (kt % GEMM_STAGES) * STAGE;
        gemm_compute_stage<TN, ACC_SHIFT>(cur, cur + GEMM_BM * GEMM_BK, acc, tx, ty);
    }
    cp_async_wait<0>();
    __syncthreads();
}

}  // namespace ptgnn
