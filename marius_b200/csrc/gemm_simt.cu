// gemm_simt.cu -- fp32 FFMA batched GEMM on the SIMT pipes (MB_PREC_FP32): the reference-exact arithmetic mode of the
// negative-score contraction (DotCompare bmm, comparators.cpp:69-72) and of its two backward products.  It is the
// numerical yardstick for the tcgen05 path (gemm_tc.cu) and the fallback for shapes TMA cannot address (d % 8 != 0).
//   C[b](m,n) = sum_k A[b](m,k) * B[b](k,n),  element strides given explicitly so one kernel serves
//   NT (scores = A . Neg^T), NN (dA = G . Neg) and TN (dNeg = G^T . A).
#include "common.cuh"

namespace mb {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256

struct GemmArgs {
    const float *A, *B;
    float* C;
    int M, N, K;
    int64_t sAm, sAk, sAb;
    int64_t sBk, sBn, sBb;
    int64_t ldc, sCb;
};

__global__ void __launch_bounds__(kThreads) gemm_simt_kernel(GemmArgs g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int b = blockIdx.z;
    const float* A = g.A + b * g.sAb;
    const float* B = g.B + b * g.sBb;
    float* C = g.C + b * g.sCb;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
    const bool a_k_contig = (g.sAk == 1);
    const bool b_n_contig = (g.sBn == 1);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
        for (int e = 0; e < (BM * BK) / kThreads; e++) {
            int lin = e * kThreads + threadIdx.x;
            int kk, mm;
            if (a_k_contig) {
                kk = lin % BK;
                mm = lin / BK;
            } else {
                mm = lin % BM;
                kk = lin / BM;
            }
            int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < g.M && k < g.K) ? A[(int64_t)m * g.sAm + (int64_t)k * g.sAk] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < (BN * BK) / kThreads; e++) {
            int lin = e * kThreads + threadIdx.x;
            int kk, nn;
            if (b_n_contig) {
                nn = lin % BN;
                kk = lin / BN;
            } else {
                kk = lin % BK;
                nn = lin / BK;
            }
            int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < g.N && k < g.K) ? B[(int64_t)k * g.sBk + (int64_t)n * g.sBn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM], bb[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; j++) bb[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int n = n0 + tx * TN + j;
            if (n < g.N) C[(int64_t)m * g.ldc + n] = acc[i][j];
        }
    }
}

}  // namespace

mb_status gemm_simt(const float* A, int64_t sAm, int64_t sAk, int64_t sAb, const float* B, int64_t sBk, int64_t sBn, int64_t sBb, float* C, int64_t ldc,
                    int64_t sCb, int M, int N, int K, int batches, cudaStream_t st) {
    if (M == 0 || N == 0 || batches == 0) return MB_OK;
    GemmArgs g{A, B, C, M, N, K, sAm, sAk, sAb, sBk, sBn, sBb, ldc, sCb};
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batches);
    gemm_simt_kernel<<<grid, kThreads, 0, st>>>(g);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
