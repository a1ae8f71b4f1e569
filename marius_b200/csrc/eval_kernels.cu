// eval_kernels.cu -- the evaluation side of the hot path: score filter and ranks of the positives among their negatives.
//   apply_score_filter            data/samplers/negative.cpp:306-311   scores[filter[:,0], filter[:,1]] = -1e9
//   LinkPredictionReporter::computeRanks   reporting/reporting.cpp:56-58   (neg >= pos.unsqueeze(1)).sum(1) + 1
// Both are integer / compare work on the score matrix the contraction just wrote: HBM (or L2) bound, one pass.
#include "kernels.h"

namespace mb {
namespace {

__global__ void score_filter_kernel(float* __restrict__ scores, int64_t rows, int64_t N, int64_t ld, const int64_t* __restrict__ filter, int64_t F,
                                    int* __restrict__ bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = filter[2 * i], c = filter[2 * i + 1];
        // torch index_put_ semantics: negative indices wrap once, anything else out of range is an error (reported to the host)
        if (r < 0) r += rows;
        if (c < 0) c += N;
        if (r < 0 || r >= rows || c < 0 || c >= N) {
            *bad = 1;
            continue;
        }
        scores[r * ld + c] = -1e9f;
    }
}

// one warp per row; 128-bit loads when the row is 16-byte aligned, scalar otherwise.  Counts are exact integers, so the order of
// the reduction does not matter.
__global__ void __launch_bounds__(256) rank_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int64_t rows, int64_t N, int64_t ld,
                                                   int64_t* __restrict__ ranks) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = warp0; i < rows; i += nwarps) {
        const float p = pos[i];
        const float* row = neg + i * ld;
        int64_t cnt = 0;
        if ((reinterpret_cast<uintptr_t>(row) & 15u) == 0) {
            const int64_t nv = N >> 2;
            const float4* r4 = reinterpret_cast<const float4*>(row);
            for (int64_t v = lane; v < nv; v += 32) {
                const float4 x = r4[v];
                cnt += (x.x >= p) + (x.y >= p) + (x.z >= p) + (x.w >= p);
            }
            for (int64_t j = (nv << 2) + lane; j < N; j += 32) cnt += row[j] >= p;
        } else {
            for (int64_t j = lane; j < N; j += 32) cnt += row[j] >= p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) ranks[i] = cnt + 1;
    }
}

}  // namespace

mb_status launch_score_filter(float* scores, int64_t rows, int64_t N, int64_t ld, const int64_t* filter, int64_t F, int* bad_flag, cudaStream_t st) {
    if (F == 0) return MB_OK;
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((F + threads - 1) / threads, (int64_t)sm_count() * 8);
    score_filter_kernel<<<blocks, threads, 0, st>>>(scores, rows, N, ld, filter, F, bad_flag);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_ranks(const float* pos, const float* neg, int64_t rows, int64_t N, int64_t ld, int64_t* ranks, cudaStream_t st) {
    if (rows == 0) return MB_OK;
    const int64_t blocks = std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 8);
    rank_kernel<<<(int)blocks, 256, 0, st>>>(pos, neg, rows, N, ld, ranks);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
