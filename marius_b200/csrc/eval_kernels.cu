// eval_kernels.cu -- the evaluation side of the hot path: score filter and ranks of the positives among their negatives.
//   apply_score_filter            data/samplers/negative.cpp:306-311   scores[filter[:,0], filter[:,1]] = -1e9
//   LinkPredictionReporter::computeRanks   reporting/reporting.cpp:56-58   (neg >= pos.unsqueeze(1)).sum(1) + 1
// Both are integer / compare work on the score matrix the contraction just wrote: HBM (or L2) bound, one pass.
#include <algorithm>

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

// ---- score-filter construction on the device (compute_filter_corruption, data/samplers/negative.cpp:62-195, GLOBAL filter) ----
// The graph's edges are sorted (stably) by the endpoint a corruption keeps; for batch edge i every graph edge with the same kept
// endpoint and relation contributes the pair (i, id of its other endpoint) -- with all nodes as negatives that id IS the column.
// Pairs come out ordered by batch edge, then in sorted-pool order: the order of the reference's loop.
__global__ void filter_keys_kernel(const int64_t* __restrict__ edges, int64_t E, int cols, int key_col, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = (uint64_t)edges[i * cols + key_col];
        vals[i] = (uint32_t)i;
    }
}
__global__ void filter_gather_kernel(const int64_t* __restrict__ edges, int cols, const uint32_t* __restrict__ order, int64_t E, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E * cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        out[i] = edges[(int64_t)order[r] * cols + (i - r * cols)];
    }
}
// one warp per batch edge.  WRITE = false: counts[i] = matches; WRITE = true: the pairs, at offsets[i] (exclusive scan of the counts)
template <bool WRITE>
__global__ void __launch_bounds__(256) filter_match_kernel(const int64_t* __restrict__ pool, int64_t E, int cols, int key_col, int cor_col,
                                                           const int64_t* __restrict__ batch, int64_t B, int64_t* __restrict__ counts,
                                                           const int64_t* __restrict__ offsets, int64_t* __restrict__ out, int64_t cap) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B) return;
    const int64_t node = batch[i * cols + key_col];
    const int64_t rid = cols == 3 ? batch[i * cols + 1] : 0;
    int64_t lo = 0, hi = E;
    while (lo < hi) {  // first pool edge whose key >= node
        const int64_t mid = (lo + hi) >> 1;
        if (pool[mid * cols + key_col] < node) lo = mid + 1; else hi = mid;
    }
    int64_t written = WRITE ? offsets[i] : 0, total = 0;
    for (int64_t base = lo; base < E; base += 32) {
        const int64_t e = base + lane;
        const bool in_range = e < E && pool[e * cols + key_col] == node;
        const bool match = in_range && (cols != 3 || pool[e * cols + 1] == rid);
        const unsigned m = __ballot_sync(0xffffffffu, match);
        if (WRITE && match) {
            const int64_t pos = written + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) {
                out[2 * pos] = i;
                out[2 * pos + 1] = pool[e * cols + cor_col];
            }
        }
        written += __popc(m);
        total += __popc(m);
        if (__ballot_sync(0xffffffffu, in_range) != 0xffffffffu) break;  // ran past the key's range
    }
    if (!WRITE && lane == 0) counts[i] = total;
}
__global__ void exclusive_scan_kernel(const int64_t* __restrict__ counts, int64_t n, int64_t* __restrict__ offsets, int64_t* __restrict__ total) {
    // single block; n is a batch size (thousands)
    __shared__ int64_t carry;
    __shared__ int64_t tmp[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n ? counts[i] : 0;
        tmp[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int64_t t = threadIdx.x >= o ? tmp[threadIdx.x - o] : 0;
            __syncthreads();
            tmp[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) offsets[i] = carry + tmp[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += tmp[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// ---- streaming all-node evaluation: score filter and rank counts on one tile of the score matrix ----------------------------------
// scores [rows, ld] hold the scores of the positives against table rows [t0, t0 + T)
__global__ void filter_tile_kernel(float* __restrict__ scores, int64_t rows, int64_t ld, int64_t t0, int64_t T, const int64_t* __restrict__ filter, int64_t F,
                                   int64_t num_nodes, int* __restrict__ bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < F; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = filter[2 * i], c = filter[2 * i + 1];
        if (r < 0) r += rows;
        if (c < 0) c += num_nodes;
        if (r < 0 || r >= rows || c < 0 || c >= num_nodes) {
            *bad = 1;
            continue;
        }
        if (c >= t0 && c < t0 + T) scores[r * ld + (c - t0)] = -1e9f;
    }
}
__global__ void __launch_bounds__(256) rank_accumulate_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int64_t rows, int64_t T, int64_t ld,
                                                              int64_t* __restrict__ ranks) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = warp0; i < rows; i += nwarps) {
        const float p = pos[i];
        const float* row = neg + i * ld;
        int64_t cnt = 0;
        for (int64_t j = lane; j < T; j += 32) cnt += row[j] >= p;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) ranks[i] += cnt;  // (one writer per row and launch)
    }
}
__global__ void fill_i64_kernel(int64_t* p, int64_t n, int64_t v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

mb_status launch_filter_sort(const int64_t* edges, int64_t E, int cols, int key_col, uint64_t* ka, uint64_t* kb, uint32_t* va, uint32_t* vb, uint32_t* hist,
                             int key_bits, int64_t* sorted_out, cudaStream_t st) {
    if (E == 0) return MB_OK;
    const int blocks = (int)std::min<int64_t>((E + 255) / 256, (int64_t)sm_count() * 8);
    filter_keys_kernel<<<blocks, 256, 0, st>>>(edges, E, cols, key_col, ka, va);
    MB_LAUNCH_CHECK();
    uint64_t* ks = nullptr;
    uint32_t* vs = nullptr;
    MB_TRY(radix_sort_pairs<uint64_t>(ka, kb, va, vb, E, key_bits, hist, &ks, &vs, st));
    filter_gather_kernel<<<blocks, 256, 0, st>>>(edges, cols, vs, E, sorted_out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_filter_match(const int64_t* pool, int64_t E, int cols, int key_col, int cor_col, const int64_t* batch, int64_t B, int64_t* counts,
                              int64_t* offsets, int64_t* out, int64_t cap, int64_t* total_dev, cudaStream_t st) {
    if (B == 0) {
        MB_CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st));
        return MB_OK;
    }
    const int grid = (int)((B + 7) / 8);
    filter_match_kernel<false><<<grid, 256, 0, st>>>(pool, E, cols, key_col, cor_col, batch, B, counts, nullptr, nullptr, 0);
    MB_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, B, offsets, total_dev);
    MB_LAUNCH_CHECK();
    filter_match_kernel<true><<<grid, 256, 0, st>>>(pool, E, cols, key_col, cor_col, batch, B, nullptr, offsets, out, cap);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_filter_tile(float* scores, int64_t rows, int64_t ld, int64_t t0, int64_t T, const int64_t* filter, int64_t F, int64_t num_nodes, int* bad,
                             cudaStream_t st) {
    if (F == 0) return MB_OK;
    const int blocks = (int)std::min<int64_t>((F + 255) / 256, (int64_t)sm_count() * 8);
    filter_tile_kernel<<<blocks, 256, 0, st>>>(scores, rows, ld, t0, T, filter, F, num_nodes, bad);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_rank_accumulate(const float* pos, const float* neg, int64_t rows, int64_t T, int64_t ld, int64_t* ranks, cudaStream_t st) {
    if (rows == 0 || T == 0) return MB_OK;
    const int64_t blocks = std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 8);
    rank_accumulate_kernel<<<(int)blocks, 256, 0, st>>>(pos, neg, rows, T, ld, ranks);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_fill_i64(int64_t* p, int64_t n, int64_t v, cudaStream_t st) {
    if (n == 0) return MB_OK;
    fill_i64_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 1024), 256, 0, st>>>(p, n, v);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

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
