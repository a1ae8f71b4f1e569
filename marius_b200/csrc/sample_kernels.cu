// sample_kernels.cu -- batch assembly on the device: negative sampling and the id concatenation / split around
// map_tensors.
//   CorruptNodeNegativeSampler::getNegatives   data/samplers/negative.cpp:328-366  (uniform part: torch::randint(num_nodes, {n}))
//   batch_sample                               data/samplers/negative.cpp:7-19     (degree part: an endpoint of a random batch edge)
//   DataLoader::edgeSample                     data/dataloader.cpp:389-471         (cat(src, dst, src_negs, dst_negs) -> map_tensors -> split)
// The reference draws from libtorch's global mt19937 / Philox generator; a device sampler cannot reproduce that stream, so the ids are
// defined by a stateless counter-based generator instead (Philox4x32-10, key = seed, counter = (element, side, batch)), which the
// oracle restates bit for bit.  The integer in [0, range) is the 64-bit draw modulo range, the transformation libtorch's randint
// applies (ATen random_from_to: `random % range + base`).
#include "kernels.h"

namespace mb {
namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
}

// Philox4x32-10: 64 random bits for (seed, element index, stream, batch)
__device__ __forceinline__ uint64_t philox_u64(uint64_t seed, uint64_t index, uint32_t stream, uint32_t batch) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), stream, batch};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return (uint64_t)c[0] | ((uint64_t)c[1] << 32);
}

__global__ void sample_negatives_kernel(int64_t num_nodes, int64_t total, int N, int num_batch, const int64_t* __restrict__ edges, int64_t B, int cols,
                                        int endpoint_col, uint64_t seed, uint32_t stream_id, uint32_t batch, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t r = philox_u64(seed, (uint64_t)i, stream_id, batch);
        const int j = (int)(i % N);
        if (j < num_batch) {  // degree-based part comes first in each chunk: torch::cat({deg_sample, uniform}) (negative.cpp:347)
            const int64_t e = (int64_t)(r % (uint64_t)B);
            out[i] = edges[e * cols + endpoint_col];
        } else {
            out[i] = (int64_t)(r % (uint64_t)num_nodes);
        }
    }
}

// all_ids = cat(src, dst, src_negs.flatten(), dst_negs.flatten())   (dataloader.cpp:398-406)
__global__ void concat_ids_kernel(const int64_t* __restrict__ edges, int64_t B, int cols, const int64_t* __restrict__ src_negs, const int64_t* __restrict__ dst_negs,
                                  int64_t CN, int64_t* __restrict__ all_ids) {
    const int64_t n_s = src_negs ? CN : 0, total = 2 * B + n_s + CN;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v;
        if (i < B)
            v = edges[i * cols];
        else if (i < 2 * B)
            v = edges[(i - B) * cols + cols - 1];
        else if (i < 2 * B + n_s)
            v = src_negs[i - 2 * B];
        else
            v = dst_negs[i - 2 * B - n_s];
        all_ids[i] = v;
    }
}

// edges_ = stack(src_mapping, rel, dst_mapping) ; the negative mappings keep their [C,N] shape   (dataloader.cpp:447-470)
__global__ void split_mapped_kernel(const int64_t* __restrict__ mapped, const int64_t* __restrict__ edges, int64_t B, int cols, int64_t n_s, int64_t CN,
                                    int64_t* __restrict__ edges_local, int64_t* __restrict__ src_negs_local, int64_t* __restrict__ dst_negs_local) {
    const int64_t total = 2 * B + n_s + CN;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = mapped[i];
        if (i < B) {
            edges_local[i * cols] = m;
            if (cols == 3) edges_local[i * cols + 1] = edges[i * cols + 1];
        } else if (i < 2 * B) {
            edges_local[(i - B) * cols + cols - 1] = m;
        } else if (i < 2 * B + n_s) {
            src_negs_local[i - 2 * B] = m;
        } else {
            dst_negs_local[i - 2 * B - n_s] = m;
        }
    }
}

inline int grid_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8)); }

}  // namespace

mb_status launch_sample_negatives(int64_t num_nodes, int C, int N, int num_batch, const int64_t* edges, int64_t B, int cols, bool inverse, uint64_t seed,
                                  uint32_t batch, int64_t* out, cudaStream_t st) {
    const int64_t total = (int64_t)C * N;
    if (total == 0) return MB_OK;
    // inverse (source corruption) samples the source column, else the destination column (negative.cpp:13-17)
    sample_negatives_kernel<<<grid_for(total), 256, 0, st>>>(num_nodes, total, N, num_batch, edges, B, cols, inverse ? 0 : cols - 1, seed, inverse ? 1u : 0u, batch,
                                                             out);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_concat_ids(const int64_t* edges, int64_t B, int cols, const int64_t* src_negs, const int64_t* dst_negs, int64_t CN, int64_t* all_ids,
                            cudaStream_t st) {
    const int64_t total = 2 * B + (src_negs ? CN : 0) + CN;
    if (total == 0) return MB_OK;
    concat_ids_kernel<<<grid_for(total), 256, 0, st>>>(edges, B, cols, src_negs, dst_negs, CN, all_ids);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_split_mapped(const int64_t* mapped, const int64_t* edges, int64_t B, int cols, bool has_src_negs, int64_t CN, int64_t* edges_local,
                              int64_t* src_negs_local, int64_t* dst_negs_local, cudaStream_t st) {
    const int64_t n_s = has_src_negs ? CN : 0, total = 2 * B + n_s + CN;
    if (total == 0) return MB_OK;
    split_mapped_kernel<<<grid_for(total), 256, 0, st>>>(mapped, edges, B, cols, n_s, CN, edges_local, src_negs_local, dst_negs_local);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
