// radix_sort.cu -- stable LSD radix sort of (key, slot) pairs + CSR / unique helpers.
//
// Two users on the hot path:
//  (1) duplicate-index accumulation of the backward pass: the reference's autograd index_select-backward is an
//      index_add WITH duplicates (decoder_methods.cpp:74-79 under model.cpp:324).  We sort the 2B+2CN gradient
//      "slots" (src | dst | src_negs | dst_negs -- the order of DataLoader::edgeSample's all_ids,
//      dataloader.cpp:399-409) by batch-local node id; each unique node then sums its slots in slot order with no
//      atomics: deterministic run-to-run.
//  (2) map_tensors (common/util.cpp:180-205): torch::_unique2(sorted=true, return_inverse=true) over global ids.
//
// 8-bit digits; per pass: tile histogram -> single-block scan over [256 digits][tiles] -> stable tile scatter
// (warp match_any ranks, tile processed in slot order).  Sizes here are O(10^4..10^6) keys: integer work in L2.
#include "common.cuh"

namespace mb {

namespace {

constexpr int kSortThreads = 256;
constexpr int kItems = 8;                        // keys per thread per tile
constexpr int kTile = kSortThreads * kItems;     // 2048 keys per block
constexpr int kWarps = kSortThreads / 32;

template <typename K>
__global__ void __launch_bounds__(kSortThreads) hist_kernel(const K* __restrict__ keys, int64_t n, int shift, int num_tiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
    for (int j = 0; j < kItems; j++) {
        int64_t i = base + j * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)((keys[i] >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// exclusive scan of `n` uint32 in place, single block (n is 256 * tiles or a flag array: small)
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 4096) {
        int64_t i0 = base + (int64_t)threadIdx.x * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (i0 + j < n) ? data[i0 + j] : 0u;
        uint32_t local = v[0] + v[1] + v[2] + v[3];
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint32_t ws = warp_sums[lane];
            uint32_t wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - ws;  // exclusive
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_sums[w] + (incl - local);
        uint32_t run = excl;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i0 + j < n) data[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

template <typename K>
__global__ void __launch_bounds__(kSortThreads) scatter_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, int64_t n, int shift, int num_tiles,
                                                               const uint32_t* __restrict__ hist_scanned, bool iota_vals) {
    __shared__ uint32_t base[256];             // running global offset of each digit for this tile
    __shared__ uint32_t warp_cnt[kWarps][256];
    base[threadIdx.x] = hist_scanned[(int64_t)threadIdx.x * num_tiles + blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int64_t tile_base = (int64_t)blockIdx.x * kTile;
    for (int j = 0; j < kItems; j++) {
#pragma unroll
        for (int ww = 0; ww < kWarps; ww++) warp_cnt[ww][threadIdx.x] = 0;
        __syncthreads();
        int64_t i = tile_base + j * kSortThreads + threadIdx.x;
        bool ok = i < n;
        K key = ok ? keys_in[i] : K(0);
        uint32_t val = ok ? (iota_vals ? (uint32_t)i : vals_in[i]) : 0u;
        uint32_t digit = ok ? (uint32_t)((key >> shift) & 0xff) : 0x100u;  // invalid lanes get a digit of their own
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        uint32_t rank = __popc(peers & lt_mask);
        if (ok && rank == 0) warp_cnt[w][digit] = __popc(peers);
        __syncthreads();
        // thread d: exclusive prefix over warps for digit d, then advance the running base
        uint32_t b = base[threadIdx.x];
        uint32_t acc = 0;
#pragma unroll
        for (int ww = 0; ww < kWarps; ww++) {
            uint32_t c = warp_cnt[ww][threadIdx.x];
            warp_cnt[ww][threadIdx.x] = b + acc;
            acc += c;
        }
        __syncthreads();
        if (ok) {
            uint32_t dst = warp_cnt[w][digit] + rank;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
        base[threadIdx.x] = b + acc;
        // (next iteration's zeroing of warp_cnt is ordered by the __syncthreads above)
    }
}

// offsets[k] = first sorted position whose key >= k, for k in [0, num_keys]; keys sorted ascending.
template <typename K>
__global__ void seg_offsets_kernel(const K* __restrict__ sorted_keys, int64_t n, int64_t num_keys, uint32_t* __restrict__ offsets) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t prev = (i == 0) ? -1 : (int64_t)sorted_keys[i - 1];
        int64_t cur = (i == n) ? num_keys : (int64_t)sorted_keys[i];
        if (cur > num_keys) cur = num_keys;
        for (int64_t k = prev + 1; k <= cur; k++) offsets[k] = (uint32_t)i;
    }
}

// head flags of a sorted key array
__global__ void head_flags_kernel(const uint64_t* __restrict__ sorted_keys, int64_t n, uint32_t* __restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1u : 0u;
}

// after an exclusive scan of head flags: rank[i] = (#heads before i); unique index of position i = rank + head - 1
__global__ void unique_write_kernel(const uint64_t* __restrict__ sorted_keys, const uint32_t* __restrict__ sorted_slots,
                                    const uint32_t* __restrict__ excl, int64_t n, int64_t* __restrict__ unique_out, int64_t* __restrict__ mapped_out,
                                    const uint32_t* __restrict__ total, int64_t* __restrict__ num_unique) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool head = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]);
        uint32_t u = excl[i] + (head ? 1u : 0u) - 1u;
        if (head) unique_out[u] = (int64_t)sorted_keys[i];
        mapped_out[sorted_slots[i]] = (int64_t)u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *num_unique = (int64_t)(*total);
}

__global__ void i64_to_u64_kernel(const int64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint64_t)in[i];
}

__global__ void i64_to_u32_kernel(const int64_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint32_t)in[i];
}

inline int blocks_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

size_t sort_scratch_bytes(int64_t n) {
    int64_t tiles = (n + kTile - 1) / kTile;
    return (size_t)(256 * tiles + 1) * sizeof(uint32_t);
}

// Sorts (keys, iota slots) by key, stable.  keys_a holds the input keys (clobbered); results end in
// *keys_sorted / *vals_sorted which point at either the a or b buffers.  `key_bits` significant bits.
template <typename K>
mb_status radix_sort_pairs(K* keys_a, K* keys_b, uint32_t* vals_a, uint32_t* vals_b, int64_t n, int key_bits, uint32_t* hist_scratch, K** keys_sorted,
                           uint32_t** vals_sorted, cudaStream_t st) {
    *keys_sorted = keys_a;
    *vals_sorted = vals_a;
    if (n == 0) return MB_OK;
    if (n >= (int64_t)1 << 32) {
        set_error("radix_sort_pairs: n >= 2^32");
        return MB_ERR_UNSUPPORTED;
    }
    int tiles = (int)((n + kTile - 1) / kTile);
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    K *kin = keys_a, *kout = keys_b;
    uint32_t *vin = vals_a, *vout = vals_b;
    for (int p = 0; p < passes; p++) {
        hist_kernel<K><<<tiles, kSortThreads, 0, st>>>(kin, n, 8 * p, tiles, hist_scratch);
        MB_LAUNCH_CHECK();
        scan_kernel<<<1, 1024, 0, st>>>(hist_scratch, (int64_t)256 * tiles, nullptr);
        MB_LAUNCH_CHECK();
        scatter_kernel<K><<<tiles, kSortThreads, 0, st>>>(kin, vin, kout, vout, n, 8 * p, tiles, hist_scratch, p == 0);
        MB_LAUNCH_CHECK();
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *keys_sorted = kin;
    *vals_sorted = vin;
    return MB_OK;
}

template mb_status radix_sort_pairs<uint32_t>(uint32_t*, uint32_t*, uint32_t*, uint32_t*, int64_t, int, uint32_t*, uint32_t**, uint32_t**, cudaStream_t);
template mb_status radix_sort_pairs<uint64_t>(uint64_t*, uint64_t*, uint32_t*, uint32_t*, int64_t, int, uint32_t*, uint64_t**, uint32_t**, cudaStream_t);

mb_status launch_i64_to_u32(const int64_t* in, uint32_t* out, int64_t n, cudaStream_t st) {
    if (n == 0) return MB_OK;
    i64_to_u32_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, out, n);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status segment_offsets_u32(const uint32_t* sorted_keys, int64_t n, int64_t num_keys, uint32_t* offsets, cudaStream_t st) {
    seg_offsets_kernel<uint32_t><<<blocks_for(n + 1, 256), 256, 0, st>>>(sorted_keys, n, num_keys, offsets);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

// map_tensors on the device.  Scratch: keys_a/keys_b [n] u64, vals_a/vals_b [n] u32, flags [n] u32, hist.
mb_status map_tensors_device(const int64_t* all_ids, int64_t n, int key_bits, uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                             uint32_t* flags, uint32_t* hist_scratch, uint32_t* total_scratch, int64_t* unique_out, int64_t* mapped_out,
                             int64_t* num_unique_dev, cudaStream_t st) {
    if (n == 0) {
        MB_CUDA_TRY(cudaMemsetAsync(num_unique_dev, 0, sizeof(int64_t), st));
        return MB_OK;
    }
    i64_to_u64_kernel<<<blocks_for(n, 256), 256, 0, st>>>(all_ids, keys_a, n);
    MB_LAUNCH_CHECK();
    uint64_t* ks;
    uint32_t* vs;
    MB_TRY(radix_sort_pairs<uint64_t>(keys_a, keys_b, vals_a, vals_b, n, key_bits, hist_scratch, &ks, &vs, st));
    MB_CUDA_TRY(cudaMemsetAsync(unique_out, 0xFF, sizeof(int64_t) * n, st));  // entries past *num_unique read as -1
    head_flags_kernel<<<blocks_for(n, 256), 256, 0, st>>>(ks, n, flags);
    MB_LAUNCH_CHECK();
    scan_kernel<<<1, 1024, 0, st>>>(flags, n, total_scratch);
    MB_LAUNCH_CHECK();
    unique_write_kernel<<<blocks_for(n, 256), 256, 0, st>>>(ks, vs, flags, n, unique_out, mapped_out, total_scratch, num_unique_dev);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
