// row_bulk.cuh -- row kernels with asynchronous bulk-copy (TMA) staging: a producer warp resolves row addresses and issues one
// cp.async.bulk (global -> shared, mbarrier completion) per table row into a ring of stages; consumer warps compute from shared
// memory.  The rows in flight per SM are bounded by the ring (~100 KB per block, two blocks per SM), not by registers: the
// register-staged kernels of decoder_vec.cuh run at 12-34 % occupancy and 2.9-4.7 TB/s at the bench shape (profiles/r2_notes_mid.md),
// because each lane has to hold its in-flight bytes in registers.
#pragma once
#include "decoder_vec.cuh"
#include "gemm_tc_ptx.cuh"

namespace mb {
namespace bulk {

using tcptx::mbar_arrive;
using tcptx::mbar_expect_tx;
using tcptx::mbar_init;
using tcptx::mbar_wait;
using tcptx::smem_u32;

__device__ __forceinline__ void bulk_load_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}

// 8 fp32 -> 8 bf16 hi + 8 bf16 lo, one 16-byte global store each
__device__ __forceinline__ void store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t elem_off, const float4& x, const float4& y) {
    const float g[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(g[2 * i + 1]), "f"(g[2 * i]));
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(g[2 * i + 1] - h1), "f"(g[2 * i] - h0));
    }
    *reinterpret_cast<uint4*>(hi + elem_off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + elem_off) = make_uint4(l[0], l[1], l[2], l[3]);
}

constexpr int kStages = 4;
constexpr int kConsumers = 4;
constexpr int kThreads = 32 * (1 + kConsumers);  // 160

// dynamic shared memory of both kernels: kStages * stage_bytes + 128 (alignment) + 64 (barriers)
inline size_t smem_bytes(int rows_per_stage, int d) { return (size_t)kStages * rows_per_stage * d * 4 + 128 + 64; }

// ---------------------------------------------------------------------------------------------------------------
// negative rows: table[negs[side][j]] -> bf16 hi/lo (and / or an fp32 copy).  16 rows per stage; consumer warp w converts rows 4w .. 4w+3.
constexpr int kNegRows = 16;

template <int NCONS>  // consumer warps: each converts kNegRows / NCONS rows of a stage
__global__ void __launch_bounds__(32 * (1 + NCONS)) neg_rows_bulk_kernel(vec::PrepArgs a) {
    constexpr int RPW = kNegRows / NCONS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const int d = a.d, dv = d >> 2;
    const uint32_t row_bytes = (uint32_t)d * 4u;
    const uint32_t bar0 = base + kStages * kNegRows * row_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), NCONS);
        }
        tcptx::fence_barrier_init();
    }
    __syncthreads();
    const int64_t total = (int64_t)a.sides * a.CN;
    const int64_t chunks = (total + kNegRows - 1) / kNegRows;
    int it = 0;
    if (warp == 0) {
        // Address resolution is a chain of dependent loads (negative -> unique-id map -> row address).  It runs as a two-deep software
        // pipeline over the block's chunks: every iteration issues the first-level load of chunk c+2 and the second-level load of chunk
        // c+1, and only consumes values loaded an iteration earlier -- the warp never waits for a load it has just issued.
        const int64_t step = gridDim.x;
        auto first_level = [&](int64_t c) -> int64_t {  // the negative's batch-local id (or -1)
            const int64_t q = c * kNegRows + lane;
            if (lane >= kNegRows || c >= chunks || q >= total) return -1;
            const int side = q >= a.CN ? 1 : 0;
            return __ldg((side ? a.negs[1] : a.negs[0]) + (q - (int64_t)side * a.CN));
        };
        auto second_level = [&](int64_t nid) -> int64_t {  // what the id maps to: a global row id, a row address, or the id itself
            if (nid < 0) return -1;
            if (a.sp.world > 1) return __ldg(a.row_map + nid);
            if (a.row_ptrs != nullptr) return (int64_t)__ldg(reinterpret_cast<const unsigned long long*>(a.row_ptrs) + nid);
            return a.row_map != nullptr ? __ldg(a.row_map + nid) : nid;
        };
        auto address = [&](int64_t g) -> const float* {
            if (g < 0) return nullptr;
            if (a.sp.world > 1) {
                const int64_t o = g / a.sp.rows_per_rank;
                return a.sp.table[o] + (g - o * a.sp.rows_per_rank) * a.emb_ld;
            }
            if (a.row_ptrs != nullptr) return reinterpret_cast<const float*>(g);
            return a.emb + g * a.emb_ld;
        };
        int64_t g_cur = second_level(first_level(blockIdx.x));
        int64_t nid_next = first_level(blockIdx.x + step);
        for (int64_t c = blockIdx.x; c < chunks; c += step, it++) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)((it / kStages) & 1);
            const int64_t nid_next2 = first_level(c + 2 * step);
            const int64_t g_next = second_level(nid_next);
            const float* src = address(g_cur);
            g_cur = g_next;
            nid_next = nid_next2;
            mbar_wait(empty(s), ph ^ 1u);
            const int nvalid = (int)min((int64_t)kNegRows, total - c * kNegRows);
            if (lane == 0) mbar_expect_tx(full(s), (uint32_t)nvalid * row_bytes);
            __syncwarp();
            if (src != nullptr) bulk_load_row(base + (uint32_t)(s * kNegRows + lane) * row_bytes, src, row_bytes, full(s));
        }
    } else {
        const int w = warp - 1;
        for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x, it++) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)((it / kStages) & 1);
            mbar_wait(full(s), ph);
            const uint32_t rows0 = base + (uint32_t)(s * kNegRows + RPW * w) * row_bytes;
            const int64_t q0 = c * kNegRows + RPW * w;
            const int d8 = d >> 3;  // (d % 8 == 0: decoder_vec_ok)
            for (int t = lane; t < RPW * d8; t += 32) {  // the warp's rows as one flat list of 8-float pieces
                int r = 0;
#pragma unroll
                for (int k = 1; k < RPW; k++) r += (t >= k * d8);
                const int v = t - r * d8;
                const int64_t q = q0 + r;
                if (q >= total) break;
                const uint32_t src = rows0 + (uint32_t)r * row_bytes + (uint32_t)v * 32u;
                const float4 x = lds4(src), y = lds4(src + 16u);
                const int side = q >= a.CN ? 1 : 0;
                const int64_t j = q - (int64_t)side * a.CN;
                float* nf = side ? a.Neg[1] : a.Neg[0];
                __nv_bfloat16* nh = side ? a.Neg_hi[1] : a.Neg_hi[0];
                __nv_bfloat16* nl = side ? a.Neg_lo[1] : a.Neg_lo[0];
                if (nf) {
                    vec::st4(nf + j * d, 2 * v, x);
                    vec::st4(nf + j * d, 2 * v + 1, y);
                }
                if (nh) store_split8(nh, nl, j * d + 8 * v, x, y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(s));
        }
    }
}

// (Two further kernels built on this staging were measured and removed: the edge rows -- relation operator + positive scores per edge,
// four consumer warps per block doing a long dependent computation per item: 94 us against 70 us for the register-staged
// vec::edge_rows_kernel -- and the fetch of remote rows of the sharded table -- correct in one process and across two GPUs at small
// shapes, a launch failure at the bench shape across real NVLink peers that was not root-caused.  DESIGN.md section 9.)

}  // namespace bulk
}  // namespace mb
