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

// ---------------------------------------------------------------------------------------------------------------
// edge rows: relation operator + positive scores for both corruption sides, adjusted rows as fp32 and / or bf16 hi/lo
// (decoder_methods.cpp:74-101, relation_operators.cpp:7-35, comparators.cpp:67-68).  Four edges per stage, four rows per edge
// (src, dst, relation, inverse relation: the relation rows come from L2); consumer warp w handles edge w of the stage.
constexpr int kEdgesPerStage = 4;
constexpr int kEdgeRows = 4 * kEdgesPerStage;

template <int DEC>
__global__ void __launch_bounds__(kThreads) edge_rows_bulk_kernel(vec::PrepArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const int d = a.d, dv = d >> 2, hv = d >> 3;
    const uint32_t row_bytes = (uint32_t)d * 4u;
    const uint32_t bar0 = base + kStages * kEdgeRows * row_bytes;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool two = a.sides == 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), kConsumers);
        }
        tcptx::fence_barrier_init();
    }
    __syncthreads();
    const int64_t chunks = (a.Bp + kEdgesPerStage - 1) / kEdgesPerStage;
    const int rows_per_edge = DEC == MB_DECODER_DOT ? 2 : (two ? 4 : 3);
    int it = 0;
    if (warp == 0) {
        // lane = 4 * edge + row kind (0 src, 1 dst, 2 relation, 3 inverse relation); two-deep address pipeline as in neg_rows_bulk_kernel
        const int64_t step = gridDim.x;
        const int e = lane >> 2, kind = lane & 3;
        const bool lane_on = lane < kEdgeRows && kind < rows_per_edge;
        auto first_level = [&](int64_t c) -> int64_t {  // node id (batch-local) or relation id
            const int64_t p = c * kEdgesPerStage + e;
            if (!lane_on || c >= chunks || p >= a.B) return -1;
            return __ldg(a.edges + p * a.cols + (kind == 0 ? 0 : (kind == 1 ? a.cols - 1 : 1)));
        };
        auto second_level = [&](int64_t id) -> int64_t {
            if (id < 0 || kind >= 2) return id;
            if (a.row_ptrs != nullptr) return (int64_t)__ldg(reinterpret_cast<const unsigned long long*>(a.row_ptrs) + id);
            return a.row_map != nullptr ? __ldg(a.row_map + id) : id;
        };
        auto address = [&](int64_t g) -> const float* {
            if (g < 0) return nullptr;
            if (kind >= 2) return (kind == 2 ? a.rel : a.inv_rel) + g * d;
            if (a.row_ptrs != nullptr) return reinterpret_cast<const float*>(g);
            return a.emb + g * a.emb_ld;
        };
        int64_t g_cur = second_level(first_level(blockIdx.x));
        int64_t id_next = first_level(blockIdx.x + step);
        for (int64_t c = blockIdx.x; c < chunks; c += step, it++) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)((it / kStages) & 1);
            const int64_t id_next2 = first_level(c + 2 * step);
            const int64_t g_next = second_level(id_next);
            const float* src = address(g_cur);
            g_cur = g_next;
            id_next = id_next2;
            mbar_wait(empty(s), ph ^ 1u);
            const int nedges = (int)max((int64_t)0, min((int64_t)kEdgesPerStage, a.B - c * kEdgesPerStage));
            if (lane == 0) mbar_expect_tx(full(s), (uint32_t)(nedges * rows_per_edge) * row_bytes);
            __syncwarp();
            if (src != nullptr) bulk_load_row(base + (uint32_t)(s * kEdgeRows + lane) * row_bytes, src, row_bytes, full(s));
        }
    } else {
        const int w = warp - 1;
        for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x, it++) {
            const int s = it % kStages;
            const uint32_t ph = (uint32_t)((it / kStages) & 1);
            mbar_wait(full(s), ph);
            const int64_t p = c * kEdgesPerStage + w;
            if (p < a.Bp) {
                if (p >= a.B) {  // zero padding rows (comparators.cpp:11-15, decoder_methods.cpp:103-111)
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int sd = 0; sd < a.sides; sd++) {
                        for (int v = lane; v < dv; v += 32) {
                            if (a.A[sd]) vec::st4(a.A[sd] + p * d, v, z);
                            if (a.A_hi[sd]) vec::store_split4(a.A_hi[sd], a.A_lo[sd], p * d + 4 * v, z);
                        }
                        if (lane == 0) a.pos[sd][p] = 0.f;
                    }
                } else {
                    const uint32_t S = base + (uint32_t)(s * kEdgeRows + 4 * w) * row_bytes, Dd = S + row_bytes, R = Dd + row_bytes;
                    const uint32_t Q = two ? R + row_bytes : R;
                    float acc0 = 0.f, acc1 = 0.f;
                    if (DEC == MB_DECODER_COMPLEX) {
                        for (int v = lane; v < hv; v += 32) {
                            const uint32_t o = (uint32_t)v * 16u, oi = (uint32_t)(hv + v) * 16u;
                            const float4 sr = lds4(S + o), sim = lds4(S + oi), dr = lds4(Dd + o), dim = lds4(Dd + oi), rr = lds4(R + o), rim = lds4(R + oi);
                            const float4 ar = vec::sub4(vec::mul4(sr, rr), vec::mul4(sim, rim));    // relation_operators.cpp:31
                            const float4 ai = vec::addrn4(vec::mul4(sr, rim), vec::mul4(sim, rr));  // relation_operators.cpp:32
                            acc0 = vec::dot4(ar, dr, acc0);
                            acc0 = vec::dot4(ai, dim, acc0);
                            if (a.A[0]) {
                                vec::st4(a.A[0] + p * d, v, ar);
                                vec::st4(a.A[0] + p * d, hv + v, ai);
                            }
                            if (a.A_hi[0]) {
                                vec::store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * v, ar);
                                vec::store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * (hv + v), ai);
                            }
                            if (two) {
                                const float4 qr = lds4(Q + o), qi = lds4(Q + oi);
                                const float4 br = vec::sub4(vec::mul4(dr, qr), vec::mul4(dim, qi));
                                const float4 bi = vec::addrn4(vec::mul4(dr, qi), vec::mul4(dim, qr));
                                acc1 = vec::dot4(br, sr, acc1);
                                acc1 = vec::dot4(bi, sim, acc1);
                                if (a.A[1]) {
                                    vec::st4(a.A[1] + p * d, v, br);
                                    vec::st4(a.A[1] + p * d, hv + v, bi);
                                }
                                if (a.A_hi[1]) {
                                    vec::store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * v, br);
                                    vec::store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * (hv + v), bi);
                                }
                            }
                        }
                    } else {
                        for (int v = lane; v < dv; v += 32) {
                            const uint32_t o = (uint32_t)v * 16u;
                            const float4 sv = lds4(S + o), dvv = lds4(Dd + o);
                            const float4 av = (DEC == MB_DECODER_DISTMULT) ? vec::mul4(sv, lds4(R + o)) : sv;  // relation_operators.cpp:11
                            acc0 = vec::dot4(av, dvv, acc0);
                            if (a.A[0]) vec::st4(a.A[0] + p * d, v, av);
                            if (a.A_hi[0]) vec::store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * v, av);
                            if (DEC == MB_DECODER_DISTMULT && two) {
                                const float4 bv = vec::mul4(dvv, lds4(Q + o));
                                acc1 = vec::dot4(bv, sv, acc1);
                                if (a.A[1]) vec::st4(a.A[1] + p * d, v, bv);
                                if (a.A_hi[1]) vec::store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * v, bv);
                            }
                        }
                    }
                    acc0 = warp_sum(acc0);
                    acc1 = warp_sum(acc1);
                    if (lane == 0) {
                        a.pos[0][p] = acc0;  // comparators.cpp:67-68
                        if (two) a.pos[1][p] = acc1;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(s));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Sharded table: resolve every unique row of the batch to an address and fetch the remote ones into the batch cache (see
// vec::fetch_remote_rows_kernel).  Remote rows travel  peer HBM --NVLink--> shared memory --> local HBM  as bulk asynchronous copies in
// both directions (cp.async.bulk global -> shared with mbarrier completion, then shared -> global as a bulk group): the bytes in flight
// per SM are bounded by the ring (3 stages x 32 rows), not by what a warp can hold in registers across a multi-microsecond NVLink round trip.
constexpr int kFetchRows = 32;    // ids resolved per chunk (one per lane) = row slots per stage
constexpr int kFetchStages = 3;
constexpr int kFetchThreads = 64;  // warp 0: resolve + load, warp 1: store

inline size_t fetch_smem_bytes(int d) { return (size_t)kFetchStages * kFetchRows * d * 4 + 128 + 64 + kFetchStages * kFetchRows * 8; }

__device__ __forceinline__ void bulk_store_row(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kFetchThreads) fetch_remote_rows_bulk_kernel(vec::FetchArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const int d = a.d;
    const uint32_t row_bytes = (uint32_t)d * 4u;
    const uint32_t bar0 = base + kFetchStages * kFetchRows * row_bytes;
    const uint32_t dst0 = bar0 + 64u;  // per stage and slot: destination address of the staged row (0 = slot unused)
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kFetchStages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kFetchStages; s++) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        tcptx::fence_barrier_init();
    }
    __syncthreads();
    const int64_t chunks = (a.U + kFetchRows - 1) / kFetchRows;
    int it = 0;
    if (warp == 0) {
        for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x, it++) {
            const int s = it % kFetchStages;
            const uint32_t ph = (uint32_t)((it / kFetchStages) & 1);
            const int64_t u = c * kFetchRows + lane;
            const float* src = nullptr;
            float* dst = nullptr;
            if (u < a.U) {
                const int64_t g = __ldg(a.ids + u);
                const float* where = a.sp.table[a.sp.rank];  // padding entries point at something valid
                if (g >= 0) {
                    const int64_t o = g / a.sp.rows_per_rank, l = g - o * a.sp.rows_per_rank;
                    if (o == a.sp.rank) {
                        where = a.sp.table[o] + l * a.ld;
                    } else {
                        src = a.sp.table[o] + l * a.ld;
                        dst = a.cache + u * d;
                        where = dst;
                    }
                }
                a.row_ptrs[u] = where;
            }
            mbar_wait(empty(s), ph ^ 1u);
            const unsigned remote = __ballot_sync(0xffffffffu, src != nullptr);
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(dst0 + (uint32_t)(s * kFetchRows + lane) * 8u), "l"(reinterpret_cast<unsigned long long>(dst)) : "memory");
            __syncwarp();
            if (lane == 0) mbar_expect_tx(full(s), (uint32_t)__popc(remote) * row_bytes);  // (arrives even when the chunk has no remote row)
            __syncwarp();
            if (src != nullptr) bulk_load_row(base + (uint32_t)(s * kFetchRows + lane) * row_bytes, src, row_bytes, full(s));
        }
    } else {
        for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x, it++) {
            const int s = it % kFetchStages;
            const uint32_t ph = (uint32_t)((it / kFetchStages) & 1);
            mbar_wait(full(s), ph);
            unsigned long long dst;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(dst) : "r"(dst0 + (uint32_t)(s * kFetchRows + lane) * 8u) : "memory");
            if (dst != 0ull) bulk_store_row(reinterpret_cast<void*>(dst), base + (uint32_t)(s * kFetchRows + lane) * row_bytes, row_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stage's shared memory has been read: it may be refilled
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(s));
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the rows are in the cache before the kernel ends
    }
}

}  // namespace bulk
}  // namespace mb
