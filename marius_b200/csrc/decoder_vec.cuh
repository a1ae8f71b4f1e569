// decoder_vec.cuh -- 128-bit vectorised versions of the decoder-side row kernels (used when d % 8 == 0 and d <= 512, the
// production shapes; the scalar kernels in decoder_kernels.cu remain the general-d fallback).  One warp per row, each lane owns
// float4 column chunks (d=400 -> 100 chunks, <= 4 per lane), all loads of a row issued before first use.
//
// Index chasing is lane-parallel: a row kernel's address needs two or three dependent global loads (edge / negative index ->
// unique-id map -> table row), and a warp that walks that chain once per row spends most of its time with nothing in flight.
// Instead lane k resolves the row pointers of the k-th of the warp's next 32 work items (items are dealt round-robin to the warps,
// so the chains of up to 32 items run side by side), and the row loop broadcasts them with shuffles: the chain is paid once per warp.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mb {
namespace vec {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

__device__ __forceinline__ float4 ld4(const float* p, int v) { return __ldg(reinterpret_cast<const float4*>(p) + v); }
__device__ __forceinline__ void st4(float* p, int v, const float4& x) { reinterpret_cast<float4*>(p)[v] = x; }
__device__ __forceinline__ float4 mul4(const float4& a, const float4& b) {
    return make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) {
    return make_float4(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z), __fsub_rn(a.w, b.w));
}
__device__ __forceinline__ float4 addrn4(const float4& a, const float4& b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 scale4(float s, const float4& a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 fma4(float s, const float4& a, const float4& c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ float4 neg4(const float4& a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
// Row loads of the lane-parallel kernels.  The pointers come out of shuffles (the compiler no longer knows they are global), and a lane
// whose chunk index is past the row end must not branch around its load (a branch per chunk keeps the compiler from hoisting the
// loads of a row above the first use, which serialises them): the chunk index is clamped instead, the lane re-reads the row's last
// chunk (same cache line as its neighbours) and simply does not store.
__device__ __forceinline__ float4 ldg_nc4(const float* p, int v) {  // read-only data (table rows during the forward, relation rows)
    float4 r;
    asm("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(reinterpret_cast<const float4*>(p) + v));
    return r;
}
__device__ __forceinline__ float4 ldg4(const float* p, int v) {  // data written earlier in the step / rewritten by this kernel
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(reinterpret_cast<const float4*>(p) + v));
    return r;
}
// fire-and-forget 128-bit add at system scope (the target may be a peer GPU's HBM): one NVLink write, no read, no reply
__device__ __forceinline__ void red_add4_sys(float4* p, const float4& v) {
    asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <typename T>
__device__ __forceinline__ T* shfl_ptr(T* p, int src_lane) {
    return reinterpret_cast<T*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(p), src_lane));
}
__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src_lane) { return (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)v, src_lane); }

// Node-partition-sharded table: rank o owns global rows [o * rows_per_rank, (o+1) * rows_per_rank).  table[o] / state[o]
// are the owners' base pointers -- local HBM for o == this rank, peer HBM mapped over NVLink (CUDA IPC) otherwise -- so the same fused
// kernels gather remote rows with plain loads and apply the Adagrad read-modify-write with plain stores: no staging, no collective.
struct ShardPtrs {
    float* table[8];
    float* state[8];
    int64_t rows_per_rank;
    int world;  // <= 1: unsharded, `emb` / `table` arguments are used directly
    int rank;   // the shard that is this GPU's own HBM
};

// where batch-local row `id` lives: the sharded step's per-batch pointer table, the table through the unique-id map, or the batch matrix
template <typename Args>
__device__ __forceinline__ const float* batch_row(const Args& a, int64_t id) {
    if (a.row_ptrs != nullptr) return reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(a.row_ptrs) + id));
    return a.emb + (a.row_map != nullptr ? __ldg(a.row_map + id) : id) * a.emb_ld;
}

// Sharded table, start of the step: resolve every unique row of the batch to an address and fetch the remote ones.
//   row_ptrs[u] = own HBM row                      when this rank owns global row ids[u]
//               = cache + u * d (filled here)      otherwise: the row is copied once from the owner's HBM over NVLink
// One copy kernel with two rows per warp in flight and nothing else in its registers keeps far more remote rows in flight per SM
// than the decoder kernels can (NVLink round trips are several microseconds), and they then read local memory only.
struct FetchArgs {
    const int64_t* ids;  // [U] global row ids (negative = padding)
    int64_t U;
    ShardPtrs sp;
    int64_t ld;
    int d;
    float* cache;            // [U, d]
    const float** row_ptrs;  // [U]  (embedding fetch only)
};

// STATE = false: embedding rows (+ row_ptrs).  STATE = true: the Adagrad state rows of the same remote rows, fetched on a side stream
// while the contractions run, so that the update at the end of the step reads them from local memory.
template <int CH, bool STATE>
__global__ void __launch_bounds__(kThreads) fetch_remote_rows_kernel(FetchArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int dv = a.d >> 2;
    for (int64_t first = warp0; first < a.U; first += 32 * nwarps) {
        const float* my_src = nullptr;
        float* my_dst = nullptr;
        {
            const int64_t u = first + lane * nwarps;
            if (u < a.U) {
                const int64_t g = __ldg(a.ids + u);
                const float* where = a.sp.table[a.sp.rank];  // padding entries point at something valid
                if (g >= 0) {
                    const int64_t o = g / a.sp.rows_per_rank, l = g - o * a.sp.rows_per_rank;
                    if (o == a.sp.rank) {
                        where = a.sp.table[o] + l * a.ld;
                    } else {
                        my_src = (STATE ? a.sp.state[o] : a.sp.table[o]) + l * a.ld;
                        my_dst = a.cache + u * a.d;
                        where = my_dst;
                    }
                }
                if (!STATE) a.row_ptrs[u] = where;
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, my_src != nullptr);
        while (todo != 0) {  // remote rows of this round, four at a time (an NVLink round trip is several microseconds: keep bytes in flight)
            int k[4];
            const float* r[4];
            float* w[4];
            int n = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                k[q] = todo != 0 ? __ffs(todo) - 1 : k[0];
                if (todo != 0) n++;
                todo &= todo - 1;  // (no-op once todo is 0)
                r[q] = shfl_ptr(my_src, k[q]);
                w[q] = shfl_ptr(my_dst, k[q]);
            }
            float4 x[4][CH];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int vc = min(lane + 32 * c, dv - 1);
#pragma unroll
                for (int q = 0; q < 4; q++) x[q][c] = ldg4(r[q], vc);  // (plain loads: the peer may have rewritten the row in an earlier step)
            }
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int v = lane + 32 * c;
                if (v < dv) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (q < n) st4(w[q], v, x[q][c]);
                }
            }
        }
    }
}

// x -> (hi, lo) bf16 with x ~= hi + lo ; 4 elements packed into two 8-byte stores
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, int64_t elem_off, const float4& x) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y), h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
    __nv_bfloat16 l0 = __float2bfloat16_rn(x.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x.y - __bfloat162float(h1));
    __nv_bfloat16 l2 = __float2bfloat16_rn(x.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(x.w - __bfloat162float(h3));
    uint2 hp, lp;
    hp.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    hp.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
    lp.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    lp.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
    *reinterpret_cast<uint2*>(hi + elem_off) = hp;
    *reinterpret_cast<uint2*>(lo + elem_off) = lp;
}

// ---------------------------------------------------------------------------------------------------------------
// prep: edge rows (relation operator + positive scores, both corruption sides) and negative-row gather/split, one launch each.
struct PrepArgs {
    const float* emb;
    int64_t emb_ld;
    const int64_t* row_map;  // null: emb is the batch-local matrix; else emb is the table and row_map = unique ids (gather fused away)
    const float* const* row_ptrs;  // sharded table: address of every unique row (own HBM, or the batch's cache of fetched remote rows)
    ShardPtrs sp;                  // sharded table: the negative-row kernel runs BESIDE the remote-row fetch, so it resolves rows through the
                                   // owners' tables directly (negatives are drawn from the resident partitions, i.e. almost always local)
    const int64_t* edges;
    int cols;
    const float* rel;      // null: no relation operator
    const float* inv_rel;  // null: no inverse side
    int64_t B, Bp, CN;
    int d, decoder, sides;
    const int64_t* negs[2];     // dst_negs, src_negs (flattened [C*N])
    float* A[2];                // adjusted rows fp32 [Bp,d] per side (null: not needed)
    float* pos[2];              // [Bp]
    __nv_bfloat16 *A_hi[2], *A_lo[2];      // [Bp,d] per side or null
    float* Neg[2];              // fp32 negative rows [CN,d] per side or null
    __nv_bfloat16 *Neg_hi[2], *Neg_lo[2];  // or null
};

// negative rows: emb[negs[side][j]] -> fp32 copy and/or bf16 hi/lo.  Items = the rows of dst_negs then src_negs, dealt round-robin to the
// warps; two rows in flight per lane.
template <int CH>  // float4 chunks per lane: d <= 128 * CH
__global__ void __launch_bounds__(kThreads) neg_rows_kernel(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, dv = d >> 2;
    const int64_t total = (int64_t)a.sides * a.CN;
    // the warp owns items warp0, warp0 + nwarps, ... ; lane k resolves the k-th of the next 32 of them
    for (int64_t first = warp0; first < total; first += 32 * nwarps) {
        const float* mine = nullptr;
        {
            const int64_t q = first + lane * nwarps;
            if (q < total) {
                const int side = q >= a.CN ? 1 : 0;
                const int64_t nid = __ldg((side ? a.negs[1] : a.negs[0]) + (q - (int64_t)side * a.CN));
                if (a.sp.world > 1) {
                    const int64_t g = __ldg(a.row_map + nid), o = g / a.sp.rows_per_rank;
                    mine = a.sp.table[o] + (g - o * a.sp.rows_per_rank) * a.emb_ld;
                } else {
                    mine = batch_row(a, nid);
                }
            }
        }
#pragma unroll 1
        for (int k = 0; k < 32; k += 2) {
            if (first + k * nwarps >= total) break;
            const float* r0 = shfl_ptr(mine, k);
            const float* r1 = shfl_ptr(mine, k + 1);
            const float* q0 = r0 ? r0 : a.emb;  // past-the-end rows of the last step: load something valid, store nothing
            const float* q1 = r1 ? r1 : a.emb;
            float4 x0[CH], x1[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int vc = min(lane + 32 * c, dv - 1);
                x0[c] = ldg_nc4(q0, vc);
                x1[c] = ldg_nc4(q1, vc);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float* r = h ? r1 : r0;
                if (r == nullptr) continue;
                const int64_t q = first + (k + h) * nwarps;
                const int side = q >= a.CN ? 1 : 0;
                const int64_t j = q - (int64_t)side * a.CN;
                float* nf = side ? a.Neg[1] : a.Neg[0];
                __nv_bfloat16* nh = side ? a.Neg_hi[1] : a.Neg_hi[0];
                __nv_bfloat16* nl = side ? a.Neg_lo[1] : a.Neg_lo[0];
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int v = lane + 32 * c;
                    if (v >= dv) continue;
                    const float4 x = h ? x1[c] : x0[c];
                    if (nf) st4(nf + j * d, v, x);
                    if (nh) store_split4(nh, nl, j * d + 4 * v, x);
                }
            }
        }
    }
}

// edge rows: relation operator + positive scores for both corruption sides, adjusted rows as fp32 and/or bf16 hi/lo.  Items = the Bp
// (padded) positives, dealt round-robin to the warps.  CHV = float4 chunks per lane of a full row (DOT / DistMult) or of a complex half (ComplEx).
template <int DEC, int CHV>  // MB_DECODER_*: relation operator fixed at compile time (DOT == identity)
__global__ void __launch_bounds__(kThreads) edge_rows_kernel(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, dv = d >> 2, hv = d >> 3;  // float4 chunks per row / per complex half
    const bool two = a.sides == 2;
    for (int64_t first = warp0; first < a.Bp; first += 32 * nwarps) {
        const float *my_src = nullptr, *my_dst = nullptr;
        int64_t my_rid = 0;
        {
            const int64_t p = first + lane * nwarps;
            if (p < a.B) {
                const int64_t si = __ldg(a.edges + p * a.cols), ti = __ldg(a.edges + p * a.cols + a.cols - 1);
                if (DEC != MB_DECODER_DOT) my_rid = __ldg(a.edges + p * a.cols + 1);
                my_src = batch_row(a, si);
                my_dst = batch_row(a, ti);
            }
        }
#pragma unroll 1
        for (int k = 0; k < 32; k++) {
            const int64_t p = first + k * nwarps;
            if (p >= a.Bp) break;
            const float* src = shfl_ptr(my_src, k);
            const float* dst = shfl_ptr(my_dst, k);
            const int64_t rid = shfl_i64(my_rid, k);
            if (p >= a.B) {  // zero padding rows (comparators.cpp:11-15, decoder_methods.cpp:103-111)
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s = 0; s < a.sides; s++) {
                    for (int v = lane; v < dv; v += 32) {
                        if (a.A[s]) st4(a.A[s] + p * d, v, z);
                        if (a.A_hi[s]) store_split4(a.A_hi[s], a.A_lo[s], p * d + 4 * v, z);
                    }
                    if (lane == 0) a.pos[s][p] = 0.f;
                }
                continue;
            }
            const float* r = (DEC != MB_DECODER_DOT) ? a.rel + rid * d : nullptr;
            const float* ri = (DEC != MB_DECODER_DOT && two) ? a.inv_rel + rid * d : nullptr;
            float acc0 = 0.f, acc1 = 0.f;
            if (DEC == MB_DECODER_COMPLEX) {
                float4 sr[CHV], sim[CHV], dr[CHV], dim[CHV], rr[CHV], rim[CHV], qr[CHV], qi[CHV];
                const float* riq = two ? ri : r;
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int vc = min(lane + 32 * c, hv - 1);
                    sr[c] = ldg_nc4(src, vc), sim[c] = ldg_nc4(src, hv + vc), dr[c] = ldg_nc4(dst, vc), dim[c] = ldg_nc4(dst, hv + vc);
                    rr[c] = ldg_nc4(r, vc), rim[c] = ldg_nc4(r, hv + vc);
                    qr[c] = ldg_nc4(riq, vc), qi[c] = ldg_nc4(riq, hv + vc);
                }
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int v = lane + 32 * c;
                    if (v >= hv) continue;
                    float4 ar = sub4(mul4(sr[c], rr[c]), mul4(sim[c], rim[c]));    // relation_operators.cpp:31
                    float4 ai = addrn4(mul4(sr[c], rim[c]), mul4(sim[c], rr[c]));  // relation_operators.cpp:32
                    acc0 = dot4(ar, dr[c], acc0);
                    acc0 = dot4(ai, dim[c], acc0);
                    if (a.A[0]) {
                        st4(a.A[0] + p * d, v, ar);
                        st4(a.A[0] + p * d, hv + v, ai);
                    }
                    if (a.A_hi[0]) {
                        store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * v, ar);
                        store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * (hv + v), ai);
                    }
                    if (two) {
                        float4 br = sub4(mul4(dr[c], qr[c]), mul4(dim[c], qi[c]));
                        float4 bi = addrn4(mul4(dr[c], qi[c]), mul4(dim[c], qr[c]));
                        acc1 = dot4(br, sr[c], acc1);
                        acc1 = dot4(bi, sim[c], acc1);
                        if (a.A[1]) {
                            st4(a.A[1] + p * d, v, br);
                            st4(a.A[1] + p * d, hv + v, bi);
                        }
                        if (a.A_hi[1]) {
                            store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * v, br);
                            store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * (hv + v), bi);
                        }
                    }
                }
            } else {
                float4 sv[CHV], dvv[CHV], rv[CHV], qv[CHV];
                const float* riq = two ? ri : r;
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int vc = min(lane + 32 * c, dv - 1);
                    sv[c] = ldg_nc4(src, vc), dvv[c] = ldg_nc4(dst, vc);
                    if (DEC == MB_DECODER_DISTMULT) rv[c] = ldg_nc4(r, vc), qv[c] = ldg_nc4(riq, vc);
                }
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int v = lane + 32 * c;
                    if (v >= dv) continue;
                    float4 av = (DEC == MB_DECODER_DISTMULT) ? mul4(sv[c], rv[c]) : sv[c];  // relation_operators.cpp:11
                    acc0 = dot4(av, dvv[c], acc0);
                    if (a.A[0]) st4(a.A[0] + p * d, v, av);
                    if (a.A_hi[0]) store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * v, av);
                    if (DEC == MB_DECODER_DISTMULT && two) {
                        float4 bv = mul4(dvv[c], qv[c]);
                        acc1 = dot4(bv, sv[c], acc1);
                        if (a.A[1]) st4(a.A[1] + p * d, v, bv);
                        if (a.A_hi[1]) store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * v, bv);
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
}

// ---------------------------------------------------------------------------------------------------------------
// SoftmaxCrossEntropy forward + gradient, the whole score row held in registers (N <= 32 * 4 * MAXV).
struct LossVArgs {
    const float* S;          // [rows, N] scores
    float* G;                // [rows, N] fp32 gradient out (may alias S; null: not needed)
    const float* pos;
    float* gpos;
    float* row_loss;
    __nv_bfloat16 *G_hi, *G_lo;  // or null
    int64_t rows;
    int N;
    float w;
    int64_t ldg;  // leading dimension of G_hi / G_lo
};

template <int MAXV>
__global__ void __launch_bounds__(kThreads) loss_kernel(LossVArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int nv = a.N >> 2;
    for (int64_t i = warp0; i < a.rows; i += nwarps) {
        const float* s = a.S + i * a.N;
        float4 x[MAXV];
#pragma unroll
        for (int c = 0; c < MAXV; c++) {
            int v = lane + 32 * c;
            x[c] = v < nv ? ld_f4(reinterpret_cast<const float4*>(s) + v) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        const float p = a.pos[i];
        float m = p;
#pragma unroll
        for (int c = 0; c < MAXV; c++) m = fmaxf(m, fmaxf(fmaxf(x[c].x, x[c].y), fmaxf(x[c].z, x[c].w)));
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < MAXV; c++) {
            if (lane + 32 * c < nv) sum += expf(x[c].x - m) + expf(x[c].y - m) + expf(x[c].z - m) + expf(x[c].w - m);
        }
        sum = warp_sum(sum);
        sum += expf(p - m);
        const float z = m + logf(sum);  // log(e^pos + sum_j e^neg_j)        loss.cpp:57-66
#pragma unroll
        for (int c = 0; c < MAXV; c++) {
            int v = lane + 32 * c;
            if (v < nv) {
                float4 g = make_float4(expf(x[c].x - z) * a.w, expf(x[c].y - z) * a.w, expf(x[c].z - z) * a.w, expf(x[c].w - z) * a.w);
                if (a.G) st4(a.G + i * a.N, v, g);
                if (a.G_hi) store_split4(a.G_hi, a.G_lo, i * a.ldg + 4 * v, g);
            }
        }
        if (lane == 0) {
            a.gpos[i] = (expf(p - z) - 1.0f) * a.w;
            a.row_loss[i] = (z - p) * a.w;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// edge backward: chain rule through the positive dot and the relation operator; a / b are recomputed from src, dst, r.
struct EdgeBwdVArgs {
    const float* emb;
    int64_t emb_ld;
    const int64_t* row_map;
    const float* const* row_ptrs;
    const int64_t* edges;
    int cols;
    const float* rel;
    const float* inv_rel;
    int64_t B, Bp;
    int d, sides;
    const float* dA[2];    // G.Neg per side [Bp,d]
    const float* gpos[2];  // [Bp]
    float* gcat;           // rows [0,B) d src, [B,2B) d dst
    float* drel[2];        // [B,d] per side or null
};

template <int DEC, int CHV>  // CHV: float4 chunks per lane of a full row (DOT / DistMult) or of a complex half (ComplEx)
__global__ void __launch_bounds__(kThreads) edge_backward_kernel(EdgeBwdVArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, dv = d >> 2, hv = d >> 3;
    const bool inverse = a.sides == 2;
    for (int64_t first = warp0; first < a.B; first += 32 * nwarps) {
        const float *my_src = nullptr, *my_dst = nullptr;
        int64_t my_rid = 0;
        float my_g0 = 0.f, my_g1 = 0.f;
        {
            const int64_t i = first + lane * nwarps;
            if (i < a.B) {
                const int64_t si = __ldg(a.edges + i * a.cols), ti = __ldg(a.edges + i * a.cols + a.cols - 1);
                if (DEC != MB_DECODER_DOT) my_rid = __ldg(a.edges + i * a.cols + 1);
                my_src = batch_row(a, si);
                my_dst = batch_row(a, ti);
                my_g0 = a.gpos[0][i];
                if (inverse) my_g1 = a.gpos[1][i];
            }
        }
#pragma unroll 1
        for (int k = 0; k < 32; k++) {
            const int64_t i = first + k * nwarps;
            if (i >= a.B) break;
            const float* src = shfl_ptr(my_src, k);
            const float* dst = shfl_ptr(my_dst, k);
            const int64_t rid = shfl_i64(my_rid, k);
            const float g0 = __shfl_sync(0xffffffffu, my_g0, k), g1 = __shfl_sync(0xffffffffu, my_g1, k);
            const float* r = (DEC != MB_DECODER_DOT) ? a.rel + rid * d : nullptr;
            const float* ri = (DEC != MB_DECODER_DOT && inverse) ? a.inv_rel + rid * d : nullptr;
            const float* da0 = a.dA[0] + i * d;
            const float* da1 = inverse ? a.dA[1] + i * d : nullptr;
            float* dsrc = a.gcat + i * d;
            float* ddst = a.gcat + (a.B + i) * d;
            if (DEC == MB_DECODER_COMPLEX) {
                float4 sr[CHV], sim[CHV], dr[CHV], dim[CHV], rr[CHV], rim[CHV], qr[CHV], qi[CHV], a0r[CHV], a0i[CHV], a1r[CHV], a1i[CHV];
                const float* riq = inverse ? ri : r;
                const float* da1q = inverse ? da1 : da0;
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int vc = min(lane + 32 * c, hv - 1);
                    sr[c] = ldg_nc4(src, vc), sim[c] = ldg_nc4(src, hv + vc), dr[c] = ldg_nc4(dst, vc), dim[c] = ldg_nc4(dst, hv + vc);
                    rr[c] = ldg_nc4(r, vc), rim[c] = ldg_nc4(r, hv + vc);
                    a0r[c] = ldg4(da0, vc), a0i[c] = ldg4(da0, hv + vc);
                    qr[c] = ldg_nc4(riq, vc), qi[c] = ldg_nc4(riq, hv + vc), a1r[c] = ldg4(da1q, vc), a1i[c] = ldg4(da1q, hv + vc);
                }
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int v = lane + 32 * c;
                    if (v >= hv) continue;
                    float4 ar = sub4(mul4(sr[c], rr[c]), mul4(sim[c], rim[c])), ai = addrn4(mul4(sr[c], rim[c]), mul4(sim[c], rr[c]));
                    float4 gar = fma4(g0, dr[c], a0r[c]), gai = fma4(g0, dim[c], a0i[c]);  // d/da: bmm backward + pos dot
                    float4 dsr = add4(mul4(gar, rr[c]), mul4(gai, rim[c]));
                    float4 dsi = sub4(mul4(gai, rr[c]), mul4(gar, rim[c]));
                    if (a.drel[0]) {
                        st4(a.drel[0] + i * d, v, add4(mul4(gar, sr[c]), mul4(gai, sim[c])));
                        st4(a.drel[0] + i * d, hv + v, sub4(mul4(gai, sr[c]), mul4(gar, sim[c])));
                    }
                    float4 ddr = scale4(g0, ar), ddi = scale4(g0, ai);  // pos = <a, dst>
                    if (inverse) {
                        float4 br = sub4(mul4(dr[c], qr[c]), mul4(dim[c], qi[c])), bi = addrn4(mul4(dr[c], qi[c]), mul4(dim[c], qr[c]));
                        float4 gbr = fma4(g1, sr[c], a1r[c]), gbi = fma4(g1, sim[c], a1i[c]);
                        ddr = add4(ddr, add4(mul4(gbr, qr[c]), mul4(gbi, qi[c])));
                        ddi = add4(ddi, sub4(mul4(gbi, qr[c]), mul4(gbr, qi[c])));
                        if (a.drel[1]) {
                            st4(a.drel[1] + i * d, v, add4(mul4(gbr, dr[c]), mul4(gbi, dim[c])));
                            st4(a.drel[1] + i * d, hv + v, sub4(mul4(gbi, dr[c]), mul4(gbr, dim[c])));
                        }
                        dsr = fma4(g1, br, dsr);  // inv_pos = <b, src>
                        dsi = fma4(g1, bi, dsi);
                    }
                    st4(dsrc, v, dsr);
                    st4(dsrc, hv + v, dsi);
                    st4(ddst, v, ddr);
                    st4(ddst, hv + v, ddi);
                }
            } else {
                float4 sv[CHV], dvv[CHV], rv[CHV], qv[CHV], a0[CHV], a1[CHV];
                const float* da1q = inverse ? da1 : da0;
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int vc = min(lane + 32 * c, dv - 1);
                    sv[c] = ldg_nc4(src, vc), dvv[c] = ldg_nc4(dst, vc), a0[c] = ldg4(da0, vc), a1[c] = ldg4(da1q, vc);
                    rv[c] = (DEC == MB_DECODER_DISTMULT) ? ldg_nc4(r, vc) : make_float4(1.f, 1.f, 1.f, 1.f);
                    qv[c] = (DEC == MB_DECODER_DISTMULT && inverse) ? ldg_nc4(ri, vc) : make_float4(1.f, 1.f, 1.f, 1.f);
                }
#pragma unroll
                for (int c = 0; c < CHV; c++) {
                    const int v = lane + 32 * c;
                    if (v >= dv) continue;
                    float4 av = (DEC == MB_DECODER_DISTMULT) ? mul4(sv[c], rv[c]) : sv[c];
                    float4 ga = fma4(g0, dvv[c], a0[c]);
                    float4 ds = (DEC == MB_DECODER_DISTMULT) ? mul4(ga, rv[c]) : ga;
                    if (a.drel[0]) st4(a.drel[0] + i * d, v, mul4(ga, sv[c]));
                    float4 dd = scale4(g0, av);
                    if (inverse) {
                        float4 bv = mul4(dvv[c], qv[c]);
                        float4 gb = fma4(g1, sv[c], a1[c]);
                        dd = add4(dd, mul4(gb, qv[c]));
                        if (a.drel[1]) st4(a.drel[1] + i * d, v, mul4(gb, dvv[c]));
                        ds = fma4(g1, bv, ds);
                    }
                    st4(dsrc, v, ds);
                    st4(ddst, v, dd);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// segmented row sum over sorted slot lists (+ Adagrad), float4, d <= 4*32*CH.
struct SegVArgs {
    const float* rows;
    const uint32_t* slots;
    const uint32_t* offsets;
    int64_t n_seg;
    int d;
    float* out;
    int64_t out_ld;
    const float* state;
    int64_t state_ld;
    float *delta_e, *delta_s;
    float *table, *state_table;
    int64_t ld;
    const int64_t* ids;
    float neg_lr;
    ShardPtrs sp;
    // sharded: the summed gradient row of a remote unique row u goes to slot u - owner_bounds[owner] of this rank's inbox at the owner
    const int64_t* owner_bounds;  // [world + 1] positions in the sorted unique-id list (owner_bounds_kernel)
    int64_t* inbox_ids[8];
    float* inbox_rows[8];
    int64_t inbox_cap;
    int part;  // sharded update: 0 = every row, 1 = only the rows this rank owns, 2 = only remote rows (their gradient rows cross NVLink)
};

__device__ __forceinline__ void adagrad4(const float4& g, const float4& s, float neg_lr, float4& de, float4& ds, float4& sn) {
    adagrad_rule(g.x, s.x, neg_lr, de.x, ds.x, sn.x);
    adagrad_rule(g.y, s.y, neg_lr, de.y, ds.y, sn.y);
    adagrad_rule(g.z, s.z, neg_lr, de.z, ds.z, sn.z);
    adagrad_rule(g.w, s.w, neg_lr, de.w, ds.w, sn.w);
}

template <int MODE, int CH>
__global__ void __launch_bounds__(kThreads) segment_reduce_kernel(SegVArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, dv = d >> 2;
    // the warp owns segments warp0, warp0 + nwarps, ... ; lane k resolves bounds, table rows and first gradient row of the k-th of the next 32
    for (int64_t first = warp0; first < a.n_seg; first += 32 * nwarps) {
        uint32_t my_beg = 0, my_end = 0;
        float *my_e = nullptr, *my_s = nullptr;
        const float* my_row = nullptr;
        int my_remote = 0;
        {
            const int64_t u = first + lane * nwarps;
            if (u < a.n_seg) {
                my_beg = __ldg(a.offsets + u);
                my_end = __ldg(a.offsets + u + 1);
                if (MODE == 2 && my_end > my_beg) {  // (empty = padding segment: graph replay runs with n_seg = capacity; no row, nothing to update)
                    const int64_t r = __ldg(a.ids + u);
                    if (a.sp.world <= 1) {
                        my_e = a.table + r * a.ld;
                        my_s = a.state_table + r * a.ld;
                    } else {  // the owner's HBM: peer-mapped over NVLink when the owner is another rank
                        const int64_t o = r / a.sp.rows_per_rank, lr_ = r - o * a.sp.rows_per_rank;
                        if (o == a.sp.rank) {
                            if (a.part != 2) {
                                my_e = a.sp.table[o] + lr_ * a.ld;
                                my_s = a.sp.state[o] + lr_ * a.ld;
                            }
                        } else if (a.part != 1) {
                            // remote row: only its gradient row crosses NVLink, into this rank's inbox at the owner (the owner applies Adagrad)
                            const int64_t slot = u - __ldg(a.owner_bounds + o);
                            my_remote = 1;
                            if (slot < a.inbox_cap) {
                                my_e = a.inbox_rows[o] + slot * d;
                                a.inbox_ids[o][slot] = r;
                            }
                        }
                    }
                }
                if (my_end > my_beg) my_row = a.rows + (int64_t)__ldg(a.slots + my_beg) * d;
            }
        }
#pragma unroll 1
        for (int k = 0; k < 32; k++) {
            const int64_t u = first + k * nwarps;
            if (u >= a.n_seg) break;
            const uint32_t beg = __shfl_sync(0xffffffffu, my_beg, k), end = __shfl_sync(0xffffffffu, my_end, k);
            float* erow = shfl_ptr(my_e, k);
            float* srow = shfl_ptr(my_s, k);
            const float* row0 = shfl_ptr(my_row, k);
            const bool remote = MODE == 2 && __shfl_sync(0xffffffffu, my_remote, k) != 0;
            if (MODE == 2 && (beg == end || erow == nullptr)) continue;
            const float* sread = remote ? a.rows : srow;  // (remote rows read no state: something valid, the value is unused)
            float4 e[CH], s[CH], acc[CH];
            const float* g0 = row0 ? row0 : a.rows;  // empty segment (modes 0 / 1): load something valid, add nothing
#pragma unroll
            for (int c = 0; c < CH; c++) {  // table row, state row and first gradient row: 3 * CH independent loads in flight per lane
                const int vc = min(lane + 32 * c, dv - 1);
                if (MODE == 2) {
                    if (!remote) e[c] = ldg4(erow, vc);  // warp-uniform predicate, not a branch: the loads below still issue back to back
                    s[c] = ldg4(sread, vc);
                } else if (MODE == 1) {
                    s[c] = ldg_nc4(a.state + u * a.state_ld, vc);
                }
                acc[c] = ldg4(g0, vc);
            }
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = row0 ? addrn4(make_float4(0.f, 0.f, 0.f, 0.f), acc[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t q = beg + 1; q < end; q++) {  // duplicates: remaining gradient rows in slot order
                const float* row = a.rows + (int64_t)a.slots[q] * d;
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int v = lane + 32 * c;
                    if (v < dv) acc[c] = addrn4(acc[c], ld_f4(reinterpret_cast<const float4*>(row) + v));
                }
            }
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int v = lane + 32 * c;
                if (v >= dv) continue;
                if (MODE == 0) {
                    st4(a.out + u * a.out_ld, v, acc[c]);
                } else if (MODE == 1) {
                    if (a.out) st4(a.out + u * a.out_ld, v, acc[c]);
                    float4 de, ds, sn;
                    adagrad4(acc[c], s[c], a.neg_lr, de, ds, sn);
                    st4(a.delta_e + u * d, v, de);
                    st4(a.delta_s + u * d, v, ds);
                } else {
                    float4 de, ds, sn;
                    adagrad4(acc[c], s[c], a.neg_lr, de, ds, sn);
                    if (remote) {
                        st4(erow, v, acc[c]);  // gradient row -> inbox slot in the owner's HBM (plain 128-bit stores over NVLink)
                    } else {
                        st_stream(reinterpret_cast<float4*>(erow) + v, addrn4(e[c], de));
                        st_stream(reinterpret_cast<float4*>(srow) + v, sn);
                    }
                }
            }
        }
    }
}

// Plain segmented row sum with the columns split over warps: warp = (segment, 32-float4 column block); blockIdx.y selects one of
// up to two (rows, out) pairs (the two relation tables).  Used where segments are few and long (relation gradients).
struct SegColArgs {
    const float* rows[2];
    float* out[2];
    const uint32_t* slots;
    const uint32_t* offsets;
    int64_t n_seg;
    int d;
    int64_t out_ld;
};

__global__ void __launch_bounds__(kThreads) segment_colsplit_kernel(SegColArgs a) {
    const int lane = threadIdx.x & 31;
    const int dv = a.d >> 2;
    const int ncb = (dv + 31) >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const float* rows = a.rows[blockIdx.y];
    float* out = a.out[blockIdx.y];
    for (int64_t w = warp0; w < a.n_seg * ncb; w += nwarps) {
        const int64_t u = w / ncb;
        const int v = (int)(w - u * ncb) * 32 + lane;
        const uint32_t beg = a.offsets[u], end = a.offsets[u + 1];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < dv) {
            uint32_t q = beg;
            for (; q + 1 < end; q += 2) {  // two rows in flight
                float4 x0 = ld_f4(reinterpret_cast<const float4*>(rows + (int64_t)a.slots[q] * a.d) + v);
                float4 x1 = ld_f4(reinterpret_cast<const float4*>(rows + (int64_t)a.slots[q + 1] * a.d) + v);
                acc = addrn4(addrn4(acc, x0), x1);
            }
            if (q < end) acc = addrn4(acc, ld_f4(reinterpret_cast<const float4*>(rows + (int64_t)a.slots[q] * a.d) + v));
            st4(out + u * a.out_ld, v, acc);
        }
    }
}

}  // namespace vec
}  // namespace mb
