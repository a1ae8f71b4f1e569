// decoder_vec.cuh -- 128-bit vectorised versions of the decoder-side row kernels (used when d % 8 == 0, the production
// shapes; the scalar kernels in decoder_kernels.cu remain the general-d fallback).  One warp per row, each lane owns
// float4 column chunks: d=400 -> 100 chunks, <= 4 per lane, all loads of a row issued before first use.
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

// Node-partition-sharded table: rank o owns global rows [o * rows_per_rank, (o+1) * rows_per_rank).  table[o] / state[o]
// are the owners' base pointers -- local HBM for o == this rank, peer HBM mapped over NVLink (CUDA IPC) otherwise -- so the same fused
// kernels gather remote rows with plain loads and apply the Adagrad read-modify-write with plain stores: no staging, no collective.
struct ShardPtrs {
    float* table[8];
    float* state[8];
    int64_t rows_per_rank;
    int world;  // <= 1: unsharded, `emb` / `table` arguments are used directly
};

__device__ __forceinline__ const float* shard_row(const ShardPtrs& sp, const float* emb, int64_t emb_ld, int64_t g) {
    if (sp.world <= 1) return emb + g * emb_ld;
    const int64_t o = g / sp.rows_per_rank;
    return sp.table[o] + (g - o * sp.rows_per_rank) * emb_ld;
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
// prep: edge rows (relation operator + positive scores, both corruption sides) AND negative-row gather/split in one launch.
struct PrepArgs {
    const float* emb;
    int64_t emb_ld;
    const int64_t* row_map;  // null: emb is the batch-local matrix; else emb is the table and row_map = unique ids (gather fused away)
    ShardPtrs sp;            // with row_map: where global row g lives
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

template <int DEC>  // MB_DECODER_*: relation operator fixed at compile time (DOT == identity)
__global__ void __launch_bounds__(kThreads) prep_kernel(PrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, dv = d >> 2, hv = d >> 3;  // float4 chunks per row / per complex half
    const int64_t total = a.Bp + (int64_t)a.sides * a.CN;
    for (int64_t t = warp0; t < total; t += nwarps) {
        if (t >= a.Bp) {
            // ---- negative row: emb[negs[side][j]] -> fp32 copy and/or bf16 hi/lo
            const int64_t q = t - a.Bp;
            const int side = q >= a.CN ? 1 : 0;
            const int64_t j = q - (int64_t)side * a.CN;
            const int64_t nid = a.negs[side][j];
            const float* src = a.row_map ? shard_row(a.sp, a.emb, a.emb_ld, a.row_map[nid]) : a.emb + nid * a.emb_ld;
            for (int v = lane; v < dv; v += 32) {
                float4 x = ld4(src, v);
                if (a.Neg[side]) st4(a.Neg[side] + j * d, v, x);
                if (a.Neg_hi[side]) store_split4(a.Neg_hi[side], a.Neg_lo[side], j * d + 4 * v, x);
            }
            continue;
        }
        const int64_t p = t;
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
        const int64_t si = a.edges[p * a.cols], ti = a.edges[p * a.cols + a.cols - 1];
        const float* src = a.row_map ? shard_row(a.sp, a.emb, a.emb_ld, a.row_map[si]) : a.emb + si * a.emb_ld;
        const float* dst = a.row_map ? shard_row(a.sp, a.emb, a.emb_ld, a.row_map[ti]) : a.emb + ti * a.emb_ld;
        const int64_t rid = (DEC != MB_DECODER_DOT) ? a.edges[p * a.cols + 1] : 0;
        const float* r = (DEC != MB_DECODER_DOT) ? a.rel + rid * d : nullptr;
        const float* ri = (DEC != MB_DECODER_DOT && a.sides == 2) ? a.inv_rel + rid * d : nullptr;
        float acc0 = 0.f, acc1 = 0.f;
        if (DEC == MB_DECODER_COMPLEX) {
            for (int v = lane; v < hv; v += 32) {
                float4 sr = ld4(src, v), sim = ld4(src, hv + v), dr = ld4(dst, v), dim = ld4(dst, hv + v);
                float4 rr = ld4(r, v), rim = ld4(r, hv + v);
                float4 ar = sub4(mul4(sr, rr), mul4(sim, rim));    // relation_operators.cpp:31
                float4 ai = addrn4(mul4(sr, rim), mul4(sim, rr));  // relation_operators.cpp:32
                acc0 = dot4(ar, dr, acc0);
                acc0 = dot4(ai, dim, acc0);
                if (a.A[0]) {
                    st4(a.A[0] + p * d, v, ar);
                    st4(a.A[0] + p * d, hv + v, ai);
                }
                if (a.A_hi[0]) {
                    store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * v, ar);
                    store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * (hv + v), ai);
                }
                if (a.sides == 2) {
                    float4 qr = ld4(ri, v), qi = ld4(ri, hv + v);
                    float4 br = sub4(mul4(dr, qr), mul4(dim, qi));
                    float4 bi = addrn4(mul4(dr, qi), mul4(dim, qr));
                    acc1 = dot4(br, sr, acc1);
                    acc1 = dot4(bi, sim, acc1);
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
            for (int v = lane; v < dv; v += 32) {
                float4 sv = ld4(src, v), dvv = ld4(dst, v);
                float4 av = (DEC == MB_DECODER_DISTMULT) ? mul4(sv, ld4(r, v)) : sv;  // relation_operators.cpp:11
                acc0 = dot4(av, dvv, acc0);
                if (a.A[0]) st4(a.A[0] + p * d, v, av);
                if (a.A_hi[0]) store_split4(a.A_hi[0], a.A_lo[0], p * d + 4 * v, av);
                if (a.sides == 2) {
                    float4 bv = mul4(dvv, ld4(ri, v));
                    acc1 = dot4(bv, sv, acc1);
                    if (a.A[1]) st4(a.A[1] + p * d, v, bv);
                    if (a.A_hi[1]) store_split4(a.A_hi[1], a.A_lo[1], p * d + 4 * v, bv);
                }
            }
        }
        acc0 = warp_sum(acc0);
        acc1 = warp_sum(acc1);
        if (lane == 0) {
            a.pos[0][p] = acc0;  // comparators.cpp:67-68
            if (a.sides == 2) a.pos[1][p] = acc1;
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
    ShardPtrs sp;
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

template <int DEC>
__global__ void __launch_bounds__(kThreads) edge_backward_kernel(EdgeBwdVArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, dv = d >> 2, hv = d >> 3;
    const bool inverse = a.sides == 2;
    for (int64_t i = warp0; i < a.B; i += nwarps) {
        const int64_t si = a.edges[i * a.cols], ti = a.edges[i * a.cols + a.cols - 1];
        const float* src = a.row_map ? shard_row(a.sp, a.emb, a.emb_ld, a.row_map[si]) : a.emb + si * a.emb_ld;
        const float* dst = a.row_map ? shard_row(a.sp, a.emb, a.emb_ld, a.row_map[ti]) : a.emb + ti * a.emb_ld;
        const int64_t rid = (DEC != MB_DECODER_DOT) ? a.edges[i * a.cols + 1] : 0;
        const float* r = (DEC != MB_DECODER_DOT) ? a.rel + rid * d : nullptr;
        const float* ri = (DEC != MB_DECODER_DOT && inverse) ? a.inv_rel + rid * d : nullptr;
        const float g0 = a.gpos[0][i];
        const float g1 = inverse ? a.gpos[1][i] : 0.f;
        const float* da0 = a.dA[0] + i * d;
        const float* da1 = inverse ? a.dA[1] + i * d : nullptr;
        float* dsrc = a.gcat + i * d;
        float* ddst = a.gcat + (a.B + i) * d;
        if (DEC == MB_DECODER_COMPLEX) {
            for (int v = lane; v < hv; v += 32) {
                float4 sr = ld4(src, v), sim = ld4(src, hv + v), dr = ld4(dst, v), dim = ld4(dst, hv + v);
                float4 rr = ld4(r, v), rim = ld4(r, hv + v);
                float4 ar = sub4(mul4(sr, rr), mul4(sim, rim)), ai = addrn4(mul4(sr, rim), mul4(sim, rr));
                float4 gar = fma4(g0, dr, ld4(da0, v)), gai = fma4(g0, dim, ld4(da0, hv + v));  // d/da: bmm backward + pos dot
                float4 dsr = add4(mul4(gar, rr), mul4(gai, rim));
                float4 dsi = sub4(mul4(gai, rr), mul4(gar, rim));
                if (a.drel[0]) {
                    st4(a.drel[0] + i * d, v, add4(mul4(gar, sr), mul4(gai, sim)));
                    st4(a.drel[0] + i * d, hv + v, sub4(mul4(gai, sr), mul4(gar, sim)));
                }
                float4 ddr = scale4(g0, ar), ddi = scale4(g0, ai);  // pos = <a, dst>
                if (inverse) {
                    float4 qr = ld4(ri, v), qi = ld4(ri, hv + v);
                    float4 br = sub4(mul4(dr, qr), mul4(dim, qi)), bi = addrn4(mul4(dr, qi), mul4(dim, qr));
                    float4 gbr = fma4(g1, sr, ld4(da1, v)), gbi = fma4(g1, sim, ld4(da1, hv + v));
                    ddr = add4(ddr, add4(mul4(gbr, qr), mul4(gbi, qi)));
                    ddi = add4(ddi, sub4(mul4(gbi, qr), mul4(gbr, qi)));
                    if (a.drel[1]) {
                        st4(a.drel[1] + i * d, v, add4(mul4(gbr, dr), mul4(gbi, dim)));
                        st4(a.drel[1] + i * d, hv + v, sub4(mul4(gbi, dr), mul4(gbr, dim)));
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
            for (int v = lane; v < dv; v += 32) {
                float4 sv = ld4(src, v), dvv = ld4(dst, v);
                float4 rv = (DEC == MB_DECODER_DISTMULT) ? ld4(r, v) : make_float4(1.f, 1.f, 1.f, 1.f);
                float4 av = (DEC == MB_DECODER_DISTMULT) ? mul4(sv, rv) : sv;
                float4 ga = fma4(g0, dvv, ld4(da0, v));
                float4 ds = (DEC == MB_DECODER_DISTMULT) ? mul4(ga, rv) : ga;
                if (a.drel[0]) st4(a.drel[0] + i * d, v, mul4(ga, sv));
                float4 dd = scale4(g0, av);
                if (inverse) {
                    float4 qv = ld4(ri, v);
                    float4 bv = mul4(dvv, qv);
                    float4 gb = fma4(g1, sv, ld4(da1, v));
                    dd = add4(dd, mul4(gb, qv));
                    if (a.drel[1]) st4(a.drel[1] + i * d, v, mul4(gb, dvv));
                    ds = fma4(g1, bv, ds);
                }
                st4(dsrc, v, ds);
                st4(ddst, v, dd);
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
    for (int64_t u = warp0; u < a.n_seg; u += nwarps) {
        const uint32_t beg = a.offsets[u], end = a.offsets[u + 1];
        float4 e[CH], s[CH];
        float *erow = nullptr, *srow = nullptr;
        if (MODE == 2 && beg == end) continue;  // padding segment (graph replay runs with n_seg = capacity): no row, nothing to update
        if (MODE == 2) {  // issue the table reads first: they do not depend on the slot list
            const int64_t r = a.ids[u];
            if (a.sp.world <= 1) {
                erow = a.table + r * a.ld;
                srow = a.state_table + r * a.ld;
            } else {  // the owner's HBM (peer-mapped when remote): Adagrad read-modify-write straight over NVLink
                const int64_t o = r / a.sp.rows_per_rank, lr_ = r - o * a.sp.rows_per_rank;
                erow = a.sp.table[o] + lr_ * a.ld;
                srow = a.sp.state[o] + lr_ * a.ld;
            }
#pragma unroll
            for (int c = 0; c < CH; c++) {
                int v = lane + 32 * c;
                if (v < dv) {
                    e[c] = reinterpret_cast<const float4*>(erow)[v];
                    s[c] = reinterpret_cast<const float4*>(srow)[v];
                }
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                int v = lane + 32 * c;
                if (v < dv) s[c] = ld4(a.state + u * a.state_ld, v);
            }
        }
        float4 acc[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t q = beg; q < end; q++) {
            const float* row = a.rows + (int64_t)a.slots[q] * d;
#pragma unroll
            for (int c = 0; c < CH; c++) {
                int v = lane + 32 * c;
                if (v < dv) acc[c] = addrn4(acc[c], ld_f4(reinterpret_cast<const float4*>(row) + v));
            }
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            int v = lane + 32 * c;
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
                st_stream(reinterpret_cast<float4*>(erow) + v, addrn4(e[c], de));
                st_stream(reinterpret_cast<float4*>(srow) + v, sn);
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
