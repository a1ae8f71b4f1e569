// decoder_kernels.cu -- the non-GEMM kernels of node_corrupt_forward / SoftmaxCrossEntropy / their backward:
//   edge_prep      : per positive edge, gather src/dst/relation rows from the batch-local embedding matrix, apply the
//                    relation operator (Hadamard / ComplexHadamard), write the adjusted rows A (both corruption sides)
//                    and the positive scores (warp-level dot).          decoder_methods.cpp:74-97, relation_operators.cpp
//   gather_split   : gather negative rows (fp32 and/or bf16 hi+lo operand arrays for the tcgen05 GEMM)   decoder_methods.cpp:79,93
//   loss_grad      : SoftmaxCrossEntropy over [pos, logsumexp(neg)] per row, forward value + gradient     loss.cpp:50-67
//   edge_backward  : chain rule through DotCompare(pos) and the relation operator                           (autograd of the above)
//   segment_reduce : duplicate-index accumulation of row gradients by sorted slot lists (no atomics), optionally fused
//                    with Batch::accumulateGradients + the two Storage::indexAdd                            batch.cpp:62-79
// One warp per row everywhere: d=400 -> 100 float4 per row, 3.1 per lane; shuffles for the dots.
#include <cuda_bf16.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "decoder_vec.cuh"
#include "kernels.h"
#include "row_bulk.cuh"

namespace mb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = kThreads / 32;

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

struct EdgePrepArgs {
    const float* emb;      // [U, emb_ld]
    int64_t emb_ld;
    const int64_t* edges;  // [B, cols]
    int cols;
    const float* rel;      // [R, d] or null
    const float* inv_rel;  // [R, d] or null (no inverse side)
    int64_t B, Bp;
    int d;
    int decoder;
    float* A0;  // [Bp, d] adjusted src   (side 0: corrupt dst)
    float* A1;  // [Bp, d] adjusted dst   (side 1: corrupt src) or null
    float* pos0;  // [Bp]
    float* pos1;  // [Bp] or null
    __nv_bfloat16 *A0_hi, *A0_lo, *A1_hi, *A1_lo;  // optional bf16 split copies [Bp, d]
};

__global__ void __launch_bounds__(kThreads) edge_prep_kernel(EdgePrepArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, h = d / 2;
    const bool inverse = a.A1 != nullptr;
    for (int64_t p = warp0; p < a.Bp; p += nwarps) {
        float* a0 = a.A0 + p * d;
        float* a1 = inverse ? a.A1 + p * d : nullptr;
        if (p >= a.B) {  // pad_and_reshape zero rows (comparators.cpp:11-15), zero-padded pos (decoder_methods.cpp:103-111)
            for (int j = lane; j < d; j += 32) {
                a0[j] = 0.f;
                if (a1) a1[j] = 0.f;
                if (a.A0_hi) {
                    a.A0_hi[p * d + j] = __float2bfloat16_rn(0.f);
                    a.A0_lo[p * d + j] = __float2bfloat16_rn(0.f);
                }
                if (a.A1_hi) {
                    a.A1_hi[p * d + j] = __float2bfloat16_rn(0.f);
                    a.A1_lo[p * d + j] = __float2bfloat16_rn(0.f);
                }
            }
            if (lane == 0) {
                a.pos0[p] = 0.f;
                if (a.pos1) a.pos1[p] = 0.f;
            }
            continue;
        }
        const int64_t s = a.edges[p * a.cols], t = a.edges[p * a.cols + a.cols - 1];
        const float* src = a.emb + s * a.emb_ld;
        const float* dst = a.emb + t * a.emb_ld;
        const bool has_rel = (a.cols == 3) && (a.decoder != MB_DECODER_DOT) && a.rel != nullptr;
        const int64_t rid = has_rel ? a.edges[p * a.cols + 1] : 0;
        const float* r = has_rel ? a.rel + rid * d : nullptr;
        const float* ri = (has_rel && a.inv_rel) ? a.inv_rel + rid * d : nullptr;
        float acc0 = 0.f, acc1 = 0.f;
        if (a.decoder == MB_DECODER_COMPLEX && has_rel) {
            // ComplexHadamardOperator (relation_operators.cpp:14-35): halves [0,h) real, [h,d) imaginary
            for (int j = lane; j < h; j += 32) {
                float sr = src[j], si = src[j + h], dr = dst[j], di = dst[j + h];
                float rr = r[j], rim = r[j + h];
                float ar = __fsub_rn(__fmul_rn(sr, rr), __fmul_rn(si, rim));
                float ai = __fadd_rn(__fmul_rn(sr, rim), __fmul_rn(si, rr));
                a0[j] = ar;
                a0[j + h] = ai;
                acc0 = fmaf(ar, dr, acc0);
                acc0 = fmaf(ai, di, acc0);
                if (a.A0_hi) {
                    split_bf16(ar, a.A0_hi[p * d + j], a.A0_lo[p * d + j]);
                    split_bf16(ai, a.A0_hi[p * d + j + h], a.A0_lo[p * d + j + h]);
                }
                if (inverse) {
                    float qr = ri[j], qi = ri[j + h];
                    float br = __fsub_rn(__fmul_rn(dr, qr), __fmul_rn(di, qi));
                    float bi = __fadd_rn(__fmul_rn(dr, qi), __fmul_rn(di, qr));
                    a1[j] = br;
                    a1[j + h] = bi;
                    acc1 = fmaf(br, sr, acc1);
                    acc1 = fmaf(bi, si, acc1);
                    if (a.A1_hi) {
                        split_bf16(br, a.A1_hi[p * d + j], a.A1_lo[p * d + j]);
                        split_bf16(bi, a.A1_hi[p * d + j + h], a.A1_lo[p * d + j + h]);
                    }
                }
            }
        } else {
            for (int j = lane; j < d; j += 32) {
                float sv = src[j], dv = dst[j];
                float av = has_rel ? __fmul_rn(sv, r[j]) : sv;  // HadamardOperator (relation_operators.cpp:7-12) / identity
                a0[j] = av;
                acc0 = fmaf(av, dv, acc0);
                if (a.A0_hi) split_bf16(av, a.A0_hi[p * d + j], a.A0_lo[p * d + j]);
                if (inverse) {
                    float bv = ri ? __fmul_rn(dv, ri[j]) : dv;
                    a1[j] = bv;
                    acc1 = fmaf(bv, sv, acc1);
                    if (a.A1_hi) split_bf16(bv, a.A1_hi[p * d + j], a.A1_lo[p * d + j]);
                }
            }
        }
        acc0 = warp_sum(acc0);
        acc1 = warp_sum(acc1);
        if (lane == 0) {
            a.pos0[p] = acc0;  // DotCompare same-shape branch (comparators.cpp:67-68)
            if (a.pos1) a.pos1[p] = acc1;
        }
    }
}

// out[r,:] = emb[idx[r],:]  (fp32 copy and/or bf16 hi/lo split), one warp per row.
__global__ void __launch_bounds__(kThreads) gather_split_kernel(const float* __restrict__ emb, int64_t emb_ld, const int64_t* __restrict__ idx, int64_t n,
                                                                int d, float* __restrict__ out, __nv_bfloat16* __restrict__ hi,
                                                                __nv_bfloat16* __restrict__ lo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const float* src = emb + idx[r] * emb_ld;
        for (int j = lane; j < d; j += 32) {
            float v = src[j];
            if (out) out[r * d + j] = v;
            if (hi) split_bf16(v, hi[r * d + j], lo[r * d + j]);
        }
    }
}

// SoftmaxCrossEntropy (loss.cpp:50-67) row-wise:  L_i = log(e^{pos_i} + sum_j e^{neg_ij}) - pos_i.
// Writes G = dL/dneg in place of the scores (and/or as bf16 hi/lo), gpos = dL/dpos, and the row loss.
struct LossArgs {
    float* S;           // [rows, N] scores in, gradient out (in place)
    const float* pos;   // [rows]
    float* gpos;        // [rows]
    float* row_loss;    // [rows]
    __nv_bfloat16 *G_hi, *G_lo;  // optional [rows, ldg]
    int64_t rows;
    int N;
    float w;            // 1 (SUM) or 1/rows_per_side (MEAN)
    int64_t ldg;        // leading dimension of G_hi / G_lo (>= N; padded so that rows start on 128-byte lines)
};

__global__ void __launch_bounds__(kThreads) loss_grad_kernel(LossArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t i = warp0; i < a.rows; i += nwarps) {
        float* s = a.S + i * a.N;
        const float p = a.pos[i];
        float m = p;
        for (int j = lane; j < a.N; j += 32) m = fmaxf(m, s[j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < a.N; j += 32) sum += expf(s[j] - m);
        sum = warp_sum(sum);
        sum += expf(p - m);
        const float z = m + logf(sum);
        for (int j = lane; j < a.N; j += 32) {
            float g = expf(s[j] - z) * a.w;
            s[j] = g;
            if (a.G_hi) split_bf16(g, a.G_hi[i * a.ldg + j], a.G_lo[i * a.ldg + j]);
        }
        if (lane == 0) {
            a.gpos[i] = (expf(p - z) - 1.0f) * a.w;
            a.row_loss[i] = (z - p) * a.w;
        }
    }
}

// loss = sum(row_loss[0..n)) in a fixed order (single block)
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ row_loss, int64_t n, float* __restrict__ loss) {
    __shared__ float ws[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += row_loss[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = ws[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) loss[0] = v;
    }
}

struct EdgeBwdArgs {
    const float* emb;
    int64_t emb_ld;
    const int64_t* edges;
    int cols;
    const float* rel;
    const float* inv_rel;
    int64_t B;
    int d;
    int decoder;
    const float *A0, *A1;        // adjusted rows [Bp,d]
    const float *dA0, *dA1;      // GEMM outputs G.Neg [Bp,d]
    const float *gpos0, *gpos1;  // [Bp]
    float* gcat;                 // [2B + 2CN, d]: rows [0,B) d src, [B,2B) d dst, then d dst_negs, d src_negs
    float *drel0, *drel1;        // per-edge relation gradients [B,d] (null when no relations)
};

__global__ void __launch_bounds__(kThreads) edge_backward_kernel(EdgeBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d, h = d / 2;
    const bool inverse = a.A1 != nullptr;
    for (int64_t i = warp0; i < a.B; i += nwarps) {
        const int64_t s = a.edges[i * a.cols], t = a.edges[i * a.cols + a.cols - 1];
        const float* src = a.emb + s * a.emb_ld;
        const float* dst = a.emb + t * a.emb_ld;
        const bool has_rel = (a.cols == 3) && (a.decoder != MB_DECODER_DOT) && a.rel != nullptr;
        const int64_t rid = has_rel ? a.edges[i * a.cols + 1] : 0;
        const float* r = has_rel ? a.rel + rid * d : nullptr;
        const float* ri = (has_rel && a.inv_rel) ? a.inv_rel + rid * d : nullptr;
        const float g0 = a.gpos0[i];
        const float g1 = inverse ? a.gpos1[i] : 0.f;
        const float* a0 = a.A0 + i * d;
        const float* a1 = inverse ? a.A1 + i * d : nullptr;
        const float* da0 = a.dA0 + i * d;
        const float* da1 = inverse ? a.dA1 + i * d : nullptr;
        float* dsrc = a.gcat + i * d;
        float* ddst = a.gcat + (a.B + i) * d;
        if (a.decoder == MB_DECODER_COMPLEX && has_rel) {
            for (int j = lane; j < h; j += 32) {
                float sr = src[j], si = src[j + h], dr = dst[j], di = dst[j + h];
                float rr = r[j], rim = r[j + h];
                // d loss / d a  (a = adjusted src): bmm backward + pos-dot backward
                float gar = fmaf(g0, dr, da0[j]), gai = fmaf(g0, di, da0[j + h]);
                // through a = src (x) r : d src, d r
                float dsr = gar * rr + gai * rim, dsi = -gar * rim + gai * rr;
                if (a.drel0) {
                    a.drel0[i * d + j] = gar * sr + gai * si;
                    a.drel0[i * d + j + h] = -gar * si + gai * sr;
                }
                float ddr = g0 * a0[j], ddi = g0 * a0[j + h];  // pos = <a, dst>
                if (inverse) {
                    float qr = ri[j], qi = ri[j + h];
                    float gbr = fmaf(g1, sr, da1[j]), gbi = fmaf(g1, si, da1[j + h]);
                    ddr += gbr * qr + gbi * qi;
                    ddi += -gbr * qi + gbi * qr;
                    if (a.drel1) {
                        a.drel1[i * d + j] = gbr * dr + gbi * di;
                        a.drel1[i * d + j + h] = -gbr * di + gbi * dr;
                    }
                    dsr = fmaf(g1, a1[j], dsr);  // inv_pos = <b, src>
                    dsi = fmaf(g1, a1[j + h], dsi);
                }
                dsrc[j] = dsr;
                dsrc[j + h] = dsi;
                ddst[j] = ddr;
                ddst[j + h] = ddi;
            }
        } else {
            for (int j = lane; j < d; j += 32) {
                float sv = src[j], dv = dst[j];
                float ga = fmaf(g0, dv, da0[j]);
                float ds = has_rel ? ga * r[j] : ga;
                if (a.drel0) a.drel0[i * d + j] = ga * sv;
                float dd = g0 * a0[j];
                if (inverse) {
                    float gb = fmaf(g1, sv, da1[j]);
                    dd += ri ? gb * ri[j] : gb;
                    if (a.drel1) a.drel1[i * d + j] = gb * dv;
                    ds = fmaf(g1, a1[j], ds);
                }
                dsrc[j] = ds;
                ddst[j] = dd;
            }
        }
    }
}

// keys[slot] = batch-local node id of every gradient slot.  Slot order: src | dst | dst_negs | src_negs -- the rows of
// `gcat`; (side, chunk) blocks of the negative-gradient GEMM output are then contiguous.  (The reference's all_ids order is
// src | dst | src_negs | dst_negs, dataloader.cpp:399-409; the order only fixes the fp32 summation order of duplicates.)
__global__ void slot_keys_kernel(const int64_t* __restrict__ edges, int cols, int64_t B, const int64_t* __restrict__ dst_negs,
                                 const int64_t* __restrict__ src_negs, int64_t CN, uint32_t* __restrict__ keys) {
    int64_t total = 2 * B + 2 * CN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k;
        if (i < B)
            k = edges[i * cols];
        else if (i < 2 * B)
            k = edges[(i - B) * cols + cols - 1];
        else if (i < 2 * B + CN)
            k = dst_negs[i - 2 * B];
        else
            k = src_negs ? src_negs[i - 2 * B - CN] : -1;
        keys[i] = (uint32_t)k;  // -1 -> 0xffffffff sorts last and falls outside [0,U): never reduced
    }
}

__global__ void rel_keys_kernel(const int64_t* __restrict__ edges, int cols, int64_t B, uint32_t* __restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) keys[i] = (uint32_t)edges[i * cols + 1];
}

// out_row(u) = sum over slots of segment u of rows[slot], slot order ascending (stable sort) => deterministic.
// MODE 0: write the sum to out[u]                          (node gradient / relation gradient)
// MODE 1: + Batch::accumulateGradients -> delta_e, delta_s given the gathered state rows   (batch.cpp:62-79)
// MODE 2: + fused Adagrad read-modify-write of table[ids[u]] and state_table[ids[u]]      (dataloader.cpp:550-557)
struct SegReduceArgs {
    const float* rows;         // [n_slots, d]
    const uint32_t* slots;     // sorted slot ids
    const uint32_t* offsets;   // [n_seg + 1]
    int64_t n_seg;
    int d;
    float* out;                // MODE 0/1: [n_seg, out_ld] gradient (may be null in MODE 1)
    int64_t out_ld;
    const float* state;        // MODE 1: [n_seg, state_ld]
    int64_t state_ld;
    float *delta_e, *delta_s;  // MODE 1: [n_seg, d]
    float *table, *state_table;  // MODE 2
    int64_t ld;
    const int64_t* ids;        // MODE 2: [n_seg] table rows
    float neg_lr;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) segment_reduce_kernel(SegReduceArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int d = a.d;
    constexpr int CH = 4;  // columns per lane per sweep (independent accumulators => 4 loads in flight per slot)
    for (int64_t u = warp0; u < a.n_seg; u += nwarps) {
        const uint32_t beg = a.offsets[u], end = a.offsets[u + 1];
        float* erow = nullptr;
        float* srow = nullptr;
        if (MODE == 2 && beg == end) continue;  // padding segment: no table row behind it
        if (MODE == 2) {
            int64_t r = a.ids[u];
            erow = a.table + r * a.ld;
            srow = a.state_table + r * a.ld;
        }
        for (int j0 = 0; j0 < d; j0 += 32 * CH) {
            float acc[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) acc[c] = 0.f;
            for (uint32_t q = beg; q < end; q++) {
                const float* row = a.rows + (int64_t)a.slots[q] * d;
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    int j = j0 + c * 32 + lane;
                    if (j < d) acc[c] = __fadd_rn(acc[c], row[j]);
                }
            }
#pragma unroll
            for (int c = 0; c < CH; c++) {
                int j = j0 + c * 32 + lane;
                if (j >= d) continue;
                float g = acc[c];
                if (MODE == 0) {
                    a.out[u * a.out_ld + j] = g;
                } else if (MODE == 1) {
                    if (a.out) a.out[u * a.out_ld + j] = g;
                    float de, ds, sn;
                    adagrad_rule(g, a.state[u * a.state_ld + j], a.neg_lr, de, ds, sn);
                    a.delta_e[u * d + j] = de;
                    a.delta_s[u * d + j] = ds;
                } else {
                    float de, ds, sn;
                    adagrad_rule(g, srow[j], a.neg_lr, de, ds, sn);
                    erow[j] = __fadd_rn(erow[j], de);
                    srow[j] = sn;
                }
            }
        }
    }
}

__global__ void split_kernel(const float* __restrict__ x, int64_t n, int64_t cols, int64_t ld_out, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = (i / cols) * ld_out + (i % cols);
        split_bf16(x[i], hi[o], lo[o]);
    }
}

inline int warp_grid(int64_t rows) {
    int64_t blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

struct EdgePrepLaunch {
    EdgePrepArgs args;
};

mb_status launch_edge_prep(const float* emb, int64_t emb_ld, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int64_t Bp,
                           int d, int decoder, float* A0, float* A1, float* pos0, float* pos1, void* A0_hi, void* A0_lo, void* A1_hi, void* A1_lo,
                           cudaStream_t st) {
    EdgePrepArgs a{emb, emb_ld, edges, cols, rel, inv_rel, B, Bp, d, decoder, A0, A1, pos0, pos1,
                   (__nv_bfloat16*)A0_hi, (__nv_bfloat16*)A0_lo, (__nv_bfloat16*)A1_hi, (__nv_bfloat16*)A1_lo};
    edge_prep_kernel<<<warp_grid(Bp), kThreads, 0, st>>>(a);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_gather_split(const float* emb, int64_t emb_ld, const int64_t* idx, int64_t n, int d, float* out, void* hi, void* lo, cudaStream_t st) {
    if (n == 0) return MB_OK;
    gather_split_kernel<<<warp_grid(n), kThreads, 0, st>>>(emb, emb_ld, idx, n, d, out, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_split(const float* x, int64_t n, void* hi, void* lo, cudaStream_t st, int64_t cols, int64_t ld_out) {
    if (n == 0) return MB_OK;
    if (cols <= 0) cols = ld_out = n;  // flat
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    split_kernel<<<(int)blocks, 256, 0, st>>>(x, n, cols, ld_out, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_loss_grad(float* S, const float* pos, float* gpos, float* row_loss, void* G_hi, void* G_lo, int64_t rows, int N, float w,
                           cudaStream_t st, int64_t ldg) {
    LossArgs a{S, pos, gpos, row_loss, (__nv_bfloat16*)G_hi, (__nv_bfloat16*)G_lo, rows, N, w, ldg > 0 ? ldg : N};
    loss_grad_kernel<<<warp_grid(rows), kThreads, 0, st>>>(a);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

// Fused SoftmaxCrossEntropy, second half (the first half is the epilogue of the score contraction, gemm_tc_group.cu): merge the
// per-slot (max, sum exp) statistics of a score row with its positive score into z = log(e^pos + sum_j e^neg_j)  (loss.cpp:57-66),
// then  row_loss = (z - pos) w ,  d loss / d pos = (e^(pos - z) - 1) w ,  zw = (z - log w) log2 e  so that  d loss / d neg_j = exp2(neg_j log2 e - zw).
// 16 lanes per score row (one statistics slot each, coalesced 128-byte reads), 16 rows per block; the block also leaves the sum of
// its rows' losses in block_loss[blockIdx.x] (fixed order: deterministic), so the final loss reduction only has rows / 16 terms.
__global__ void __launch_bounds__(256) loss_merge_kernel(const float2* __restrict__ stats, int slots, const float* __restrict__ pos,
                                                          float* __restrict__ gpos, float* __restrict__ row_loss, float* __restrict__ zw, int64_t rows,
                                                          float w, float log_w, float* __restrict__ block_loss) {
    __shared__ float part[16];
    const int sub = threadIdx.x & 15, rib = threadIdx.x >> 4;
    const int64_t i = (int64_t)blockIdx.x * 16 + rib;
    const bool on = i < rows;
    const float p = on ? pos[i] : 0.f;
    float m = p, lsum = 0.f;
    float2 x[2];  // up to 32 slots (N <= 2048) in registers, more in a second pass
    const float2* st = stats + (on ? i : 0) * slots;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int t = sub + 16 * k;
        x[k] = (on && t < slots) ? st[t] : make_float2(0.f, 0.f);
        if (x[k].y > 0.f) m = fmaxf(m, x[k].x);  // a slot no tile wrote has sum == 0
    }
    for (int t = sub + 32; on && t < slots; t += 16) {
        const float2 y = st[t];
        if (y.y > 0.f) m = fmaxf(m, y.x);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
#pragma unroll
    for (int k = 0; k < 2; k++)
        if (x[k].y > 0.f) lsum += x[k].y * expf(x[k].x - m);
    for (int t = sub + 32; on && t < slots; t += 16) {
        const float2 y = st[t];
        if (y.y > 0.f) lsum += y.y * expf(y.x - m);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);  // (xor butterfly: every lane holds the same sum)
    float rl = 0.f;
    if (on) {
        const float sum = lsum + expf(p - m);
        const float z = m + logf(sum);
        rl = (z - p) * w;
        if (sub == 0) {
            gpos[i] = (expf(p - z) - 1.0f) * w;
            row_loss[i] = rl;
            zw[i] = (z - log_w) * 1.4426950408889634f;  // pre-scaled by log2(e): the converter warps evaluate exp2(S * log2 e - zw)
        }
    }
    if (sub == 0) part[rib] = rl;
    __syncthreads();
    if (threadIdx.x == 0) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) acc += part[k];
        block_loss[blockIdx.x] = acc;
    }
}

// rows / 16 partial sums, written by loss_merge_kernel into `block_loss`; returns how many
int64_t loss_merge_blocks(int64_t rows) { return (rows + 15) / 16; }

mb_status launch_loss_merge(const float2* stats, int slots, const float* pos, float* gpos, float* row_loss, float* zw, int64_t rows, float w,
                            float* block_loss, cudaStream_t st) {
    if (rows == 0) return MB_OK;
    loss_merge_kernel<<<(unsigned)loss_merge_blocks(rows), 256, 0, st>>>(stats, slots, pos, gpos, row_loss, zw, rows, w, logf(w), block_loss);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_loss_reduce(const float* row_loss, int64_t n, float* loss, cudaStream_t st) {
    loss_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, n, loss);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_edge_backward(const float* emb, int64_t emb_ld, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int d,
                               int decoder, const float* A0, const float* A1, const float* dA0, const float* dA1, const float* gpos0,
                               const float* gpos1, float* gcat, float* drel0, float* drel1, cudaStream_t st) {
    if (B == 0) return MB_OK;
    EdgeBwdArgs a{emb, emb_ld, edges, cols, rel, inv_rel, B, d, decoder, A0, A1, dA0, dA1, gpos0, gpos1, gcat, drel0, drel1};
    edge_backward_kernel<<<warp_grid(B), kThreads, 0, st>>>(a);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_slot_keys(const int64_t* edges, int cols, int64_t B, const int64_t* dst_negs, const int64_t* src_negs, int64_t CN, uint32_t* keys,
                           cudaStream_t st) {
    int64_t total = 2 * B + 2 * CN;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    slot_keys_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(edges, cols, B, dst_negs, src_negs, CN, keys);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_rel_keys(const int64_t* edges, int cols, int64_t B, uint32_t* keys, cudaStream_t st) {
    int blocks = (int)((B + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    rel_keys_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(edges, cols, B, keys);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_segment_reduce(int mode, const float* rows, const uint32_t* slots, const uint32_t* offsets, int64_t n_seg, int d, float* out, int64_t out_ld,
                                const float* state, int64_t state_ld, float* delta_e, float* delta_s, float* table, float* state_table, int64_t ld,
                                const int64_t* ids, float lr, cudaStream_t st) {
    if (n_seg == 0) return MB_OK;
    SegReduceArgs a{rows, slots, offsets, n_seg, d, out, out_ld, state, state_ld, delta_e, delta_s, table, state_table, ld, ids, -lr};
    int grid = warp_grid(n_seg);
    if (mode == 0)
        segment_reduce_kernel<0><<<grid, kThreads, 0, st>>>(a);
    else if (mode == 1)
        segment_reduce_kernel<1><<<grid, kThreads, 0, st>>>(a);
    else
        segment_reduce_kernel<2><<<grid, kThreads, 0, st>>>(a);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

// ---- unified launchers: 128-bit vector kernels when the shapes allow, scalar kernels otherwise ------------------------------
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool decoder_vec_ok(const float* emb, int64_t emb_ld, int d, bool has_rel, const float* rel, const float* inv_rel, int sides) {
    return (d % 8 == 0) && d <= 512 && (emb_ld % 4 == 0) && al16(emb) && (!has_rel || (al16(rel) && (sides == 1 || al16(inv_rel))));
}

static vec::ShardPtrs make_sp(const mb_shards* sh) {
    vec::ShardPtrs sp;
    std::memset(&sp, 0, sizeof(sp));
    if (sh != nullptr && sh->world > 1) {
        sp.world = sh->world;
        sp.rows_per_rank = sh->rows_per_rank;
        sp.rank = sh->rank;
        for (int i = 0; i < sh->world && i < 8; i++) {
            sp.table[i] = sh->tables[i];
            sp.state[i] = sh->states[i];
        }
    }
    return sp;
}

// opt a bulk-staged kernel in to its dynamic shared memory, once per (kernel, device)
static mb_status bulk_attr(const void* fn, size_t smem, int slot) {
    static std::atomic<size_t> done[8][64];
    int dev = 0;
    MB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || done[slot][dev].load(std::memory_order_acquire) < smem) {
        MB_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) done[slot][dev].store(smem, std::memory_order_release);
    }
    return MB_OK;
}

mb_status launch_prep(const mb_shards* sh, cudaEvent_t rows_fetched, const float* const* row_ptrs, const float* emb, int64_t emb_ld, const int64_t* row_map, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int64_t Bp,
                      int64_t CN, int d, int decoder, int sides, const int64_t* dst_negs, const int64_t* src_negs, float* A /*[sides][Bp][d] or null*/,
                      float* pos /*[sides][Bp]*/, void* A_hi, void* A_lo /*[sides][Bp][d] or null*/, float* Neg /*[sides][CN][d] or null*/,
                      void* Neg_hi, void* Neg_lo, cudaStream_t st) {
    const bool has_rel = (cols == 3) && (decoder != MB_DECODER_DOT) && rel != nullptr;
    const int dec = has_rel ? decoder : MB_DECODER_DOT;
    const bool vec_ok = decoder_vec_ok(emb, emb_ld, d, has_rel, rel, inv_rel, sides);
    if (vec_ok) {
        vec::PrepArgs a;
        a.emb = emb;
        a.emb_ld = emb_ld;
        a.row_map = row_map;
        a.row_ptrs = row_ptrs;
        a.sp = make_sp(sh);
        a.edges = edges;
        a.cols = cols;
        a.rel = rel;
        a.inv_rel = inv_rel;
        a.B = B;
        a.Bp = Bp;
        a.CN = CN;
        a.d = d;
        a.decoder = dec;
        a.sides = sides;
        a.negs[0] = dst_negs;
        a.negs[1] = src_negs;
        for (int s = 0; s < 2; s++) {
            const bool on = s < sides;
            a.A[s] = (on && A) ? A + s * Bp * d : nullptr;
            a.pos[s] = on ? pos + s * Bp : nullptr;
            a.A_hi[s] = (on && A_hi) ? (__nv_bfloat16*)A_hi + s * Bp * d : nullptr;
            a.A_lo[s] = (on && A_lo) ? (__nv_bfloat16*)A_lo + s * Bp * d : nullptr;
            a.Neg[s] = (on && Neg) ? Neg + s * CN * d : nullptr;
            a.Neg_hi[s] = (on && Neg_hi) ? (__nv_bfloat16*)Neg_hi + s * CN * d : nullptr;
            a.Neg_lo[s] = (on && Neg_lo) ? (__nv_bfloat16*)Neg_lo + s * CN * d : nullptr;
        }
        // negative rows first (two in flight per warp): with a sharded table they run while the remote rows are still being fetched
        // MB_ROW_BULK: bit 0 = negative rows through the bulk-copy staged kernel (default; 0 = the register-staged kernel),
        // bit 2 = eight consumer warps per block instead of four
        static const int bulk_mode = [] { const char* e = getenv("MB_ROW_BULK"); return e ? atoi(e) : 1; }();
        if (CN > 0 && (bulk_mode & 1)) {
            // bulk-copy staged version (row_bulk.cuh): one cp.async.bulk per row into a shared-memory ring, two blocks per SM
            const size_t smem = bulk::smem_bytes(bulk::kNegRows, d);
            const int64_t chunks = ((int64_t)sides * CN + bulk::kNegRows - 1) / bulk::kNegRows;
            const int grid = (int)std::min<int64_t>(chunks, (int64_t)sm_count() * 2);
            if (bulk_mode & 4) {
                MB_TRY(bulk_attr(reinterpret_cast<const void*>(bulk::neg_rows_bulk_kernel<8>), smem, 4));
                bulk::neg_rows_bulk_kernel<8><<<grid, 32 * 9, smem, st>>>(a);
            } else {
                MB_TRY(bulk_attr(reinterpret_cast<const void*>(bulk::neg_rows_bulk_kernel<4>), smem, 0));
                bulk::neg_rows_bulk_kernel<4><<<grid, 32 * 5, smem, st>>>(a);
            }
            MB_LAUNCH_CHECK();
        } else if (CN > 0) {
            const int grid = warp_grid((sides * CN + 1) / 2);
            if (d <= 128)
                vec::neg_rows_kernel<1><<<grid, vec::kThreads, 0, st>>>(a);
            else
                vec::neg_rows_kernel<4><<<grid, vec::kThreads, 0, st>>>(a);
            MB_LAUNCH_CHECK();
        }
        if (rows_fetched != nullptr) MB_CUDA_TRY(cudaStreamWaitEvent(st, rows_fetched, 0));
        if (Bp > 0) {
            const int grid = warp_grid(Bp);
            const bool small = d <= 128;  // chunks per lane: full row <= 1 (d <= 128) / 4 (d <= 512); complex half <= 1 (d <= 256) / 2
            if (dec == MB_DECODER_COMPLEX) {
                if (d <= 256)
                    vec::edge_rows_kernel<MB_DECODER_COMPLEX, 1><<<grid, vec::kThreads, 0, st>>>(a);
                else
                    vec::edge_rows_kernel<MB_DECODER_COMPLEX, 2><<<grid, vec::kThreads, 0, st>>>(a);
            } else if (dec == MB_DECODER_DISTMULT) {
                if (small)
                    vec::edge_rows_kernel<MB_DECODER_DISTMULT, 1><<<grid, vec::kThreads, 0, st>>>(a);
                else
                    vec::edge_rows_kernel<MB_DECODER_DISTMULT, 4><<<grid, vec::kThreads, 0, st>>>(a);
            } else {
                if (small)
                    vec::edge_rows_kernel<MB_DECODER_DOT, 1><<<grid, vec::kThreads, 0, st>>>(a);
                else
                    vec::edge_rows_kernel<MB_DECODER_DOT, 4><<<grid, vec::kThreads, 0, st>>>(a);
            }
            MB_LAUNCH_CHECK();
        }
        return MB_OK;
    }
    // scalar fallback needs the fp32 adjusted rows and a batch-local embedding matrix
    if (row_map != nullptr || row_ptrs != nullptr) {
        set_error("launch_prep: row_map requires the vector path");
        return MB_ERR_INVALID;
    }
    MB_TRY(launch_edge_prep(emb, emb_ld, edges, cols, rel, sides == 2 ? inv_rel : nullptr, B, Bp, d, decoder, A, sides == 2 ? A + Bp * d : nullptr, pos,
                            sides == 2 ? pos + Bp : nullptr, A_hi, A_lo, (A_hi && sides == 2) ? (void*)((__nv_bfloat16*)A_hi + Bp * d) : nullptr,
                            (A_lo && sides == 2) ? (void*)((__nv_bfloat16*)A_lo + Bp * d) : nullptr, st));
    for (int s = 0; s < sides; s++) {
        MB_TRY(launch_gather_split(emb, emb_ld, s == 0 ? dst_negs : src_negs, CN, d, Neg ? Neg + s * CN * d : nullptr,
                                   Neg_hi ? (void*)((__nv_bfloat16*)Neg_hi + s * CN * d) : nullptr,
                                   Neg_lo ? (void*)((__nv_bfloat16*)Neg_lo + s * CN * d) : nullptr, st));
    }
    return MB_OK;
}

// S -> G (fp32, may be in place or null) and/or bf16 hi/lo
mb_status launch_loss(const float* S, float* G, const float* pos, float* gpos, float* row_loss, void* G_hi, void* G_lo, int64_t rows, int N, float w,
                      cudaStream_t st, int64_t ldg) {
    if (rows == 0) return MB_OK;
    if (ldg <= 0) ldg = N;
    if ((N % 4 == 0) && (ldg % 4 == 0) && N <= 1024 && al16(S) && (!G || al16(G))) {
        vec::LossVArgs a{S, G, pos, gpos, row_loss, (__nv_bfloat16*)G_hi, (__nv_bfloat16*)G_lo, rows, N, w, ldg};
        int grid = warp_grid(rows);
        if (N <= 256)
            vec::loss_kernel<2><<<grid, vec::kThreads, 0, st>>>(a);
        else
            vec::loss_kernel<8><<<grid, vec::kThreads, 0, st>>>(a);
        MB_LAUNCH_CHECK();
        return MB_OK;
    }
    // scalar kernel works in place on an fp32 buffer
    if (G == nullptr) {
        set_error("launch_loss: scalar fallback needs an fp32 gradient buffer");
        return MB_ERR_INVALID;
    }
    if (G != S) MB_CUDA_TRY(cudaMemcpyAsync(G, S, sizeof(float) * rows * N, cudaMemcpyDeviceToDevice, st));
    return launch_loss_grad(G, pos, gpos, row_loss, G_hi, G_lo, rows, N, w, st, ldg);
}

mb_status launch_edge_bwd(const float* const* row_ptrs, const float* emb, int64_t emb_ld, const int64_t* row_map, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int64_t Bp,
                          int d, int decoder, int sides, const float* A /*[sides][Bp][d], scalar path only*/, const float* dA, const float* gpos,
                          float* gcat, float* drel /*[sides][B][d] or null*/, cudaStream_t st) {
    if (B == 0) return MB_OK;
    const bool has_rel = (cols == 3) && (decoder != MB_DECODER_DOT) && rel != nullptr;
    const int dec = has_rel ? decoder : MB_DECODER_DOT;
    const bool vec_ok = decoder_vec_ok(emb, emb_ld, d, has_rel, rel, inv_rel, sides);
    if (vec_ok) {
        vec::EdgeBwdVArgs a;
        a.emb = emb;
        a.emb_ld = emb_ld;
        a.row_map = row_map;
        a.row_ptrs = row_ptrs;
        a.edges = edges;
        a.cols = cols;
        a.rel = rel;
        a.inv_rel = inv_rel;
        a.B = B;
        a.Bp = Bp;
        a.d = d;
        a.sides = sides;
        for (int s = 0; s < 2; s++) {
            const bool on = s < sides;
            a.dA[s] = on ? dA + s * Bp * d : nullptr;
            a.gpos[s] = on ? gpos + s * Bp : nullptr;
            a.drel[s] = (on && drel && has_rel) ? drel + s * B * d : nullptr;
        }
        a.gcat = gcat;
        const int grid = warp_grid(B);
        const bool small = d <= 128;
        if (dec == MB_DECODER_COMPLEX) {
            if (d <= 256)
                vec::edge_backward_kernel<MB_DECODER_COMPLEX, 1><<<grid, vec::kThreads, 0, st>>>(a);
            else
                vec::edge_backward_kernel<MB_DECODER_COMPLEX, 2><<<grid, vec::kThreads, 0, st>>>(a);
        } else if (dec == MB_DECODER_DISTMULT) {
            if (small)
                vec::edge_backward_kernel<MB_DECODER_DISTMULT, 1><<<grid, vec::kThreads, 0, st>>>(a);
            else
                vec::edge_backward_kernel<MB_DECODER_DISTMULT, 4><<<grid, vec::kThreads, 0, st>>>(a);
        } else {
            if (small)
                vec::edge_backward_kernel<MB_DECODER_DOT, 1><<<grid, vec::kThreads, 0, st>>>(a);
            else
                vec::edge_backward_kernel<MB_DECODER_DOT, 4><<<grid, vec::kThreads, 0, st>>>(a);
        }
        MB_LAUNCH_CHECK();
        return MB_OK;
    }
    if (row_map != nullptr || row_ptrs != nullptr) {
        set_error("launch_edge_bwd: row_map requires the vector path");
        return MB_ERR_INVALID;
    }
    return launch_edge_backward(emb, emb_ld, edges, cols, rel, sides == 2 ? inv_rel : nullptr, B, d, decoder, A, sides == 2 ? A + Bp * d : nullptr, dA,
                                sides == 2 ? dA + Bp * d : nullptr, gpos, sides == 2 ? gpos + Bp : nullptr, gcat, (drel && has_rel) ? drel : nullptr,
                                (drel && has_rel && sides == 2) ? drel + B * d : nullptr, st);
}

// sharded table: row_ptrs[u] for every unique row + the remote rows copied into `cache` [U,d] (see fetch_remote_rows_kernel)
mb_status launch_fetch_remote_rows(const mb_shards* sh, const int64_t* ids, int64_t U, int64_t ld, int d, float* cache, const float** row_ptrs,
                                   bool state_rows, cudaStream_t st) {
    if (U == 0) return MB_OK;
    if (sh == nullptr || sh->world < 1 || sh->rank < 0 || sh->rank >= sh->world || (d % 4) != 0 || d > 512 || (ld % 4) != 0 || !al16(cache)) {
        set_error("launch_fetch_remote_rows: bad shard description / unsupported row shape");
        return MB_ERR_INVALID;
    }
    vec::FetchArgs a{ids, U, make_sp(sh), ld, d, cache, row_ptrs};
    const int grid = warp_grid(U);
    if (state_rows) {
        if (d <= 128)
            vec::fetch_remote_rows_kernel<1, true><<<grid, vec::kThreads, 0, st>>>(a);
        else
            vec::fetch_remote_rows_kernel<4, true><<<grid, vec::kThreads, 0, st>>>(a);
    } else {
        if (d <= 128)
            vec::fetch_remote_rows_kernel<1, false><<<grid, vec::kThreads, 0, st>>>(a);
        else
            vec::fetch_remote_rows_kernel<4, false><<<grid, vec::kThreads, 0, st>>>(a);
    }
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_seg_reduce(const mb_shards* sh, int mode, const float* rows, const uint32_t* slots, const uint32_t* offsets, int64_t n_seg, int d, float* out, int64_t out_ld,
                            const float* state, int64_t state_ld, float* delta_e, float* delta_s, float* table, float* state_table, int64_t ld,
                            const int64_t* ids, float lr, cudaStream_t st, const int64_t* owner_bounds, int part) {
    if (n_seg == 0) return MB_OK;
    const bool vec_ok = (d % 4 == 0) && d <= 512 && al16(rows) && (!out || (al16(out) && out_ld % 4 == 0)) && (!state || (al16(state) && state_ld % 4 == 0)) &&
                        (!delta_e || (al16(delta_e) && al16(delta_s))) && (!table || (al16(table) && al16(state_table) && ld % 4 == 0));
    if (!vec_ok && sh != nullptr && sh->world > 1) {
        set_error("sharded update needs the vector kernels (d % 4 == 0, aligned tables)");
        return MB_ERR_UNSUPPORTED;
    }
    if (!vec_ok)
        return launch_segment_reduce(mode, rows, slots, offsets, n_seg, d, out, out_ld, state, state_ld, delta_e, delta_s, table, state_table, ld, ids, lr, st);
    vec::SegVArgs a{rows, slots, offsets, n_seg, d, out, out_ld, state, state_ld, delta_e, delta_s, table, state_table, ld, ids, -lr, make_sp(sh), owner_bounds};
    for (int i = 0; i < 8; i++) {
        a.inbox_ids[i] = nullptr;
        a.inbox_rows[i] = nullptr;
    }
    a.inbox_cap = 0;
    a.part = part;
    if (sh != nullptr && sh->world > 1) {
        if (mode == 2 && owner_bounds == nullptr) {
            set_error("sharded update needs the owner bounds of the batch");
            return MB_ERR_INVALID;
        }
        shard_inbox_ptrs(sh, d, a.inbox_ids, a.inbox_rows);
        a.inbox_cap = sh->exchange_rows;
    }
    int grid = warp_grid(n_seg);
#define MB_SEG(MODE)                                                                      \
    if (d <= 128)                                                                         \
        vec::segment_reduce_kernel<MODE, 1><<<grid, vec::kThreads, 0, st>>>(a);           \
    else                                                                                  \
        vec::segment_reduce_kernel<MODE, 4><<<grid, vec::kThreads, 0, st>>>(a);
    if (mode == 0) {
        MB_SEG(0)
    } else if (mode == 1) {
        MB_SEG(1)
    } else {
        MB_SEG(2)
    }
#undef MB_SEG
    MB_LAUNCH_CHECK();
    return MB_OK;
}

// relation gradients: rel_grad[r] = sum of the per-edge gradients of relation r (both relation tables in one launch)
mb_status launch_rel_reduce(const float* drel0, const float* drel1, float* out0, float* out1, const uint32_t* slots, const uint32_t* offsets, int64_t R,
                            int d, cudaStream_t st) {
    if (R == 0) return MB_OK;
    const int n_out = (out0 ? 1 : 0) + (out1 ? 1 : 0);
    if (n_out == 0) return MB_OK;
    const bool vec_ok = (d % 4 == 0) && al16(drel0) && (!drel1 || al16(drel1)) && (!out0 || al16(out0)) && (!out1 || al16(out1));
    if (!vec_ok) {
        if (out0) MB_TRY(launch_segment_reduce(0, drel0, slots, offsets, R, d, out0, d, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0.f, st));
        if (out1) MB_TRY(launch_segment_reduce(0, drel1, slots, offsets, R, d, out1, d, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0.f, st));
        return MB_OK;
    }
    vec::SegColArgs a;
    int k = 0;
    if (out0) {
        a.rows[k] = drel0;
        a.out[k++] = out0;
    }
    if (out1) {
        a.rows[k] = drel1;
        a.out[k++] = out1;
    }
    if (k == 1) {
        a.rows[1] = a.rows[0];
        a.out[1] = a.out[0];
    }
    a.slots = slots;
    a.offsets = offsets;
    a.n_seg = R;
    a.d = d;
    a.out_ld = d;
    const int ncb = ((d >> 2) + 31) >> 5;
    dim3 grid((unsigned)warp_grid(R * ncb), (unsigned)k);
    vec::segment_colsplit_kernel<<<grid, vec::kThreads, 0, st>>>(a);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
