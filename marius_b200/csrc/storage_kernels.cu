// storage_kernels.cu -- HBM-bound row kernels of the embedding table:
//   gather (Storage::indexRead), scatter-add / put (Storage::indexAdd / indexPut), the sparse Adagrad rule
//   (Batch::accumulateGradients) and its fusion with the two scatter-adds, getGlobalToLocalMap, dense Adagrad.
//
// Layout: table[num_rows][ld] fp32 row-major, rows of d floats (1600 B at d=400, 16 B aligned when d % 4 == 0).
// Mapping: a flat index v over (row, 16-byte column) pairs, so consecutive lanes touch consecutive 16 B of the
// same row (and roll over into the next index's row): every 128-byte line of every touched row is moved by
// exactly one fully-used request.  Each thread keeps UNROLL independent rows in flight (loads first, then
// stores) to cover the ~600-cycle HBM latency; the grid is a multiple of the SM count and grid-strides.
// Algorithmic bytes: gather 8*n*d (read row + write staged row); scatter-add 12*n*d; fused Adagrad update
// 20*n*d (read g,e,s ; write e,s).
#include "common.cuh"

namespace mb {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <typename T>
struct Vec;
template <>
struct Vec<float4> {
    static constexpr int W = 4;
};
template <>
struct Vec<float> {
    static constexpr int W = 1;
};

__device__ __forceinline__ float4 vload(const float4* p) { return ld_stream(p); }
__device__ __forceinline__ float vload(const float* p) { return __ldg(p); }
__device__ __forceinline__ void vstore(float4* p, const float4& v) { st_stream(p, v); }
__device__ __forceinline__ void vstore(float* p, float v) { *p = v; }
__device__ __forceinline__ float4 vadd(const float4& a, const float4& b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }

// Flat (row, vec-column) iterator: position v = row * dv + col advanced by a fixed stride without divisions.
struct FlatIter {
    int64_t v, total, row;
    int col, dv, srow_col;
    int64_t srow, stride;
    __device__ FlatIter(int64_t n_rows, int dv_) {
        dv = dv_;
        total = n_rows * dv;
        stride = (int64_t)gridDim.x * blockDim.x;
        v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        row = v / dv;
        col = (int)(v - row * dv);
        srow = stride / dv;
        srow_col = (int)(stride - srow * dv);
    }
    __device__ __forceinline__ bool valid() const { return v < total; }
    __device__ __forceinline__ void next() {
        v += stride;
        row += srow;
        col += srow_col;
        if (col >= dv) {
            col -= dv;
            row += 1;
        }
    }
};

// out[r, :] = table[idx[r], :]
template <typename V>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const float* __restrict__ table, int64_t ld, const int64_t* __restrict__ idx, int64_t n,
                                                               int dv, float* __restrict__ out, int64_t out_ld) {
    constexpr int W = Vec<V>::W;
    FlatIter it(n, dv);
    while (it.valid()) {
        V val[kUnroll];
        int64_t orow[kUnroll];
        int ocol[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            ok[u] = it.valid();
            orow[u] = it.row;
            ocol[u] = it.col;
            if (ok[u]) {
                int64_t src = __ldg(idx + it.row);
                // a negative id is padding (mb_edge_sample / mb_reduce_rows_by_key fill unused capacity with -1): reads as a zero row
                val[u] = src >= 0 ? vload(reinterpret_cast<const V*>(table + src * ld) + it.col) : V{};
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            if (ok[u]) vstore(reinterpret_cast<V*>(out + orow[u] * out_ld) + ocol[u], val[u]);
        }
    }
    (void)W;
}

// MODE 0: table[idx[r], :] += vals[r, :]   (unique idx: plain read-modify-write, no atomics -- buffer.cpp:459)
// MODE 1: table[idx[r], :]  = vals[r, :]
template <typename V, int MODE>
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(float* __restrict__ table, int64_t ld, const int64_t* __restrict__ idx, int64_t n, int dv,
                                                                const float* __restrict__ vals, int64_t vals_ld) {
    FlatIter it(n, dv);
    while (it.valid()) {
        V a[kUnroll], b[kUnroll];
        V* dst[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            ok[u] = it.valid();
            if (ok[u]) {
                int64_t r = __ldg(idx + it.row);
                dst[u] = reinterpret_cast<V*>(table + r * ld) + it.col;
                b[u] = vload(reinterpret_cast<const V*>(vals + it.row * vals_ld) + it.col);
                if (MODE == 0) a[u] = *dst[u];
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            if (ok[u]) {
                if (MODE == 0)
                    vstore(dst[u], vadd(a[u], b[u]));
                else
                    vstore(dst[u], b[u]);
            }
        }
    }
}

// Batch::accumulateGradients: delta_e, delta_s from (grad, gathered state)
__global__ void __launch_bounds__(kThreads) adagrad_deltas_kernel(const float* __restrict__ grad, const float* __restrict__ state, int64_t n, int d,
                                                                  int64_t ld, float neg_lr, float* __restrict__ de, float* __restrict__ ds) {
    int64_t total = n * d;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = v / d;
        int c = (int)(v - r * d);
        float e, s, sn;
        adagrad_rule(grad[r * ld + c], state[r * ld + c], neg_lr, e, s, sn);
        de[r * ld + c] = e;
        ds[r * ld + c] = s;
    }
}

__device__ __forceinline__ void adagrad_apply(float g, float& e, float& s, float neg_lr) {
    float de, ds, sn;
    adagrad_rule(g, s, neg_lr, de, ds, sn);
    e = __fadd_rn(e, de);  // Storage::indexAdd of node_gradients_   (storage.cpp:656-657)
    s = sn;                // Storage::indexAdd of node_state_update_ : s + g*g == s'
}

// accumulateGradients + indexAdd(emb) + indexAdd(state) fused: one RMW per unique row of each table.
template <typename V>
__global__ void __launch_bounds__(kThreads) adagrad_update_rows_kernel(float* __restrict__ table, float* __restrict__ state_table, int64_t ld,
                                                                       const int64_t* __restrict__ idx, int64_t n, int dv,
                                                                       const float* __restrict__ grad, int64_t grad_ld, float neg_lr) {
    FlatIter it(n, dv);
    while (it.valid()) {
        V g[kUnroll], e[kUnroll], s[kUnroll];
        V *pe[kUnroll], *ps[kUnroll];
        bool ok[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            ok[u] = it.valid();
            if (ok[u]) {
                int64_t r = __ldg(idx + it.row);
                if (r < 0) {  // padding entry (mb_reduce_rows_by_key pads its unique list with -1): nothing to update
                    ok[u] = false;
                    it.next();
                    continue;
                }
                pe[u] = reinterpret_cast<V*>(table + r * ld) + it.col;
                ps[u] = reinterpret_cast<V*>(state_table + r * ld) + it.col;
                g[u] = vload(reinterpret_cast<const V*>(grad + it.row * grad_ld) + it.col);
                e[u] = *pe[u];
                s[u] = *ps[u];
            }
            it.next();
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            if (ok[u]) {
                if constexpr (Vec<V>::W == 4) {
                    adagrad_apply(g[u].x, e[u].x, s[u].x, neg_lr);
                    adagrad_apply(g[u].y, e[u].y, s[u].y, neg_lr);
                    adagrad_apply(g[u].z, e[u].z, s[u].z, neg_lr);
                    adagrad_apply(g[u].w, e[u].w, s[u].w, neg_lr);
                } else {
                    adagrad_apply(g[u], e[u], s[u], neg_lr);
                }
                vstore(pe[u], e[u]);
                vstore(ps[u], s[u]);
            }
        }
    }
}

__global__ void fill_i64_kernel(int64_t* p, int64_t n, int64_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

struct ResidentList {
    int32_t part[64];
    int32_t slot[64];
};

__global__ void g2l_map_kernel(int64_t* map, int64_t total_rows, int64_t psize, ResidentList rl, int n_res) {
    // one block-row per resident partition
    int p = blockIdx.y;
    if (p >= n_res) return;
    int64_t lo = (int64_t)rl.part[p] * psize;
    int64_t hi = lo + psize < total_rows ? lo + psize : total_rows;
    int64_t base = (int64_t)rl.slot[p] * psize;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) map[i] = base + (i - lo);
}

// AdagradOptimizer::step (nn/optim.cpp:114-145): sum += g*g ; p += -lr * g / (sqrt(sum) + eps)
__global__ void dense_adagrad_kernel(float* __restrict__ p, float* __restrict__ sum, const float* __restrict__ g, int64_t n, float lr, float eps) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        float s = fmaf(gi, gi, sum[i]);  // addcmul_
        sum[i] = s;
        float den = __fadd_rn(__fsqrt_rn(s), eps);
        p[i] = fmaf(-lr, __fdiv_rn(gi, den), p[i]);  // addcdiv_
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int grid_for(int64_t work_items) {
    int64_t per_block = (int64_t)kThreads * kUnroll;
    int64_t blocks = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)sm_count() * 8;  // 8 x 256 threads = 2048 resident threads / SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

mb_status gather_rows(const float* table, int64_t ld, int64_t d, const int64_t* idx, int64_t n, float* out, int64_t out_ld, cudaStream_t st) {
    if (n == 0 || d == 0) return MB_OK;
    bool v4 = (d % 4 == 0) && (ld % 4 == 0) && (out_ld % 4 == 0) && aligned16(table) && aligned16(out);
    if (v4) {
        int dv = (int)(d / 4);
        gather_rows_kernel<float4><<<grid_for(n * dv), kThreads, 0, st>>>(table, ld, idx, n, dv, out, out_ld);
    } else {
        gather_rows_kernel<float><<<grid_for(n * d), kThreads, 0, st>>>(table, ld, idx, n, (int)d, out, out_ld);
    }
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status scatter_rows(float* table, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* vals, int64_t vals_ld, bool add,
                       cudaStream_t st) {
    if (n == 0 || d == 0) return MB_OK;
    bool v4 = (d % 4 == 0) && (ld % 4 == 0) && (vals_ld % 4 == 0) && aligned16(table) && aligned16(vals);
    if (v4) {
        int dv = (int)(d / 4);
        if (add)
            scatter_rows_kernel<float4, 0><<<grid_for(n * dv), kThreads, 0, st>>>(table, ld, idx, n, dv, vals, vals_ld);
        else
            scatter_rows_kernel<float4, 1><<<grid_for(n * dv), kThreads, 0, st>>>(table, ld, idx, n, dv, vals, vals_ld);
    } else {
        if (add)
            scatter_rows_kernel<float, 0><<<grid_for(n * d), kThreads, 0, st>>>(table, ld, idx, n, (int)d, vals, vals_ld);
        else
            scatter_rows_kernel<float, 1><<<grid_for(n * d), kThreads, 0, st>>>(table, ld, idx, n, (int)d, vals, vals_ld);
    }
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status adagrad_deltas(const float* grad, const float* state, int64_t n, int64_t d, int64_t ld, float lr, float* de, float* ds, cudaStream_t st) {
    if (n == 0 || d == 0) return MB_OK;
    adagrad_deltas_kernel<<<grid_for(n * d), kThreads, 0, st>>>(grad, state, n, (int)d, ld, -lr, de, ds);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status adagrad_update_rows(float* table, float* state_table, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* grad,
                              int64_t grad_ld, float lr, cudaStream_t st) {
    if (n == 0 || d == 0) return MB_OK;
    bool v4 = (d % 4 == 0) && (ld % 4 == 0) && (grad_ld % 4 == 0) && aligned16(table) && aligned16(state_table) && aligned16(grad);
    if (v4) {
        int dv = (int)(d / 4);
        adagrad_update_rows_kernel<float4><<<grid_for(n * dv), kThreads, 0, st>>>(table, state_table, ld, idx, n, dv, grad, grad_ld, -lr);
    } else {
        adagrad_update_rows_kernel<float><<<grid_for(n * d), kThreads, 0, st>>>(table, state_table, ld, idx, n, (int)d, grad, grad_ld, -lr);
    }
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status global_to_local_map(int64_t* map, int64_t total_rows, int64_t psize, const int32_t* part_ids, const int32_t* slots, int n_res,
                              cudaStream_t st) {
    if (n_res > 64) {
        set_error("global_to_local_map: more than 64 resident partitions");
        return MB_ERR_UNSUPPORTED;
    }
    fill_i64_kernel<<<grid_for(total_rows), kThreads, 0, st>>>(map, total_rows, -1);
    MB_LAUNCH_CHECK();
    if (n_res > 0) {
        ResidentList rl;
        for (int i = 0; i < n_res; i++) {
            rl.part[i] = part_ids[i];
            rl.slot[i] = slots[i];
        }
        dim3 grid((unsigned)((psize + kThreads - 1) / kThreads > 64 ? 64 : (psize + kThreads - 1) / kThreads), (unsigned)n_res);
        g2l_map_kernel<<<grid, kThreads, 0, st>>>(map, total_rows, psize, rl, n_res);
        MB_LAUNCH_CHECK();
    }
    return MB_OK;
}

mb_status dense_adagrad_step(float* p, float* sum, const float* g, int64_t n, float lr, float eps, cudaStream_t st) {
    if (n == 0) return MB_OK;
    dense_adagrad_kernel<<<grid_for(n), kThreads, 0, st>>>(p, sum, g, n, lr, eps);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
