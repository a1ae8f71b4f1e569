// shard_kernels.cu -- the exchange step of the node-partition-sharded table: device-side flag barriers between the
// ranks of one box, owner bounds of a batch's sorted unique ids, and the owner-side Adagrad apply of the gradient rows a rank received.
//
// Exchange area of one owner (mb_shard_exchange_bytes), zero-initialised by the caller, mapped into every peer process (CUDA IPC):
//   [0, 64)      flags[8]   uint64: flags[s] = number of barriers rank s has entered (written by rank s, st.release.sys)
//   [64, 72)     epoch      uint64: number of barriers THIS rank has entered (only this rank touches it)
//   [72, 76)     error      int   : sticky, set when a barrier timed out
//   [256, ...)   inbox[s], s = 0..world-1, stride = inbox_stride(rows, d):
//                  [0, 8)                     count: gradient rows sender s wrote this step
//                  [256, 256 + 8 rows)        global ids of those rows
//                  [256 + align(8 rows), ..)  the gradient rows, [rows][d] fp32
#include "kernels.h"

namespace mb {

namespace {

constexpr int64_t kExHeader = 256;
__host__ __device__ inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }
__host__ __device__ inline int64_t inbox_stride(int64_t rows, int64_t d) { return 256 + align256(8 * rows) + align256(4 * rows * d); }

struct ExPtrs {
    char* area[8];
    int world, rank;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One warp.  Everything this rank wrote before (kernels earlier on the stream: inbox rows in peer memory, its own table rows) is made
// visible system-wide, then the rank's barrier count is published in every peer's flag array and the warp waits until every peer has
// published the same count in ours.  ~2 s without progress: give up and raise the sticky error flag (never hang the GPU).
__global__ void shard_barrier_kernel(ExPtrs ex) {
    const int lane = threadIdx.x;
    char* me = ex.area[ex.rank];
    unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(me);
    unsigned long long* my_epoch = reinterpret_cast<unsigned long long*>(me + 64);
    unsigned long long epoch = 0;
    if (lane == 0) {
        epoch = *my_epoch + 1;
        *my_epoch = epoch;
        __threadfence_system();
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    if (lane < ex.world && lane != ex.rank) {
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(ex.area[lane]) + ex.rank, epoch);
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(my_flags + lane) < epoch) {
            if (clock64() - t0 > 4000000000ll) {
                *reinterpret_cast<int*>(me + 72) = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    __syncwarp();
    __threadfence_system();
}

// bounds[o] = first position of the (ascending, -1 padded) unique-id list whose row belongs to owner >= o; the number of rows this
// rank will ship to owner o is published in o's inbox header right away.
__global__ void owner_bounds_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t rows_per_rank, ExPtrs ex, int64_t ex_rows, int64_t d,
                                    int64_t* __restrict__ bounds) {
    const int o = threadIdx.x;
    if (o > ex.world) return;
    const unsigned long long key = (unsigned long long)o * (unsigned long long)rows_per_rank;
    int64_t lo = 0, hi = n;
    while (lo < hi) {  // padding entries are negative: as unsigned they sort after every valid id
        const int64_t mid = (lo + hi) >> 1;
        if ((unsigned long long)ids[mid] < key) lo = mid + 1; else hi = mid;
    }
    bounds[o] = lo;
    __syncthreads();
    if (o < ex.world && o != ex.rank) {
        const int64_t cnt = bounds[o + 1] - bounds[o];
        *reinterpret_cast<int64_t*>(ex.area[o] + kExHeader + ex.rank * inbox_stride(ex_rows, d)) = cnt < ex_rows ? cnt : ex_rows;
    }
}

__device__ __forceinline__ float4 ldg_f4(const float* p, int v) {
    float4 r;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(reinterpret_cast<const float4*>(p) + v));
    return r;
}

// Owner side: one sparse-Adagrad step (batch.cpp:62-79 + both indexAdds, dataloader.cpp:550-564) with the gradient rows one sender
// left in this rank's inbox.  One warp per row, two rows in flight; ids within one inbox are unique, so plain read-modify-write.
template <int CH>
__global__ void __launch_bounds__(256) inbox_apply_kernel(const char* __restrict__ inbox, int64_t ex_rows, float* __restrict__ table, float* __restrict__ state,
                                                           int64_t ld, int d, int64_t row0, int64_t rows_per_rank, float neg_lr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    const int64_t count = *reinterpret_cast<const int64_t*>(inbox);
    const int64_t* ids = reinterpret_cast<const int64_t*>(inbox + 256);
    const float* rows = reinterpret_cast<const float*>(inbox + 256 + align256(8 * ex_rows));
    const int dv = d >> 2;
    for (int64_t i = warp0; i < count; i += nwarps) {
        const int64_t l = ids[i] - row0;
        if (l < 0 || l >= rows_per_rank) continue;  // (never: the sender routes by owner)
        float* e = table + l * ld;
        float* s = state + l * ld;
        const float* g = rows + i * d;
        float4 ge[CH], ee[CH], ss[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int vc = min(lane + 32 * c, dv - 1);
            ge[c] = ldg_f4(g, vc);
            ee[c] = ldg_f4(e, vc);
            ss[c] = ldg_f4(s, vc);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int v = lane + 32 * c;
            if (v >= dv) continue;
            float4 de, ds, sn;
            adagrad_rule(ge[c].x, ss[c].x, neg_lr, de.x, ds.x, sn.x);
            adagrad_rule(ge[c].y, ss[c].y, neg_lr, de.y, ds.y, sn.y);
            adagrad_rule(ge[c].z, ss[c].z, neg_lr, de.z, ds.z, sn.z);
            adagrad_rule(ge[c].w, ss[c].w, neg_lr, de.w, ds.w, sn.w);
            st_stream(reinterpret_cast<float4*>(e) + v, make_float4(__fadd_rn(ee[c].x, de.x), __fadd_rn(ee[c].y, de.y), __fadd_rn(ee[c].z, de.z), __fadd_rn(ee[c].w, de.w)));
            st_stream(reinterpret_cast<float4*>(s) + v, sn);
        }
    }
}

ExPtrs make_ex(const mb_shards* sh) {
    ExPtrs ex;
    for (int i = 0; i < 8; i++) ex.area[i] = i < sh->world ? static_cast<char*>(sh->exchange[i]) : nullptr;
    ex.world = sh->world;
    ex.rank = sh->rank;
    return ex;
}

}  // namespace

int64_t shard_exchange_bytes(int world, int64_t rows, int64_t d) { return kExHeader + (int64_t)world * inbox_stride(rows, d); }

mb_status launch_shard_barrier(const mb_shards* sh, cudaStream_t st) {
    if (sh->single_process) return MB_OK;
    shard_barrier_kernel<<<1, 32, 0, st>>>(make_ex(sh));
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status launch_owner_bounds(const mb_shards* sh, const int64_t* ids, int64_t n, int64_t d, int64_t* bounds, cudaStream_t st) {
    owner_bounds_kernel<<<1, 32, 0, st>>>(ids, n, sh->rows_per_rank, make_ex(sh), sh->exchange_rows, d, bounds);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

// this rank's sender slot in owner o's exchange area
void shard_inbox_ptrs(const mb_shards* sh, int64_t d, int64_t** ids_out, float** rows_out) {
    for (int o = 0; o < 8; o++) {
        ids_out[o] = nullptr;
        rows_out[o] = nullptr;
        if (o >= sh->world) continue;
        char* inbox = static_cast<char*>(sh->exchange[o]) + kExHeader + sh->rank * inbox_stride(sh->exchange_rows, d);
        ids_out[o] = reinterpret_cast<int64_t*>(inbox + 256);
        rows_out[o] = reinterpret_cast<float*>(inbox + 256 + align256(8 * sh->exchange_rows));
    }
}

// apply the inbox sender `sender` left at owner `owner` (owner == sh->rank in the one-process-per-GPU deployment)
mb_status launch_inbox_apply(const mb_shards* sh, int owner, int sender, int64_t ld, int d, float lr, int64_t max_rows, cudaStream_t st) {
    const char* inbox = static_cast<const char*>(sh->exchange[owner]) + kExHeader + sender * inbox_stride(sh->exchange_rows, d);
    int64_t warps = max_rows < 1 ? 1 : max_rows;
    int grid = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)sm_count() * 4);
    if (d <= 128)
        inbox_apply_kernel<1><<<grid, 256, 0, st>>>(inbox, sh->exchange_rows, sh->tables[owner], sh->states[owner], ld, d, (int64_t)owner * sh->rows_per_rank,
                                                    sh->rows_per_rank, -lr);
    else
        inbox_apply_kernel<4><<<grid, 256, 0, st>>>(inbox, sh->exchange_rows, sh->tables[owner], sh->states[owner], ld, d, (int64_t)owner * sh->rows_per_rank,
                                                    sh->rows_per_rank, -lr);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

mb_status shard_error_flag(const mb_shards* sh, int* out) {
    int v = 0;
    MB_CUDA_TRY(cudaMemcpy(&v, static_cast<const char*>(sh->exchange[sh->rank]) + 72, sizeof(int), cudaMemcpyDeviceToHost));
    *out = v;
    return MB_OK;
}

}  // namespace mb
