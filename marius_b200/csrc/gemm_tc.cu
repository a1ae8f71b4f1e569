// gemm_tc.cu -- tcgen05 (5th-gen tensor core) batched GEMM for the only dense contraction on the Marius hot path:
//   scores  S  = A . Neg^T     (DotCompare bmm, comparators.cpp:69-72)          A K-major,  B K-major
//   dA         = G . Neg       (bmm backward w.r.t. the adjusted positives)     A K-major,  B MN-major
//   dNeg       = G^T . A       (bmm backward w.r.t. the negative rows)          A MN-major, B MN-major
// fp32 semantics on a bf16 tensor pipe: every fp32 operand x is pre-split into bf16 hi = rn(x), lo = rn(x - hi) and the
// kernel accumulates  hi.hi + hi.lo + lo.hi  into one fp32 TMEM accumulator (3 UMMA products per k-step, |rel err| ~ 2^-16;
// the dropped lo.lo term is ~2^-18).  MB_PREC_BF16 issues only hi.hi.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps: inner, rows, batch) -> SWIZZLE_128B smem tiles, mbarrier tx-count
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, M=128, N=BLOCK_N, K=16), tcgen05.commit -> mbarriers
//   warps 2..5  epilogue: tcgen05.ld 32x32b (each warp owns its TMEM lane quarter) -> registers -> fp32 rows in global
//   two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1.
// OOB handling is TMA zero fill: ragged M / N / K tiles need no special cases in the MMA loop (zero rows/columns contribute 0).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "gemm_tc_ptx.cuh"

namespace mb {

namespace {

using namespace tcptx;
// BLOCK_K (template parameter, bf16 elements per k-block): 64 -> K-major tiles are SWIZZLE_128B rows, 32 -> SWIZZLE_64B rows
// (half the bytes per stage => twice the pipeline depth in the same shared memory).  MN-major tiles are always SWIZZLE_128B
// slabs of 64 MN-elements x BLOCK_K k-rows.

struct TcParams {
    float* D;
    int64_t ldd;      // elements
    int64_t sDb;      // batch stride (elements)
    int M, N, K, batches;
    int passes;       // 1 (hi.hi) or 3 (hi.hi + hi.lo + lo.hi)
    int m_tiles, n_tiles;
    int debug_flags;  // bit 0: skip the epilogue's global stores (bandwidth experiments only; MB_TC_DEBUG)
    int tma_store;    // 1: epilogue stores through shared memory + TMA (needs 16-byte aligned D, ldd % 4 == 0)
};

template <int BLOCK_N, int BLOCK_K, int STAGES>
struct SmemLayout {
    static constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + kEpilogueSmemBytes;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
    static_assert(TOTAL <= 232448, "shared memory budget (227 KB)");
};

template <int BLOCK_N, int BLOCK_K, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, const __grid_constant__ CUtensorMap tmB_hi,
               const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmD, const TcParams p) {
    using L = SmemLayout<BLOCK_N, BLOCK_K, STAGES>;
    constexpr int A_TILE_BYTES = L::A_TILE_BYTES;
    constexpr int TMEM_COLS = 2 * BLOCK_N;  // double-buffered fp32 accumulator: 256 or 512 columns (power of two)
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512 || TMEM_COLS == 128, "TMEM columns must be a power of two");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_holder = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int tiles_per_batch = p.m_tiles * p.n_tiles;
    const int num_tiles = tiles_per_batch * p.batches;
    const uint32_t stage_tx = (uint32_t)((p.passes == 3 ? 2 : 1) * (A_TILE_BYTES + L::B_TILE_BYTES));

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi);
        prefetch_tmap(&tmB_hi);
        if (p.passes == 3) {
            prefetch_tmap(&tmA_lo);
            prefetch_tmap(&tmB_lo);
        }
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc<TMEM_COLS>(tmem_holder);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int b = t / tiles_per_batch;
                const int rem = t - b * tiles_per_batch;
                const int m0 = (rem / p.n_tiles) * BLOCK_M;
                const int n0 = (rem % p.n_tiles) * BLOCK_N;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA_hi = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sA_lo = sA_hi + A_TILE_BYTES;
                    const uint32_t sB_hi = sA_lo + A_TILE_BYTES;
                    const uint32_t sB_lo = sB_hi + L::B_TILE_BYTES;
                    if (p.debug_flags & 2) {  // experiment: no TMA traffic, the MMAs run on whatever is in shared memory
                        mbar_arrive(full_bar(stage));
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    mbar_expect_tx(full_bar(stage), stage_tx);
                    const int k0 = kb * BLOCK_K;
                    if (!A_MN) {
                        tma_load_3d(sA_hi, &tmA_hi, full_bar(stage), k0, m0, b);
                        if (p.passes == 3) tma_load_3d(sA_lo, &tmA_lo, full_bar(stage), k0, m0, b);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; j++) {
                            tma_load_3d(sA_hi + j * (BLOCK_K * 128), &tmA_hi, full_bar(stage), m0 + 64 * j, k0, b);
                            if (p.passes == 3) tma_load_3d(sA_lo + j * (BLOCK_K * 128), &tmA_lo, full_bar(stage), m0 + 64 * j, k0, b);
                        }
                    }
                    if (!B_MN) {
                        tma_load_3d(sB_hi, &tmB_hi, full_bar(stage), k0, n0, b);
                        if (p.passes == 3) tma_load_3d(sB_lo, &tmB_lo, full_bar(stage), k0, n0, b);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; j++) {
                            tma_load_3d(sB_hi + j * (BLOCK_K * 128), &tmB_hi, full_bar(stage), n0 + 64 * j, k0, b);
                            if (p.passes == 3) tma_load_3d(sB_lo + j * (BLOCK_K * 128), &tmB_lo, full_bar(stage), n0 + 64 * j, k0, b);
                        }
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            // K-major: rows of BLOCK_K*2 bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), 8-row groups SBO = 8 rows apart, LBO unused
            //          (=16 B); k-step = +32 B inside the swizzle span.
            // MN-major SW128: 64-element MN slabs BLOCK_K*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO); k-step = +2048 B.
            constexpr uint32_t K_ROW_BYTES = BLOCK_K * 2;
            constexpr uint32_t K_LAYOUT = (K_ROW_BYTES == 128) ? 2u : 4u;
            constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 16, A_SBO = A_MN ? 1024 : 8 * K_ROW_BYTES, A_KSTEP = A_MN ? 2048 : 32;
            constexpr uint32_t B_LBO = B_MN ? BLOCK_K * 128 : 16, B_SBO = B_MN ? 1024 : 8 * K_ROW_BYTES, B_KSTEP = B_MN ? 2048 : 32;
            constexpr uint32_t A_LAYOUT = A_MN ? 2u : K_LAYOUT, B_LAYOUT = B_MN ? 2u : K_LAYOUT;
            int stage = 0;
            uint32_t phase = 0;
            int local_tile = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, local_tile++) {
                const int buf = local_tile & 1;
                const uint32_t buf_phase = (uint32_t)((local_tile >> 1) & 1);
                // ragged last N tile: shrink the instruction's N (multiple of 16) so no tensor cycles are spent on zero columns
                const int n0 = ((t % tiles_per_batch) % p.n_tiles) * BLOCK_N;
                const int n_eff = min(BLOCK_N, ((p.N - n0 + 15) / 16) * 16);
                const uint32_t idesc = make_idesc(BLOCK_M, n_eff, A_MN, B_MN);
                mbar_wait(tempty_bar(buf), buf_phase ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BLOCK_N);
                uint32_t accumulate = 0;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sA_hi = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sA_lo = sA_hi + A_TILE_BYTES;
                    const uint32_t sB_hi = sA_lo + A_TILE_BYTES;
                    const uint32_t sB_lo = sB_hi + L::B_TILE_BYTES;
                    const int k_valid = min(BLOCK_K, p.K - kb * BLOCK_K);
                    const int ksteps = (k_valid + UMMA_K - 1) / UMMA_K;
                    for (int prod = 0; prod < ((p.debug_flags & 4) ? 0 : p.passes); prod++) {
                        const uint32_t sa = (prod == 2) ? sA_lo : sA_hi;  // hi.hi, hi.lo, lo.hi
                        const uint32_t sb = (prod == 1) ? sB_lo : sB_hi;
                        for (int ks = 0; ks < ksteps; ks++) {
                            uint64_t adesc = make_smem_desc(sa + ks * A_KSTEP, A_LBO, A_SBO, A_LAYOUT);
                            uint64_t bdesc = make_smem_desc(sb + ks * B_KSTEP, B_LBO, B_SBO, B_LAYOUT);
                            umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(buf));  // accumulator complete -> epilogue
            }
        }
    } else {
        // ================= epilogue warps 2..5 =================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int local_tile = 0;
        uint32_t epi_chunk = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, local_tile++) {
            const int buf = local_tile & 1;
            const uint32_t buf_phase = (uint32_t)((local_tile >> 1) & 1);
            const int b = t / tiles_per_batch;
            const int rem = t - b * tiles_per_batch;
            const int m0 = (rem / p.n_tiles) * BLOCK_M;
            const int n0 = (rem % p.n_tiles) * BLOCK_N;
            mbar_wait(tfull_bar(buf), buf_phase);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            float* drow = p.D + (int64_t)b * p.sDb + (int64_t)row * p.ldd;
            const bool row_ok = row < p.M && !(p.debug_flags & 1);
            const bool vec_ok = ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15u) == 0) && ((p.sDb & 3) == 0);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; c++) {
                const int col0 = n0 + c * 32;
                if (col0 >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BLOCK_N + c * 32), r);
                tmem_ld_wait();
                if (p.tma_store) {
                    if (!(p.debug_flags & 1))
                        stage_and_store(r, smem_base + L::EPI_OFFSET + (uint32_t)((warp - 2) * 2 + (epi_chunk & 1)) * kStageTileBytes, lane, &tmD, col0,
                                        m0 + q * 32, b);
                    epi_chunk++;
                } else if (row_ok) {
                    if (vec_ok && col0 + 32 <= p.N) {
                        float4* dst = reinterpret_cast<float4*>(drow + col0);
#pragma unroll
                        for (int v = 0; v < 8; v++)
                            dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                                 __uint_as_float(r[4 * v + 3]));
                    } else {
#pragma unroll
                        for (int v = 0; v < 32; v++)
                            if (col0 + v < p.N) drow[col0 + v] = __uint_as_float(r[v]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        if (p.tma_store && elect_one()) bulk_wait_all();  // all bulk stores of this warp have completed before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---- host side: tensor maps -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// 3-D bf16 map: dims {inner, rows, batches}; box {box_inner (64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B), box_rows, 1}; zero OOB fill.
mb_status make_map(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches, uint64_t row_stride_elems,
                   uint64_t batch_stride_elems, uint32_t box_rows, uint32_t box_inner = 64) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return MB_ERR_CUDA;
    }
    cuuint64_t dims[3] = {inner, rows, batches};
    cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
    if (batches == 1) strides[1] = strides[0] * rows;
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (strides[0] & 15u) || (strides[1] & 15u)) {
        set_error("gemm_tc: operand not 16-byte aligned / stride not a multiple of 16 bytes");
        return MB_ERR_INVALID;
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return MB_ERR_CUDA;
    }
    return MB_OK;
}

template <int BLOCK_N, int BLOCK_K, int STAGES, bool A_MN, bool B_MN>
mb_status launch_variant(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo, const CUtensorMap& d_map,
                         const TcParams& p, cudaStream_t st) {
    using L = SmemLayout<BLOCK_N, BLOCK_K, STAGES>;
    auto kern = gemm_tc_kernel<BLOCK_N, BLOCK_K, STAGES, A_MN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        MB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    int tiles = p.m_tiles * p.n_tiles * p.batches;
    int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, kTcThreads, L::TOTAL, st>>>(a_hi, a_lo, b_hi, b_lo, d_map, p);
    MB_LAUNCH_CHECK();
    return MB_OK;
}


// =================================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs (one TPC) computes a 256 x 256 tile.  Each CTA stages ITS 128 rows
// of A and ITS 128-column half of B; the leader's single thread issues tcgen05.mma.cta_group::2 (M = 256), the hardware
// reads the B halves from both CTAs' shared memory and each CTA accumulates its 128 rows in its own TMEM.  Per SM and
// k-block this moves 2/3 of the bytes of the 1-CTA kernel through L2 -> SMEM and through the SMEM read port (the 1-CTA
// kernel needs ~96 B/clk of operand reads + ~62 B/clk of TMA fill against a 128 B/clk port: DESIGN.md 4).
//   - TMA loads use .cta_group::2 and signal the LEADER's full barrier (peer bit 24 of the shared::cluster address cleared)
//   - tcgen05.commit .multicast::cluster releases the stage in both CTAs and hands the accumulator to both epilogues
//   - the peer's epilogue warps arrive remotely (mapa) on the leader's tmem-empty barrier
// =================================================================================================================
template <int BLOCK_K, int STAGES>
struct SmemLayout2 {
    static constexpr int HALF_N = 128;                         // this CTA's share of the 256 tile columns
    static constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // this CTA's 128 rows of A
    static constexpr int B_TILE_BYTES = HALF_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + kEpilogueSmemBytes;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static_assert(TOTAL <= 232448, "shared memory budget (227 KB)");
};

template <int BLOCK_K, int STAGES, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo, const __grid_constant__ CUtensorMap tmB_hi,
                const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmD, const TcParams p) {
    using L = SmemLayout2<BLOCK_K, STAGES>;
    constexpr int TILE_M = 256, TILE_N = 256, HALF_N = L::HALF_N;
    constexpr int A_TILE_BYTES = L::A_TILE_BYTES, B_TILE_BYTES = L::B_TILE_BYTES;
    constexpr int TMEM_COLS = 2 * TILE_N;  // 512: double-buffered 128-lane x 256-column accumulator per CTA
    constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // shared::cluster address of the even (leader) CTA of the pair
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
    const uint32_t tmem_holder = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int tiles_per_batch = p.m_tiles * p.n_tiles;
    const int num_tiles = tiles_per_batch * p.batches;
    const uint32_t stage_tx_pair = (uint32_t)(2 * (p.passes == 3 ? 2 : 1) * (A_TILE_BYTES + B_TILE_BYTES));  // both CTAs' bytes

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi);
        prefetch_tmap(&tmB_hi);
        if (p.passes == 3) {
            prefetch_tmap(&tmA_lo);
            prefetch_tmap(&tmB_lo);
        }
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 8);  // 4 epilogue warps x 2 CTAs (only the leader's copy is used)
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barrier inits of both CTAs are visible before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder_ptr;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int b = t / tiles_per_batch;
                const int rem = t - b * tiles_per_batch;
                const int m0 = (rem / p.n_tiles) * TILE_M + (int)rank * BLOCK_M;
                const int n_tile0 = (rem % p.n_tiles) * TILE_N;
                const int n_eff = min(TILE_N, ((p.N - n_tile0 + 31) / 32) * 32);  // multiple of 32: each CTA's half is a multiple of 16
                const int n0 = n_tile0 + (int)rank * (n_eff / 2);
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA_hi = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sA_lo = sA_hi + A_TILE_BYTES;
                    const uint32_t sB_hi = sA_lo + A_TILE_BYTES;
                    const uint32_t sB_lo = sB_hi + B_TILE_BYTES;
                    const uint32_t lbar = full_bar(stage) & kPeerMask;
                    if (p.debug_flags & 2) {
                        if (leader) mbar_arrive(full_bar(stage));
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    if (leader) mbar_expect_tx(full_bar(stage), stage_tx_pair);
                    const int k0 = kb * BLOCK_K;
                    if (!A_MN) {
                        tma_load_3d_2sm(sA_hi, &tmA_hi, lbar, k0, m0, b);
                        if (p.passes == 3) tma_load_3d_2sm(sA_lo, &tmA_lo, lbar, k0, m0, b);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; j++) {
                            tma_load_3d_2sm(sA_hi + j * (BLOCK_K * 128), &tmA_hi, lbar, m0 + 64 * j, k0, b);
                            if (p.passes == 3) tma_load_3d_2sm(sA_lo + j * (BLOCK_K * 128), &tmA_lo, lbar, m0 + 64 * j, k0, b);
                        }
                    }
                    if (!B_MN) {
                        tma_load_3d_2sm(sB_hi, &tmB_hi, lbar, k0, n0, b);
                        if (p.passes == 3) tma_load_3d_2sm(sB_lo, &tmB_lo, lbar, k0, n0, b);
                    } else {
#pragma unroll
                        for (int j = 0; j < HALF_N / 64; j++) {
                            tma_load_3d_2sm(sB_hi + j * (BLOCK_K * 128), &tmB_hi, lbar, n0 + 64 * j, k0, b);
                            if (p.passes == 3) tma_load_3d_2sm(sB_lo + j * (BLOCK_K * 128), &tmB_lo, lbar, n0 + 64 * j, k0, b);
                        }
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: one thread of the leader CTA =================
        if (leader && lane == 0) {
            constexpr uint32_t K_ROW_BYTES = BLOCK_K * 2;
            constexpr uint32_t K_LAYOUT = (K_ROW_BYTES == 128) ? 2u : 4u;
            constexpr uint32_t A_LBO = A_MN ? BLOCK_K * 128 : 16, A_SBO = A_MN ? 1024 : 8 * K_ROW_BYTES, A_KSTEP = A_MN ? 2048 : 32;
            constexpr uint32_t B_LBO = B_MN ? BLOCK_K * 128 : 16, B_SBO = B_MN ? 1024 : 8 * K_ROW_BYTES, B_KSTEP = B_MN ? 2048 : 32;
            constexpr uint32_t A_LAYOUT = A_MN ? 2u : K_LAYOUT, B_LAYOUT = B_MN ? 2u : K_LAYOUT;
            int stage = 0;
            uint32_t phase = 0;
            int local_tile = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters, local_tile++) {
                const int buf = local_tile & 1;
                const uint32_t buf_phase = (uint32_t)((local_tile >> 1) & 1);
                const int n_tile0 = ((t % tiles_per_batch) % p.n_tiles) * TILE_N;
                const int n_eff = min(TILE_N, ((p.N - n_tile0 + 31) / 32) * 32);
                const uint32_t idesc = make_idesc(TILE_M, n_eff, A_MN, B_MN);
                mbar_wait(tempty_bar(buf), buf_phase ^ 1u);  // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TILE_N);
                uint32_t accumulate = 0;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);  // both CTAs' tiles have landed
                    tc_fence_after();
                    const uint32_t sA_hi = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sA_lo = sA_hi + A_TILE_BYTES;
                    const uint32_t sB_hi = sA_lo + A_TILE_BYTES;
                    const uint32_t sB_lo = sB_hi + B_TILE_BYTES;
                    const int k_valid = min(BLOCK_K, p.K - kb * BLOCK_K);
                    const int ksteps = (k_valid + UMMA_K - 1) / UMMA_K;
                    for (int prod = 0; prod < ((p.debug_flags & 4) ? 0 : p.passes); prod++) {
                        const uint32_t sa = (prod == 2) ? sA_lo : sA_hi;
                        const uint32_t sb = (prod == 1) ? sB_lo : sB_hi;
                        for (int ks = 0; ks < ksteps; ks++) {
                            uint64_t adesc = make_smem_desc(sa + ks * A_KSTEP, A_LBO, A_SBO, A_LAYOUT);
                            uint64_t bdesc = make_smem_desc(sb + ks * B_KSTEP, B_LBO, B_SBO, B_LAYOUT);
                            umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    umma_commit_2sm(empty_bar(stage));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_2sm(tfull_bar(buf));
            }
        }
    } else {
        // ================= epilogue warps 2..5 (both CTAs): this CTA's 128 rows =================
        const int q = warp & 3;
        int local_tile = 0;
        uint32_t epi_chunk = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters, local_tile++) {
            const int buf = local_tile & 1;
            const uint32_t buf_phase = (uint32_t)((local_tile >> 1) & 1);
            const int b = t / tiles_per_batch;
            const int rem = t - b * tiles_per_batch;
            const int m0 = (rem / p.n_tiles) * TILE_M + (int)rank * BLOCK_M;
            const int n0 = (rem % p.n_tiles) * TILE_N;
            mbar_wait(tfull_bar(buf), buf_phase);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            float* drow = p.D + (int64_t)b * p.sDb + (int64_t)row * p.ldd;
            const bool row_ok = row < p.M && !(p.debug_flags & 1);
            const bool vec_ok = ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15u) == 0) && ((p.sDb & 3) == 0);
#pragma unroll 1
            for (int c = 0; c < TILE_N / 32; c++) {
                const int col0 = n0 + c * 32;
                if (col0 >= p.N) break;
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TILE_N + c * 32), r);
                tmem_ld_wait();
                if (p.tma_store) {
                    if (!(p.debug_flags & 1))
                        stage_and_store(r, smem_base + L::EPI_OFFSET + (uint32_t)((warp - 2) * 2 + (epi_chunk & 1)) * kStageTileBytes, lane, &tmD, col0,
                                        m0 + q * 32, b);
                    epi_chunk++;
                } else if (row_ok) {
                    if (vec_ok && col0 + 32 <= p.N) {
                        float4* dst = reinterpret_cast<float4*>(drow + col0);
#pragma unroll
                        for (int v = 0; v < 8; v++)
                            dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                                 __uint_as_float(r[4 * v + 3]));
                    } else {
#pragma unroll
                        for (int v = 0; v < 32; v++)
                            if (col0 + v < p.N) drow[col0 + v] = __uint_as_float(r[v]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty_bar(buf), 0);  // the leader's barrier (local for rank 0)
        }
        if (p.tma_store && elect_one()) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA may free TMEM / exit while the pair can still touch its shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    }
}

template <int BLOCK_K, int STAGES, bool A_MN, bool B_MN>
mb_status launch_variant2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo, const CUtensorMap& d_map,
                          const TcParams& p, cudaStream_t st) {
    using L = SmemLayout2<BLOCK_K, STAGES>;
    auto kern = gemm_tc2_kernel<BLOCK_K, STAGES, A_MN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        MB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    int tiles = p.m_tiles * p.n_tiles * p.batches;
    int clusters = sm_count() / 2;
    if (tiles < clusters) clusters = tiles;
    kern<<<2 * clusters, kTcThreads, L::TOTAL, st>>>(a_hi, a_lo, b_hi, b_lo, d_map, p);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace

bool gemm_tc_supported(int64_t a_inner, int64_t b_inner) {
    // TMA: 16-byte global strides => inner extents (bf16) multiples of 8
    return (a_inner % 8 == 0) && (b_inner % 8 == 0) && get_encode_fn() != nullptr;
}

// D[b] (M x N, fp32, ld = ldd) = A[b] . B[b]^T-like contraction over K with bf16 hi/lo operands.
//   a_mn == false : A stored [batch][M][K] (K contiguous, row stride lda)   a_mn == true : stored [batch][K][M] (M contiguous)
//   b_mn == false : B stored [batch][N][K]                                  b_mn == true : stored [batch][K][N]
mb_status gemm_tc(const void* A_hi, const void* A_lo, int64_t lda, int64_t sAb, bool a_mn, const void* B_hi, const void* B_lo, int64_t ldb, int64_t sBb,
                  bool b_mn, float* D, int64_t ldd, int64_t sDb, int M, int N, int K, int batches, int passes, int block_n, cudaStream_t st) {
    if (M == 0 || N == 0 || batches == 0) return MB_OK;
    if (K == 0) {
        set_error("gemm_tc: K == 0");
        return MB_ERR_INVALID;
    }
    if (!A_lo || !B_lo) passes = 1;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    // tile configurations: block_n 256 -> (BLOCK_N 256, BLOCK_K 32, 4 stages), 2560 -> (256, 64, 2), 128 -> (128, 64, 3),
    // 512 -> 2-CTA pair, 256x256 cluster tile, BLOCK_K 64, 3 stages ; 5120 -> 2-CTA, BLOCK_K 32, 6 stages
    const bool two_cta = (block_n == 512 || block_n == 5120);
    const int bk = (block_n == 256 || block_n == 5120) ? 32 : 64;
    const int bn_real = (block_n == 2560 || two_cta) ? 256 : block_n;
    const uint32_t bn = two_cta ? 128u : (uint32_t)bn_real;  // TMA box rows of a K-major B tile (a CTA's half in 2-CTA mode)
    if (!a_mn) {
        MB_TRY(make_map(&ma_hi, A_hi, K, M, batches, lda, sAb, BLOCK_M, bk));
        if (passes == 3) MB_TRY(make_map(&ma_lo, A_lo, K, M, batches, lda, sAb, BLOCK_M, bk));
    } else {
        MB_TRY(make_map(&ma_hi, A_hi, M, K, batches, lda, sAb, bk));
        if (passes == 3) MB_TRY(make_map(&ma_lo, A_lo, M, K, batches, lda, sAb, bk));
    }
    if (!b_mn) {
        MB_TRY(make_map(&mb_hi, B_hi, K, N, batches, ldb, sBb, bn, bk));
        if (passes == 3) MB_TRY(make_map(&mb_lo, B_lo, K, N, batches, ldb, sBb, bn, bk));
    } else {
        MB_TRY(make_map(&mb_hi, B_hi, N, K, batches, ldb, sBb, bk));
        if (passes == 3) MB_TRY(make_map(&mb_lo, B_lo, N, K, batches, ldb, sBb, bk));
    }
    if (passes != 3) {
        ma_lo = ma_hi;
        mb_lo = mb_hi;
    }
    // fp32 output map for the TMA-store epilogue: dims {N, M, batches}, box {32 columns, 32 rows, 1}, SWIZZLE_128B
    CUtensorMap md;
    std::memset(&md, 0, sizeof(md));
    static int want_tma_store = [] { const char* e = getenv("MB_TC_TMA_STORE"); return e ? atoi(e) : 1; }();
    bool tma_store = want_tma_store && ((reinterpret_cast<uintptr_t>(D) & 15u) == 0) && (ldd % 4 == 0) && (batches == 1 || sDb % 4 == 0);
    if (tma_store) {
        EncodeTiledFn fn = get_encode_fn();
        cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)batches};
        cuuint64_t strides[2] = {(cuuint64_t)ldd * 4, (cuuint64_t)(batches == 1 ? (int64_t)M * ldd : sDb) * 4};
        cuuint32_t box[3] = {32, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = fn(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tma_store = false;
    }
    TcParams p;
    p.tma_store = tma_store ? 1 : 0;
    p.D = D;
    p.ldd = ldd;
    p.sDb = sDb;
    p.M = M;
    p.N = N;
    p.K = K;
    p.batches = batches;
    p.passes = passes;
    {
        static int dbg = [] { const char* e = getenv("MB_TC_DEBUG"); return e ? atoi(e) : 0; }();
        p.debug_flags = dbg;
    }
    p.m_tiles = two_cta ? (M + 255) / 256 : (M + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = (N + bn_real - 1) / bn_real;
    if (two_cta) {
#define MB_TC2_DISPATCH(BK, ST)                                                                           \
    if (!a_mn && !b_mn) return launch_variant2<BK, ST, false, false>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);  \
    if (!a_mn && b_mn) return launch_variant2<BK, ST, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);    \
    if (a_mn && b_mn) return launch_variant2<BK, ST, true, true>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);      \
    return launch_variant2<BK, ST, true, false>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);
        if (block_n == 512) {
            MB_TC2_DISPATCH(64, 3)
        } else {
            MB_TC2_DISPATCH(32, 6)
        }
#undef MB_TC2_DISPATCH
    }
#define MB_TC_DISPATCH(BN, BK, ST)                                                                          \
    if (!a_mn && !b_mn) return launch_variant<BN, BK, ST, false, false>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st); \
    if (!a_mn && b_mn) return launch_variant<BN, BK, ST, false, true>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);   \
    if (a_mn && b_mn) return launch_variant<BN, BK, ST, true, true>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);     \
    return launch_variant<BN, BK, ST, true, false>(ma_hi, ma_lo, mb_hi, mb_lo, md, p, st);
    if (block_n == 256) {
        MB_TC_DISPATCH(256, 32, 4)
    } else if (block_n == 2560) {
        MB_TC_DISPATCH(256, 64, 2)
    } else if (block_n == 128) {
        MB_TC_DISPATCH(128, 64, 3)
    }
#undef MB_TC_DISPATCH
    set_error("gemm_tc: unsupported block_n");
    return MB_ERR_UNSUPPORTED;
}

}  // namespace mb
