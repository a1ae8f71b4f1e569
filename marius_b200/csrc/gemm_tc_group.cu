// gemm_tc_group.cu -- grouped, table-scheduled 2-CTA tcgen05 GEMMs (the tensor-core contraction kernels of the library):
// gemm_tc_group_kernel<false> (operands from bf16 hi/lo arrays; forward scores), gemm_tc_group_kernel<true> and gemm_tc_ts_kernel (backward:
// the A operand is produced inside the kernel from the fp32 scores, staged in shared memory / in tensor memory).
//
// One persistent CTA pair per two SMs (cta_group::2), warp-specialised (TMA producer / MMA issuer / 4 epilogue warps), with two
// properties that matter at the named shape (ComplEx d=400, 1000 negatives: each backward contraction has few cluster tiles per
// CTA pair, 2.16 waves at batch 10k):
//   (1) up to two independent problems (dA = G.Neg and dNeg = G^T.A) share ONE persistent launch: operand majors, shapes and
//       tensor maps are per-problem run-time data, so the tail of one contraction is filled with tiles of the other;
//   (2) tiles are handed out from a host-built table: tiles sorted by cost (ragged N tiles are cheaper), assigned to CTA pairs
//       in snake order (longest-processing-time-first), so every pair gets the same work within one small tile.
// All three warp roles read the same table entries, so no in-kernel tile broadcast is needed.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "gemm_tc_ptx.cuh"
#include "kernels.h"

namespace mb {

namespace {

using namespace tcptx;

constexpr int GTILE_M = 256, GTILE_N = 256, GHALF_N = 128;

struct GProblem {
    float* D;
    int64_t ldd, sDb;
    int M, N, K, batches;
    int a_mn, b_mn, tma_store;
    // forward: per-(row, 64-column slot) softmax statistics of the score tile (null: none)
    float2* stats;
    int stat_slots;
    // backward (CONV kernel): the A operand is produced in shared memory from the fp32 matrix `conv_src` [batches][conv_rows][conv_cols]
    const float* conv_src;
    const float* conv_z;  // [batches][conv_rows] per-row shift (conv_mode 1) or null
    int64_t conv_ld, conv_sb;
    int conv_rows, conv_cols, conv_mode;
};
struct GParams {
    GProblem prob[2];
    const int4* table;  // [rounds][num_clusters] : (problem | width << 8, or -1; batch; m0; n0), width = tile columns (multiple of 32 off the N edge)
    int rounds;
    int passes;
    int debug_flags;
};
struct GMaps {
    CUtensorMap m[2][5];  // per problem: A_hi, A_lo, B_hi, B_lo, D
};

constexpr int kConvGroups = 3;                 // converter groups, each takes every kConvGroups-th k-block
constexpr int kConvWarps = 4 * kConvGroups;    // one warp of every group per SM sub-partition
constexpr int kConvThreads = kTcThreads + 32 * kConvWarps;  // 576: producer, MMA, 4 epilogue warps, 12 converter warps
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 4 fp32 -> 4 bf16 hi + 4 bf16 lo (x ~= hi + lo), one 8-byte shared-memory store each
__device__ __forceinline__ void split4_store(uint32_t dst_hi, uint32_t dst_lo, const float (&g)[4], bool want_lo) {
    uint32_t h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(g[2 * i + 1]), "f"(g[2 * i]));  // first operand -> upper half
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(g[2 * i + 1] - h1), "f"(g[2 * i] - h0));
    }
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst_hi), "r"(h[0]), "r"(h[1]) : "memory");
    if (want_lo) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst_lo), "r"(l[0]), "r"(l[1]) : "memory");
}
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {  // read once per tile: keep it out of the (tiny) L1
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void named_barrier_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// Shared-memory layout and pipeline geometry of the two instantiations.
//   CONV = false (forward scores, diagnostics): k-blocks of 64, 3 operand stages of 64 KB (A hi/lo + B hi/lo), 192 threads.
//   CONV = true  (backward contractions): the A operand of every problem is produced in the kernel from an fp32 matrix (the scores S):
//                G = exp(S - z_row) (the SoftmaxCrossEntropy gradient, loss.cpp:50-67) or G = S.  k-blocks of 32: 4 operand stages of
//                32 KB + 3 raw fp32 score tiles of 16 KB (one per converter group), so that the score tiles are in flight (TMA) ahead
//                of their conversion, independently of the operand stages; 576 threads (12 converter warps).  The gradient matrix never
//                exists in global memory.  Used when the output is wider than the tensor-memory-A kernel below can hold (N > 416).
template <bool CONV>
struct Geo {
    static constexpr int BK = CONV ? 32 : 64;                 // k-block (bf16 elements)
    static constexpr int STAGES = CONV ? 4 : 3;
    static constexpr int A_TILE = BLOCK_M * BK * 2;           // one of hi / lo: this CTA's 128 rows
    static constexpr int B_TILE = GHALF_N * BK * 2;           // one of hi / lo: this CTA's half of the tile columns
    static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    static constexpr int RAW_SLOTS = CONV ? kConvGroups : 0;  // one private slot per converter group
    static constexpr int RAW_TILE = BLOCK_M * BK * 4;         // fp32 score tile of one k-block (this CTA's part)
    static constexpr int RAW_OFFSET = STAGES * STAGE;
    static constexpr int EPI_OFFSET = RAW_OFFSET + RAW_SLOTS * RAW_TILE;
    static constexpr int BAR_OFFSET = EPI_OFFSET + kEpilogueSmemBytes;
    static constexpr int SMEM_TOTAL = BAR_OFFSET + 256 + 1024;
    static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
};

template <bool CONV>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV ? kConvThreads : kTcThreads, 1) gemm_tc_group_kernel(const __grid_constant__ GMaps maps, const GParams p) {
    using G = Geo<CONV>;
    constexpr int BK = G::BK, NST = G::STAGES;
    constexpr int TMEM_COLS = 2 * GTILE_N;
    constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + G::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * NST + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * NST + 2 + b); };
    const uint32_t tmem_holder = bar_base + 8u * (2 * NST + 4);
    auto raw_bar = [&](int s) { return bar_base + 8u * (2 * NST + 6 + s); };  // CONV: this CTA's fp32 score tile in raw slot s has landed
    auto afull_bar = [&](int s) { return bar_base + 8u * (2 * NST + 6 + G::RAW_SLOTS + s); };  // CONV, non-leader CTA: its A tiles of stage s are written
    volatile uint32_t* tmem_holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int lo_mult = p.passes == 3 ? 2 : 1;
    const uint32_t stage_tx_pair = (uint32_t)(2 * lo_mult * ((CONV ? 0 : G::A_TILE) + G::B_TILE));

    if (warp == 0 && lane == 0) {
        for (int q = 0; q < 2; q++)
            for (int j = 0; j < 5; j++) prefetch_tmap(&maps.m[q][j]);
        for (int s = 0; s < NST; s++) {
            mbar_init(full_bar(s), CONV ? 1 + 2 : 1);  // + one arrival per CTA once its converter warps have written the A tiles
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < G::RAW_SLOTS; s++) mbar_init(raw_bar(s), 1);
        if (CONV)
            for (int s = 0; s < NST; s++) mbar_init(afull_bar(s), 1);
        for (int b = 0; b < 2; b++) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder_ptr;

    if (warp == 0) {
        // ================= TMA producer (both CTAs): the whole warp runs the loop, one elected lane issues =================
        int stage = 0;
        uint32_t phase = 0;
        int4 e_next = p.table[cluster_id];  // the next table entry is fetched one tile ahead: its latency never sits between two tiles
        for (int r = 0; r < p.rounds; r++) {
            const int4 e = e_next;
            if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
            if (e.x < 0) continue;
            const int pi = e.x & 0xff;
            const GProblem& pr = p.prob[pi];
            const CUtensorMap* mA_hi = &maps.m[pi][0];
            const CUtensorMap* mA_lo = &maps.m[pi][1];
            const CUtensorMap* mB_hi = &maps.m[pi][2];
            const CUtensorMap* mB_lo = &maps.m[pi][3];
            const int b = e.y;
            const int m0 = e.z + (int)rank * BLOCK_M;
            const int n_eff = ((min(e.x >> 8, pr.N - e.w) + 31) / 32) * 32;
            const int n0 = e.w + (int)rank * (n_eff / 2);
            const int num_k_blocks = (pr.K + BK - 1) / BK;
            const bool a_mn = pr.a_mn != 0, b_mn = pr.b_mn != 0, three = p.passes == 3;
            for (int kb = 0; kb < num_k_blocks; kb++) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sA_hi = smem_base + stage * G::STAGE;
                const uint32_t sA_lo = sA_hi + G::A_TILE;
                const uint32_t sB_hi = sA_lo + G::A_TILE;
                const uint32_t sB_lo = sB_hi + G::B_TILE;
                const uint32_t lbar = full_bar(stage) & kPeerMask;
                const int k0 = kb * BK;
                if (elect_one()) {
                    if (p.debug_flags & 2) {  // ablation: no TMA loads, the MMAs run on whatever is in shared memory
                        if (leader) mbar_arrive(full_bar(stage));
                    } else {
                        if (leader) mbar_expect_tx(full_bar(stage), stage_tx_pair);
                        if (CONV) {
                            // the A tiles of the stage are written by the converter warps
                        } else if (!a_mn) {
                            tma_load_3d_2sm(sA_hi, mA_hi, lbar, k0, m0, b);
                            if (three) tma_load_3d_2sm(sA_lo, mA_lo, lbar, k0, m0, b);
                        } else {
#pragma unroll
                            for (int j = 0; j < BLOCK_M / 64; j++) {
                                tma_load_3d_2sm(sA_hi + j * (BK * 128), mA_hi, lbar, m0 + 64 * j, k0, b);
                                if (three) tma_load_3d_2sm(sA_lo + j * (BK * 128), mA_lo, lbar, m0 + 64 * j, k0, b);
                            }
                        }
                        if (!b_mn) {
                            tma_load_3d_2sm(sB_hi, mB_hi, lbar, k0, n0, b);
                            if (three) tma_load_3d_2sm(sB_lo, mB_lo, lbar, k0, n0, b);
                        } else {
#pragma unroll
                            for (int j = 0; j < GHALF_N / 64; j++) {
                                tma_load_3d_2sm(sB_hi + j * (BK * 128), mB_hi, lbar, n0 + 64 * j, k0, b);
                                if (three) tma_load_3d_2sm(sB_lo + j * (BK * 128), mB_lo, lbar, n0 + 64 * j, k0, b);
                            }
                        }
                    }
                }
                __syncwarp();
                if (++stage == NST) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: warp 1 of the leader CTA, warp-uniform loop, one elected lane issues =================
        if (leader) {
            int stage = 0;
            uint32_t phase = 0;
            int local_tile = 0;
            // K-major tiles: rows of BK * 2 bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B), 8-row groups SBO apart, LBO unused, k-step +32 B.
            // MN-major tiles (SWIZZLE_128B): 64-element MN slabs BK * 128 B apart (LBO), 8-k-row groups 1024 B apart (SBO), k-step +2048 B.
            constexpr uint32_t kLayoutK = (BK * 2 == 128) ? 2u : 4u;
            constexpr uint32_t kDescHiK = ((uint32_t)((8 * BK * 2) >> 4) & 0x3fffu) | (1u << 14) | (kLayoutK << 29);
            constexpr uint32_t kDescHiMN = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
            const uint32_t npass = (p.debug_flags & 4) ? 0u : (uint32_t)p.passes;  // bit 2: ablation, no MMAs
            int4 e_next = p.table[cluster_id];
            for (int r = 0; r < p.rounds; r++) {
                const int4 e = e_next;
                if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
                if (e.x < 0) continue;
                const GProblem& pr = p.prob[e.x & 0xff];
                const bool a_mn = pr.a_mn != 0, b_mn = pr.b_mn != 0;
                const uint32_t a_lbo = (a_mn ? (uint32_t)(BK * 128) >> 4 : 1u) << 16, a_kstep = a_mn ? 2048u >> 4 : 32u >> 4;
                const uint32_t b_lbo = (b_mn ? (uint32_t)(BK * 128) >> 4 : 1u) << 16, b_kstep = b_mn ? 2048u >> 4 : 32u >> 4;
                const uint32_t a_hi32 = a_mn ? kDescHiMN : kDescHiK, b_hi32 = b_mn ? kDescHiMN : kDescHiK;
                const int buf = local_tile & 1;
                const uint32_t buf_phase = (uint32_t)((local_tile >> 1) & 1);
                local_tile++;
                const int n_eff = ((min(e.x >> 8, pr.N - e.w) + 31) / 32) * 32;
                const uint32_t idesc = make_idesc(GTILE_M, n_eff, a_mn, b_mn);
                const int num_k_blocks = (pr.K + BK - 1) / BK;
                mbar_wait(tempty_bar(buf), buf_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * GTILE_N);
                uint32_t accumulate = 0;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sA_hi = (smem_base + stage * G::STAGE) >> 4;  // 16-byte units, < 2^14
                    const uint32_t sA_lo = sA_hi + (G::A_TILE >> 4);
                    const uint32_t sB_hi = sA_lo + (G::A_TILE >> 4);
                    const uint32_t sB_lo = sB_hi + (G::B_TILE >> 4);
                    const int k_valid = min(BK, pr.K - kb * BK);
                    const int ksteps = (k_valid + UMMA_K - 1) / UMMA_K;
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t prod = 0; prod < 3; prod++) {  // hi.hi, hi.lo, lo.hi
                            if (prod < npass) {
                                const uint32_t sa = a_lbo | ((prod == 2) ? sA_lo : sA_hi);
                                const uint32_t sb = b_lbo | ((prod == 1) ? sB_lo : sB_hi);
#pragma unroll
                                for (int ks = 0; ks < BK / UMMA_K; ks++) {
                                    if (ks < ksteps) {
                                        const uint64_t adesc = ((uint64_t)a_hi32 << 32) | (uint64_t)(sa + ks * a_kstep);
                                        const uint64_t bdesc = ((uint64_t)b_hi32 << 32) | (uint64_t)(sb + ks * b_kstep);
                                        umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, accumulate);
                                        accumulate = 1;
                                    }
                                }
                            }
                        }
                        umma_commit_2sm(empty_bar(stage));
                        if (kb == num_k_blocks - 1) umma_commit_2sm(tfull_bar(buf));
                    }
                    __syncwarp();
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        } else if (CONV) {
            // Non-leader CTA, CONV: this otherwise idle warp forwards "this CTA's A tiles of the stage are written" to the leader's full
            // barrier.  The converter warps only pay a CTA-local arrive; the cross-CTA hop happens here, off their critical path, and
            // without a GPU-scope memory barrier (measured ~0.9 us per arrive with .release.cluster: the A tiles are settled in this CTA's
            // shared memory -- every writer fenced and synchronised -- before this warp is told).
            int stage = 0;
            uint32_t phase = 0;
            int4 e_next = p.table[cluster_id];
            for (int r = 0; r < p.rounds; r++) {
                const int4 e = e_next;
                if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
                if (e.x < 0) continue;
                const int num_k_blocks = (p.prob[e.x & 0xff].K + BK - 1) / BK;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(afull_bar(stage), phase);
                    if (elect_one()) mbar_arrive_remote_light(full_bar(stage), 0);
                    __syncwarp();
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (!CONV || warp < 6) {
        // ================= epilogue warps 2..5 (both CTAs) =================
        const int q = warp & 3;
        int local_tile = 0;
        uint32_t epi_chunk = 0;
        int4 e_next = p.table[cluster_id];
        for (int r = 0; r < p.rounds; r++) {
            const int4 e = e_next;
            if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
            if (e.x < 0) continue;
            const GProblem& pr = p.prob[e.x & 0xff];
            const CUtensorMap* mD = &maps.m[e.x & 0xff][4];
            const int buf = local_tile & 1;
            const uint32_t buf_phase = (uint32_t)((local_tile >> 1) & 1);
            local_tile++;
            const int b = e.y;
            const int m0 = e.z + (int)rank * BLOCK_M;
            const int n0 = e.w;
            mbar_wait(tfull_bar(buf), buf_phase);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            float* drow = pr.D + (int64_t)b * pr.sDb + (int64_t)row * pr.ldd;
            const bool row_ok = row < pr.M && !(p.debug_flags & 1);
            // 32-column chunks of this warp's 32 accumulator rows; the TMEM load of chunk c+1 is in flight while chunk c is staged / stored
            const int n_lim = min(pr.N, n0 + (e.x >> 8));
            const int n_chunks = (p.debug_flags & 8) ? 0 : (n_lim - n0 + 31) / 32;  // bit 3: ablation, no epilogue at all
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * GTILE_N);
            uint32_t ra[32], rb[32];
            // forward contraction: running (max, sum exp) of this thread's score row over the tile's columns -- the SoftmaxCrossEntropy
            // statistics (loss.cpp:57-66), merged over the row's column tiles by loss_merge_kernel
            float m_run = -INFINITY, l_run = 0.f;
            const bool want_stats = !CONV && pr.stats != nullptr;
            auto emit = [&](const uint32_t (&rg)[32], int c) {
                const int col0 = n0 + c * 32;
                if (want_stats) {
                    float cm = -INFINITY;
#pragma unroll
                    for (int v = 0; v < 32; v++)
                        if (col0 + v < pr.N) cm = fmaxf(cm, __uint_as_float(rg[v]));
                    const float mn = fmaxf(m_run, cm);
                    const float mnl = mn * kLog2e;
                    float acc = 0.f;
#pragma unroll
                    for (int v = 0; v < 32; v++)
                        if (col0 + v < pr.N) acc += ex2_approx(fmaf(__uint_as_float(rg[v]), kLog2e, -mnl));
                    l_run = fmaf(l_run, ex2_approx((m_run - mn) * kLog2e), acc);
                    m_run = mn;
                }
                if (pr.tma_store) {
                    if (!(p.debug_flags & 1))
                        stage_and_store(rg, smem_base + G::EPI_OFFSET + (uint32_t)((warp - 2) * 2 + (epi_chunk & 1)) * kStageTileBytes, lane, mD, col0,
                                        m0 + q * 32, b);
                    epi_chunk++;
                } else if (row_ok) {
#pragma unroll
                    for (int v = 0; v < 32; v++)
                        if (col0 + v < pr.N) drow[col0 + v] = __uint_as_float(rg[v]);
                }
            };
            if (n_chunks > 0) tmem_ld_32x32b_x32(taddr, ra);
#pragma unroll 1
            for (int c = 0; c < n_chunks; c += 2) {
                tmem_ld_wait();
                if (c + 1 < n_chunks) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), rb);
                emit(ra, c);
                if (c + 1 < n_chunks) {
                    tmem_ld_wait();
                    if (c + 2 < n_chunks) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 2) * 32), ra);
                    emit(rb, c + 1);
                }
            }
            if (want_stats && row < pr.M && n_chunks > 0)
                pr.stats[((int64_t)b * pr.M + row) * pr.stat_slots + (n0 >> 6)] = make_float2(m_run, l_run);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty_bar(buf), 0);
        }
        if (elect_one()) bulk_wait_all();
    } else {
        // ================= converter warps 6..13 (both CTAs, CONV only): fp32 scores -> G -> bf16 hi/lo operand tiles =================
        // kConvGroups groups of four warps work on INTERLEAVED k-blocks (group g takes the k-blocks whose running number is g mod
        // kConvGroups), so the serial chain of one k-block (wait for the raw tile, load, synchronise, convert, fence, synchronise, arrive)
        // may take kConvGroups k-block periods of the tensor core: the groups pipeline against each other instead of splitting every
        // k-block twelve ways.  Measured (B = 50 000, both backward contractions): 2 groups 545 us, 3 groups 475 us, 4 groups (80
        // registers per thread at 704 threads: spills) 516 us; 3 groups reading the scores straight from global memory into registers
        // one own k-block ahead (no raw ring, 6 operand stages) 673 us -- L2 latency under this load exceeds the slack; the contraction
        // alone, operands from global bf16 arrays, 369 us.  At 3 groups the kernel sits at ~75 % of the shared-memory bandwidth (operand
        // reads 34 %, converter loads / stores 42 %) and 65 % tensor-pipe utilisation (profiles/r2_ncu_step_mid.csv); with private raw slots
        // 460 us.  gemm_tc_ts_kernel below (A through tensor memory, full-width tiles) does the same work in 374 us and is the one used
        // whenever the output fits its 416 accumulator columns.
        // Raw ring: this CTA's fp32 score tile of a k-block (128 operand rows x 32 K, 16 KB, row-major, no swizzle) is loaded by TMA into
        // the private slot of the group that converts it; thread 0 of the group issues the load of the group's next k-block as soon as
        // every thread of the group has read the current one (the load cursor walks the same tile table, across tile boundaries).
        // Operand stage: the A region gets what the TMA loads of the non-CONV kernel would produce for BK = 32:
        //   K-major A (dA = G . Neg):    128 rows (M row m0 + i) of 64 bytes (32 K), SWIZZLE_64B: 16-byte chunk index XOR ((row >> 1) & 3)
        //   MN-major A (dNeg = G^T . A): two slabs (64 M columns each) of 32 rows (K row k0 + r) of 128 bytes, SWIZZLE_128B
        // hi tile then lo tile.  Thread mapping: consecutive threads take consecutive float4 of a raw row (conflict-free 128-bit loads);
        // their 8-byte stores cover whole operand rows (all 32 banks).
        constexpr int kGroupThreads = 128;
        constexpr int kPieces = (BLOCK_M * BK / 4) / kGroupThreads;           // float4 pieces per thread per k-block (8)
        const int ct = (int)threadIdx.x - kTcThreads;
        const int grp = ct / kGroupThreads, gt = ct % kGroupThreads;
        const int bar_read = 1 + 2 * grp, bar_done = 2 + 2 * grp;             // named barriers of the group
        const uint32_t raw_base = smem_base + G::RAW_OFFSET;
        // ---- load cursor (thread 0 of each group): the group's own k-blocks, one ahead (see gemm_tc_ts_kernel on why the slot is private)
        int lr = -1, lkb = 0, lnkb = 0;  // table round, k-block, k-blocks of that tile
        int4 le = make_int4(-1, 0, 0, 0);
        auto next_tile = [&]() {
            while (++lr < p.rounds) {
                le = p.table[lr * num_clusters + cluster_id];
                if (le.x >= 0) {
                    lnkb = (p.prob[le.x & 0xff].K + BK - 1) / BK;
                    lkb = 0;
                    return;
                }
            }
        };
        auto load_seek = [&](int steps) {  // the cursor moves `steps` k-blocks forward, across tiles; lr >= p.rounds: past the end
            while (lr < p.rounds) {
                if (lkb + steps < lnkb) {
                    lkb += steps;
                    return;
                }
                steps -= lnkb - lkb;
                next_tile();
            }
        };
        auto load_issue = [&]() {  // TMA of the score tile at the load cursor into the group's raw slot
            if (lr >= p.rounds) return;
            const int pi = le.x & 0xff;
            const GProblem& lp = p.prob[pi];
            const int lm = le.z + (int)rank * BLOCK_M;
            mbar_expect_tx(raw_bar(grp), G::RAW_TILE);
            if (lp.a_mn == 0)
                tma_load_3d(raw_base + grp * G::RAW_TILE, &maps.m[pi][0], raw_bar(grp), lkb * BK, lm, le.y);  // box {32 K columns, 128 M rows}
            else
                tma_load_3d(raw_base + grp * G::RAW_TILE, &maps.m[pi][0], raw_bar(grp), lm, lkb * BK, le.y);  // box {128 M columns, 32 K rows}
        };
        if (gt == 0) {
            next_tile();
            load_seek(grp);
            load_issue();
        }
        int stage = 0;
        uint32_t phase = 0, raw_phase = 0;
        int n = 0;  // running k-block number
        int4 e_next = p.table[cluster_id];
        for (int r = 0; r < p.rounds; r++) {
            const int4 e = e_next;
            if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
            if (e.x < 0) continue;
            const GProblem& pr = p.prob[e.x & 0xff];
            const int b = e.y;
            const int m_base = e.z + (int)rank * BLOCK_M;
            const int num_k_blocks = (pr.K + BK - 1) / BK;
            const bool a_mn = pr.a_mn != 0, want_lo = p.passes == 3, expo = pr.conv_mode == 1;
            const int rows_lim = pr.conv_rows, cols_lim = pr.conv_cols;
            const int lim_m = a_mn ? cols_lim : rows_lim, lim_k = a_mn ? rows_lim : cols_lim;  // extents along the tile's M / K directions
            const bool m_interior = m_base + BLOCK_M <= lim_m;
            const float* zb = expo ? pr.conv_z + (int64_t)b * rows_lim : nullptr;
            // piece j of this thread: float4 index f = gt + 128 j of the raw tile
            //   K-major  raw [128 rows][8 float4]:   row i = f >> 3, c = f & 7      -> operand row i, K elements 4c .. 4c+3
            //   MN-major raw [32 K rows][32 float4]: K row kr = f >> 5, c = f & 31   -> slab c >> 4, operand row kr, M elements 4 (c & 15) ..
            // zfix[j]: what is known about the piece's shift for the whole tile.  K-major: the shift itself (the piece's score row does not
            // change with the k-block), +inf if the row lies outside the matrix.  MN-major: +inf if the piece's columns lie outside, else
            // -inf (the shift of its row is loaded per k-block and combined with fmax).  A +inf shift turns the TMA's zero fill into
            // exp2(0 - inf) = 0.
            uint32_t dst_off[kPieces];
            float zfix[kPieces];
            const int kr0 = gt >> 5;  // MN-major: the K row of piece j is kr0 + 4 j
#pragma unroll
            for (int j = 0; j < kPieces; j++) {
                const int f = gt + kGroupThreads * j;
                if (!a_mn) {
                    const int i = f >> 3, c = f & 7;
                    dst_off[j] = (uint32_t)i * 64u + (uint32_t)(((c >> 1) ^ ((i >> 1) & 3)) << 4) + (uint32_t)(c & 1) * 8u;
                    zfix[j] = (expo && m_base + i < rows_lim) ? __ldg(zb + m_base + i) : INFINITY;
                } else {
                    const int kr = f >> 5, c = f & 31;
                    dst_off[j] = (uint32_t)(c >> 4) * (uint32_t)(BK * 128) + (uint32_t)kr * 128u + (uint32_t)((((c & 15) >> 1) ^ (kr & 7)) << 4) +
                                 (uint32_t)(c & 1) * 8u;
                    zfix[j] = (m_base + 4 * c < cols_lim) ? -INFINITY : INFINITY;
                }
            }
            const int kcol = 4 * (gt & 7);  // K-major: first K column of this thread's pieces within the k-block
#pragma unroll 1
            for (int kb = 0; kb < num_k_blocks; kb++, n++) {
                if (n % kConvGroups == grp) {
                    const int k0 = kb * BK;
                    const bool k_interior = k0 + BK <= lim_k;  // (warp-uniform)
                    float z[kPieces];
                    if (expo) {
                        if (!a_mn) {
                            const bool col_ok = k_interior || k0 + kcol < cols_lim;
#pragma unroll
                            for (int j = 0; j < kPieces; j++) z[j] = col_ok ? zfix[j] : INFINITY;
                        } else {
                            // issued before the wait for the raw tile; a batch's shifts (4 KB) stay in L1
                            const float* zk = zb + k0 + kr0;  // (nothing consumes these loads before the raw tile has been read)
#pragma unroll
                            for (int j = 0; j < kPieces; j++) z[j] = (k_interior || k0 + kr0 + 4 * j < rows_lim) ? __ldg(zk + 4 * j) : INFINITY;
                        }
                    }
                    mbar_wait(raw_bar(grp), raw_phase);
                    raw_phase ^= 1u;
                    float4 v[kPieces];
#pragma unroll
                    for (int j = 0; j < kPieces; j++) v[j] = lds_f4(raw_base + grp * G::RAW_TILE + (uint32_t)(gt + kGroupThreads * j) * 16u);
                    named_barrier_sync(bar_read, kGroupThreads);  // every thread of the group holds its part of the raw tile: the slot is free
                    if (gt == 0) {
                        load_seek(kConvGroups);
                        load_issue();
                    }
                    mbar_wait(empty_bar(stage), phase ^ 1u);  // the tensor core is done with the operand stage's previous contents
                    const uint32_t sA = smem_base + stage * G::STAGE;
                    if (expo) {
                        if (a_mn) {
#pragma unroll
                            for (int j = 0; j < kPieces; j++) z[j] = fmaxf(z[j], zfix[j]);
                        }
#pragma unroll
                        for (int j = 0; j < kPieces; j++) {
                            v[j].x = ex2_approx(fmaf(v[j].x, kLog2e, -z[j]));
                            v[j].y = ex2_approx(fmaf(v[j].y, kLog2e, -z[j]));
                            v[j].z = ex2_approx(fmaf(v[j].z, kLog2e, -z[j]));
                            v[j].w = ex2_approx(fmaf(v[j].w, kLog2e, -z[j]));
                        }
                    }
                    if (want_lo) {
#pragma unroll
                        for (int j = 0; j < kPieces; j++) {
                            const float g[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                            split4_store(sA + dst_off[j], sA + G::A_TILE + dst_off[j], g, true);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < kPieces; j++) {
                            const float g[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                            split4_store(sA + dst_off[j], sA + G::A_TILE + dst_off[j], g, false);
                        }
                    }
                    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    named_barrier_sync(bar_done, kGroupThreads);  // the CTA's A tiles are complete: one CTA-local arrival (the non-leader's is
                    if (gt == 32) mbar_arrive(leader ? full_bar(stage) : afull_bar(stage));  // forwarded to the leader's full barrier by its warp 1)
                }
                if (++stage == NST) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    }
}


// =====================================================================================================================================
// TS variant of the backward contractions: the A operand (G, or G^T) is written by the converter warps straight into TENSOR MEMORY and
// read from there by tcgen05.mma ([a_tmem] operand), instead of going through shared memory.  Two consequences:
//   * the operand stages in shared memory hold only B, so the shared-memory traffic per k-block drops from ~112 KB to ~96 KB *for a tile
//     1.6x as wide* (the smem-A kernel above is bound by shared-memory bandwidth, not by the tensor core);
//   * one tile covers ALL output columns of a 256-row block (N <= 416: accumulator columns [0, 224) and [224, 416), two MMAs per k-step
//     with the same A): every score element is converted once per problem instead of once per column tile, and read once.
// Tensor memory: accumulator 416 columns (single-buffered: the epilogue of a tile is not overlapped with the next tile's MMAs, ~10 % of
// a tile at K = 1000) + 3 A stages of 32 columns (bf16 hi: 16 columns = 32 K, bf16 lo: 16 columns), one per converter group.
// A from tensor memory cannot be transposed: for dNeg = G^T . A the converter thread of operand row m reads COLUMN m of the fp32 score
// tile (conflict-free: consecutive threads, consecutive addresses); for dA = G . Neg it reads row m of a SWIZZLE_128B tile.
struct GeoTS {
    static constexpr int BK = 32;
    static constexpr int STAGES = 4;                    // B operand stages (shared memory)
    static constexpr int A_STAGES = kConvGroups;        // A operand stages (tensor memory)
    static constexpr int PIECE0 = 224;                  // accumulator columns of the first column piece
    static constexpr int ACC_COLS = 416;                // 224 + 192
    static constexpr int A_COLS = 32;                   // per A stage: hi 16 columns, lo 16 columns
    static constexpr int B_TILE = GHALF_N * BK * 2;     // one of hi / lo of one column piece: this CTA's half (8 KB)
    static constexpr int STAGE = 4 * B_TILE;            // piece 0 hi, piece 0 lo, piece 1 hi, piece 1 lo
    static constexpr int RAW_SLOTS = kConvGroups;       // fp32 score tiles in flight: one private slot per converter group
    static constexpr int RAW_TILE = BLOCK_M * BK * 4;
    static constexpr int RAW_OFFSET = STAGES * STAGE;
    static constexpr int EPI_OFFSET = RAW_OFFSET + RAW_SLOTS * RAW_TILE;
    static constexpr int BAR_OFFSET = EPI_OFFSET + kEpilogueSmemBytes;
    static constexpr int SMEM_TOTAL = BAR_OFFSET + 256 + 1024;
    static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
    static_assert(ACC_COLS + A_STAGES * A_COLS <= 512, "tensor memory budget");
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}
// 32 fp32 -> 16 columns of packed bf16 hi + 16 columns of packed bf16 lo (x ~= hi + lo; even k in the low half of a column)
__device__ __forceinline__ void split32(const float (&v)[32], uint32_t (&h)[16], uint32_t (&l)[16]) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
        const float h0 = __uint_as_float(h[j] << 16), h1 = __uint_as_float(h[j] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[j]) : "f"(v[2 * j + 1] - h1), "f"(v[2 * j] - h0));
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1) gemm_tc_ts_kernel(const __grid_constant__ GMaps maps, const GParams p) {
    using G = GeoTS;
    constexpr int BK = G::BK, NST = G::STAGES, NA = G::A_STAGES;
    constexpr int TMEM_COLS = 512;
    constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + G::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // leader: the B tiles of stage s have landed (both CTAs' loads)
    auto empty_bar = [&](int s) { return bar_base + 8u * (NST + s); };        // both: the tensor core is done with B stage s
    const uint32_t tfull_bar = bar_base + 8u * (2 * NST);                     // both: the tile's accumulator is complete
    const uint32_t tempty_bar = bar_base + 8u * (2 * NST + 1);                // leader: all eight epilogue warps have drained it
    const uint32_t tmem_holder = bar_base + 8u * (2 * NST + 2);
    auto raw_bar = [&](int s) { return bar_base + 8u * (2 * NST + 3 + s); };  // this CTA's fp32 score tile in raw slot s has landed
    auto afull_bar = [&](int a) { return bar_base + 8u * (2 * NST + 3 + G::RAW_SLOTS + a); };           // leader: both CTAs wrote A stage a
    auto aloc_bar = [&](int a) { return bar_base + 8u * (2 * NST + 3 + G::RAW_SLOTS + NA + a); };       // non-leader: its part of A stage a is written
    auto aempty_bar = [&](int a) { return bar_base + 8u * (2 * NST + 3 + G::RAW_SLOTS + 2 * NA + a); }; // both: the tensor core is done with A stage a
    static_assert(8 * (2 * NST + 3 + G::RAW_SLOTS + 3 * NA) <= 256, "barrier area");
    volatile uint32_t* tmem_holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int lo_mult = p.passes == 3 ? 2 : 1;

    if (warp == 0 && lane == 0) {
        for (int q = 0; q < 2; q++)
            for (int j = 0; j < 5; j++) prefetch_tmap(&maps.m[q][j]);
        for (int s = 0; s < NST; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < G::RAW_SLOTS; s++) mbar_init(raw_bar(s), 1);
        for (int a = 0; a < NA; a++) {
            mbar_init(afull_bar(a), 2);
            mbar_init(aloc_bar(a), 1);
            mbar_init(aempty_bar(a), 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<TMEM_COLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder_ptr;
    auto ceil32 = [](int x) { return (x + 31) & ~31; };

    if (warp == 0) {
        // ================= TMA producer (both CTAs): the B tiles (MN-major, SWIZZLE_128B: 64-column slabs of BK rows of 128 bytes) =================
        int stage = 0;
        uint32_t phase = 0;
        int4 e_next = p.table[cluster_id];
        for (int r = 0; r < p.rounds; r++) {
            const int4 e = e_next;
            if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
            if (e.x < 0) continue;
            const int pi = e.x & 0xff;
            const GProblem& pr = p.prob[pi];
            const CUtensorMap* mB_hi = &maps.m[pi][2];
            const CUtensorMap* mB_lo = &maps.m[pi][3];
            const int b = e.y;
            const int w0 = min(pr.N, G::PIECE0), w1 = pr.N - w0;
            const int nc0 = (int)rank * (ceil32(w0) / 2), nc1 = G::PIECE0 + (int)rank * (ceil32(w1) / 2);
            const uint32_t tx_pair = (uint32_t)(2 * lo_mult * (w1 > 0 ? 2 : 1) * G::B_TILE);
            const int num_k_blocks = (pr.K + BK - 1) / BK;
            const bool three = p.passes == 3;
            for (int kb = 0; kb < num_k_blocks; kb++) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sB = smem_base + stage * G::STAGE;
                const uint32_t lbar = full_bar(stage) & kPeerMask;
                const int k0 = kb * BK;
                if (elect_one() && (p.debug_flags & 16)) {
                    if (leader) mbar_arrive(full_bar(stage));
                } else if (elect_one()) {
                    if (leader) mbar_expect_tx(full_bar(stage), tx_pair);
#pragma unroll
                    for (int j = 0; j < GHALF_N / 64; j++) {
                        tma_load_3d_2sm(sB + j * (BK * 128), mB_hi, lbar, nc0 + 64 * j, k0, b);
                        if (three) tma_load_3d_2sm(sB + G::B_TILE + j * (BK * 128), mB_lo, lbar, nc0 + 64 * j, k0, b);
                    }
                    if (w1 > 0) {
#pragma unroll
                        for (int j = 0; j < GHALF_N / 64; j++) {
                            tma_load_3d_2sm(sB + 2 * G::B_TILE + j * (BK * 128), mB_hi, lbar, nc1 + 64 * j, k0, b);
                            if (three) tma_load_3d_2sm(sB + 3 * G::B_TILE + j * (BK * 128), mB_lo, lbar, nc1 + 64 * j, k0, b);
                        }
                    }
                }
                __syncwarp();
                if (++stage == NST) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ================= MMA issuer: A from tensor memory, B from shared memory, two column pieces per k-step =================
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0, tile_phase = 0;
            constexpr uint32_t kDescHiMN = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
            constexpr uint32_t b_lbo = ((uint32_t)(BK * 128) >> 4) << 16, b_kstep = 2048u >> 4;
            const uint32_t npass = (p.debug_flags & 1) ? 0u : (uint32_t)p.passes;
            int4 e_next = p.table[cluster_id];
            for (int r = 0; r < p.rounds; r++) {
                const int4 e = e_next;
                if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
                if (e.x < 0) continue;
                const GProblem& pr = p.prob[e.x & 0xff];
                const int w0 = min(pr.N, G::PIECE0), w1 = pr.N - w0;
                const uint32_t idesc0 = make_idesc(GTILE_M, ceil32(w0), false, true);
                const uint32_t idesc1 = make_idesc(GTILE_M, ceil32(max(w1, 1)), false, true);
                const int num_k_blocks = (pr.K + BK - 1) / BK;
                mbar_wait(tempty_bar, tile_phase ^ 1u);  // the epilogue has drained the previous tile
                tile_phase ^= 1u;
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    mbar_wait(afull_bar(astage), aphase);
                    tc_fence_after();
                    const uint32_t sB = (smem_base + stage * G::STAGE) >> 4;  // 16-byte units
                    const uint32_t tA = tmem_base + (uint32_t)(G::ACC_COLS + astage * G::A_COLS);
                    const int k_valid = min(BK, pr.K - kb * BK);
                    const int ksteps = (k_valid + UMMA_K - 1) / UMMA_K;
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t prod = 0; prod < 3; prod++) {  // hi.hi, hi.lo, lo.hi
                            if (prod < npass) {
                                const uint32_t ta = tA + ((prod == 2) ? 16u : 0u);
                                const uint32_t sb0 = b_lbo | (sB + ((prod == 1) ? (uint32_t)(G::B_TILE >> 4) : 0u));
                                const uint32_t sb1 = sb0 + (uint32_t)((2 * G::B_TILE) >> 4);
#pragma unroll
                                for (int ks = 0; ks < BK / UMMA_K; ks++) {
                                    if (ks < ksteps) {
                                        umma_bf16_2sm_ts(tmem_base, ta + 8u * ks, ((uint64_t)kDescHiMN << 32) | (uint64_t)(sb0 + ks * b_kstep), idesc0, accumulate);
                                        if (w1 > 0)
                                            umma_bf16_2sm_ts(tmem_base + (uint32_t)G::PIECE0, ta + 8u * ks, ((uint64_t)kDescHiMN << 32) | (uint64_t)(sb1 + ks * b_kstep),
                                                             idesc1, accumulate);
                                        accumulate = 1;
                                    }
                                }
                            }
                        }
                        umma_commit_2sm(empty_bar(stage));
                        umma_commit_2sm(aempty_bar(astage));
                        if (kb == num_k_blocks - 1) umma_commit_2sm(tfull_bar);
                    }
                    __syncwarp();
                    if (p.debug_flags & 64) mbar_wait(aempty_bar(astage), aphase);  // ablation: one k-block in flight at a time
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                    if (++astage == NA) {
                        astage = 0;
                        aphase ^= 1u;
                    }
                }
            }
        } else {
            // Non-leader CTA: forward "this CTA's part of A stage a is written" to the leader's barrier (see the smem-A kernel above)
            int astage = 0;
            uint32_t aphase = 0;
            int4 e_next = p.table[cluster_id];
            for (int r = 0; r < p.rounds; r++) {
                const int4 e = e_next;
                if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
                if (e.x < 0) continue;
                const int num_k_blocks = (p.prob[e.x & 0xff].K + BK - 1) / BK;
                for (int kb = 0; kb < num_k_blocks; kb++) {
                    mbar_wait(aloc_bar(astage), aphase);
                    if (elect_one()) {
                        if (p.debug_flags & 32) mbar_arrive_remote(afull_bar(astage), 0);
                        else mbar_arrive_remote_light(afull_bar(astage), 0);
                    }
                    __syncwarp();
                    if (++astage == NA) {
                        astage = 0;
                        aphase ^= 1u;
                    }
                }
            }
        }
    } else if (warp < 6) {
        // ================= epilogue warps 2..5 (both CTAs): 32-column chunks of the 416-column accumulator -> TMA store =================
        const int q = warp & 3;
        uint32_t tile_phase = 0;
        uint32_t epi_chunk = 0;
        int4 e_next = p.table[cluster_id];
        for (int r = 0; r < p.rounds; r++) {
            const int4 e = e_next;
            if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
            if (e.x < 0) continue;
            const GProblem& pr = p.prob[e.x & 0xff];
            const CUtensorMap* mD = &maps.m[e.x & 0xff][4];
            const int b = e.y;
            const int m0 = e.z + (int)rank * BLOCK_M;
            mbar_wait(tfull_bar, tile_phase);
            tile_phase ^= 1u;
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            float* drow = pr.D + (int64_t)b * pr.sDb + (int64_t)row * pr.ldd;
            const bool row_ok = row < pr.M;
            const int n_chunks = (p.debug_flags & 4) ? 0 : (pr.N + 31) / 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            uint32_t ra[32], rb[32];
            auto emit = [&](const uint32_t (&rg)[32], int c) {
                const int col0 = c * 32;
                if (pr.tma_store) {
                    stage_and_store(rg, smem_base + G::EPI_OFFSET + (uint32_t)((warp - 2) * 2 + (epi_chunk & 1)) * kStageTileBytes, lane, mD, col0, m0 + q * 32, b);
                    epi_chunk++;
                } else if (row_ok) {
#pragma unroll
                    for (int v = 0; v < 32; v++)
                        if (col0 + v < pr.N) drow[col0 + v] = __uint_as_float(rg[v]);
                }
            };
            if (n_chunks > 0) tmem_ld_32x32b_x32(taddr, ra);
#pragma unroll 1
            for (int c = 0; c < n_chunks; c += 2) {
                tmem_ld_wait();
                if (c + 1 < n_chunks) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 1) * 32), rb);
                emit(ra, c);
                if (c + 1 < n_chunks) {
                    tmem_ld_wait();
                    if (c + 2 < n_chunks) tmem_ld_32x32b_x32(taddr + (uint32_t)((c + 2) * 32), ra);
                    emit(rb, c + 1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty_bar, 0);
        }
        if (elect_one()) bulk_wait_all();
    } else {
        // ================= converter warps 6..17 (both CTAs): fp32 scores -> G -> bf16 hi / lo in tensor memory =================
        // Three groups of four warps on interleaved k-blocks (group g: running k-block numbers g mod 3, A stage g).  Thread (sub-partition
        // q = warp mod 4, lane) owns operand row i = 32 q + lane = TMEM lane i, and produces that row's 32 K values of the k-block.
        constexpr int kGroupThreads = 128;
        const int ct = (int)threadIdx.x - kTcThreads;
        const int grp = ct / kGroupThreads, gt = ct % kGroupThreads;
        const int bar_read = 1 + 2 * grp, bar_done = 2 + 2 * grp;
        const int i = (warp & 3) * 32 + lane;
        const uint32_t raw_base = smem_base + G::RAW_OFFSET;
        const uint32_t tA = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(G::ACC_COLS + grp * G::A_COLS);
        // Raw ring: one slot per converter group.  A barrier that several groups polled in turn would be unsafe: a group returns to a
        // given slot only every 12 k-blocks, so it can start polling two phases ahead of the barrier, and a parity wait cannot tell
        // "two phases behind" from "complete" (seen as a corrupted tile followed by an arrive on a phase that is still waiting for
        // its bytes).  With a private slot the group that waits for use j is the one that consumed use j - 1.  The load of the group's
        // next k-block (3 ahead) is issued as soon as the group holds the current tile in registers: one group period (~3 k-block
        // periods of the tensor core) of lookahead.
        int lr = -1, lkb = 0, lnkb = 0;  // load cursor of thread 0 of the group: table round, k-block, k-blocks of that tile
        int4 le = make_int4(-1, 0, 0, 0);
        auto next_tile = [&]() {
            while (++lr < p.rounds) {
                le = p.table[lr * num_clusters + cluster_id];
                if (le.x >= 0) {
                    lnkb = (p.prob[le.x & 0xff].K + BK - 1) / BK;
                    lkb = 0;
                    return;
                }
            }
        };
        auto load_seek = [&](int steps) {  // the cursor moves `steps` k-blocks forward, across tiles; lr >= p.rounds: past the end
            while (lr < p.rounds) {
                if (lkb + steps < lnkb) {
                    lkb += steps;
                    return;
                }
                steps -= lnkb - lkb;
                next_tile();
            }
        };
        auto load_issue = [&]() {  // TMA of the score tile at the cursor into the group's raw slot
            if (lr >= p.rounds) return;
            const int pi = le.x & 0xff;
            const GProblem& lp = p.prob[pi];
            const int lm = le.z + (int)rank * BLOCK_M;
            mbar_expect_tx(raw_bar(grp), G::RAW_TILE);
            if (lp.a_mn == 0)
                tma_load_3d(raw_base + grp * G::RAW_TILE, &maps.m[pi][0], raw_bar(grp), lkb * BK, lm, le.y);  // box {32 K columns, 128 M rows}, SWIZZLE_128B
            else
                tma_load_3d(raw_base + grp * G::RAW_TILE, &maps.m[pi][0], raw_bar(grp), lm, lkb * BK, le.y);  // box {128 M columns, 32 K rows}, no swizzle
        };
        if (gt == 0) {
            next_tile();
            load_seek(grp);  // the group's first k-block
            load_issue();
        }
        uint32_t raw_phase = 0, aphase = 0;
        int n = 0;
        int4 e_next = p.table[cluster_id];
        for (int r = 0; r < p.rounds; r++) {
            const int4 e = e_next;
            if (r + 1 < p.rounds) e_next = p.table[(r + 1) * num_clusters + cluster_id];
            if (e.x < 0) continue;
            const GProblem& pr = p.prob[e.x & 0xff];
            const int b = e.y;
            const int m_base = e.z + (int)rank * BLOCK_M;
            const int num_k_blocks = (pr.K + BK - 1) / BK;
            const bool a_mn = pr.a_mn != 0, want_lo = p.passes == 3, expo = pr.conv_mode == 1;
            const int rows_lim = pr.conv_rows, cols_lim = pr.conv_cols;
            const float* zb = expo ? pr.conv_z + (int64_t)b * rows_lim : nullptr;
            // the operand row's position in the score matrix: K-major (dA): score row m_base + i; MN-major (dNeg): score column m_base + i
            const bool m_ok = m_base + i < (a_mn ? cols_lim : rows_lim);
            // K-major: the shift of the thread's score row, fixed for the tile; a +inf shift turns a value into exp2(-inf) = 0
            const float zrow = (expo && !a_mn && m_ok) ? __ldg(zb + m_base + i) : INFINITY;
            const int lim_k = a_mn ? rows_lim : cols_lim;
#pragma unroll 1
            for (int kb = 0; kb < num_k_blocks; kb++, n++) {
                if (n % kConvGroups == grp) {
                    const int k0 = kb * BK;
                    const bool k_interior = k0 + BK <= lim_k;
                    // MN-major: lane l holds the shift of score row k0 + l (all lanes use it: it must not depend on this lane's own column;
                    // operand rows beyond the matrix only feed output rows that the store clips)
                    float zl = INFINITY;
                    if (expo && a_mn && k0 + lane < rows_lim) zl = __ldg(zb + k0 + lane);
                    mbar_wait(raw_bar(grp), raw_phase);
                    raw_phase ^= 1u;
                    const uint32_t raw = raw_base + grp * G::RAW_TILE;
                    float v[32];
                    if (!a_mn) {
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const float4 t = lds_f4(raw + (uint32_t)i * 128u + (uint32_t)((c ^ (i & 7)) << 4));
                            v[4 * c] = t.x, v[4 * c + 1] = t.y, v[4 * c + 2] = t.z, v[4 * c + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; k++) v[k] = lds_f32(raw + (uint32_t)k * 512u + (uint32_t)i * 4u);
                    }
                    named_barrier_sync(bar_read, kGroupThreads);  // every thread of the group holds its part of the raw tile: the slot is free
                    if (gt == 0) {
                        load_seek(kConvGroups);
                        load_issue();
                    }
                    if (expo) {
                        if (!a_mn) {
                            const float zs = -zrow;
#pragma unroll
                            for (int k = 0; k < 32; k++) v[k] = ex2_approx(fmaf(v[k], kLog2e, zs));
                            if (!k_interior) {
#pragma unroll
                                for (int k = 0; k < 32; k++)
                                    if (k0 + k >= cols_lim) v[k] = 0.f;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; k++) v[k] = ex2_approx(fmaf(v[k], kLog2e, -__shfl_sync(0xffffffffu, zl, k)));
                        }
                    }
                    uint32_t h[16], l[16];
                    split32(v, h, l);
                    mbar_wait(aempty_bar(grp), aphase ^ 1u);  // the tensor core is done with the stage's previous contents
                    tc_fence_after();
                    if (!(p.debug_flags & 2)) {
                        tmem_st_32x32b_x16(tA, h);
                        if (want_lo) tmem_st_32x32b_x16(tA + 16u, l);
                        tmem_st_wait();
                    }
                    tc_fence_before();
                    named_barrier_sync(bar_done, kGroupThreads);
                    if (gt == 32) mbar_arrive(leader ? afull_bar(grp) : aloc_bar(grp));
                    aphase ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    }
}

// ---- host ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

mb_status bf16_map(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches, uint64_t row_stride, uint64_t batch_stride,
                   uint32_t box_rows) {
    cuuint64_t dims[3] = {inner, rows, batches};
    cuuint64_t strides[2] = {row_stride * 2, batch_stride * 2};
    if (batches == 1) strides[1] = strides[0] * rows;
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (strides[0] & 15u) || (strides[1] & 15u)) {
        set_error("gemm_tc_grouped: operand not 16-byte aligned / stride not a multiple of 16 bytes");
        return MB_ERR_INVALID;
    }
    CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return MB_ERR_CUDA;
    }
    return MB_OK;
}

// tile tables are pure functions of the shapes: build once, keep on the device
struct TableKey {
    int dev, clusters, n, slot64;
    int M[2], N[2], K[2], batches[2];
    bool operator<(const TableKey& o) const {
        return std::tie(dev, clusters, n, slot64, M[0], N[0], K[0], batches[0], M[1], N[1], K[1], batches[1]) <
               std::tie(o.dev, o.clusters, o.n, o.slot64, o.M[0], o.N[0], o.K[0], o.batches[0], o.M[1], o.N[1], o.K[1], o.batches[1]);
    }
};
struct TableVal {
    int4* dev_ptr;
    int rounds;
};
std::map<TableKey, TableVal>& table_cache() {
    static std::map<TableKey, TableVal> c;
    return c;
}
std::mutex& table_mutex() {
    static std::mutex m;
    return m;
}

}  // namespace

// MB_TC_WAITLOG=1: bounded waits that give up leave a record in mapped host memory before the kernel traps (diagnostics)
static unsigned long long* g_wait_log_host = nullptr;
static void wait_log_setup() {
    static std::once_flag once;
    std::call_once(once, [] {
#ifdef MB_WAITLOG
        const char* e = getenv("MB_TC_WAITLOG");
        if (e == nullptr || atoi(e) == 0) return;
        unsigned long long* h = nullptr;
        if (cudaHostAlloc(&h, 1001 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess) return;
        std::memset(h, 0, 1001 * sizeof(unsigned long long));
        unsigned long long* d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return;
        if (cudaMemcpyToSymbol(tcptx::g_wait_log, &d, sizeof(d)) != cudaSuccess) return;
        g_wait_log_host = h;
#endif
    });
}
int gemm_tc_wait_log(unsigned long long* out, int cap) {
    if (g_wait_log_host == nullptr) return 0;
    const int n = (int)std::min<unsigned long long>(g_wait_log_host[0], 500ull);
    for (int i = 0; i < 2 * n && i < cap; i++) out[i] = g_wait_log_host[1 + i];
    return n;
}

bool gemm_tc_supported(int64_t a_inner, int64_t b_inner) {
    // TMA: 16-byte global strides => inner extents (bf16) multiples of 8
    return (a_inner % 8 == 0) && (b_inner % 8 == 0) && encode_fn() != nullptr;
}

mb_status gemm_tc_grouped(const TcGroupProblem* probs, int n, int passes, cudaStream_t st, bool force_smem_a) {
    if (n < 1 || n > 2) {
        set_error("gemm_tc_grouped: 1 or 2 problems");
        return MB_ERR_INVALID;
    }
    if (!encode_fn()) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return MB_ERR_CUDA;
    }
    int dev = 0;
    MB_CUDA_TRY(cudaGetDevice(&dev));
    wait_log_setup();
    const int clusters_max = sm_count() / 2;
    GMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    bool conv = false, with_stats = false;
    GParams p;
    std::memset(&p, 0, sizeof(p));
    p.passes = passes;
    {
        static int dbg = [] { const char* e = getenv("MB_TC_DEBUG"); return e ? atoi(e) : 0; }();
        p.debug_flags = dbg;
    }
    TableKey key;
    std::memset(&key, 0, sizeof(key));
    key.dev = dev;
    key.n = n;
    int64_t total_tiles = 0;
    // backward contractions: A operand through tensor memory, one tile per 256-row block over all columns (gemm_tc_ts_kernel), when
    // every problem's output fits the 416 accumulator columns; MB_CONV_TS=0 keeps the shared-memory-A kernel
    const char* ts_env = getenv("MB_CONV_TS");  // (read per call: tests switch it inside one process)
    bool ts = (ts_env == nullptr || atoi(ts_env) != 0) && !force_smem_a;
    for (int i = 0; i < n; i++) ts = ts && probs[i].conv_mode != 0 && probs[i].N <= GeoTS::ACC_COLS && probs[i].b_mn;
    for (int i = 0; i < n; i++) {
        const TcGroupProblem& g = probs[i];
        if (g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batches <= 0) {
            set_error("gemm_tc_grouped: empty problem");
            return MB_ERR_INVALID;
        }
        const bool lo = passes == 3;
        const bool conv_i = g.conv_mode != 0;
        const uint32_t bk = conv_i ? (uint32_t)Geo<true>::BK : (uint32_t)Geo<false>::BK;  // k-block = K rows of an MN-major box
        if (i == 0) conv = conv_i;
        if (conv_i != conv) {
            set_error("gemm_tc_grouped: all problems of a launch must agree on conv_mode");
            return MB_ERR_INVALID;
        }
        if (conv_i) {
            if (g.conv_src == nullptr || (g.conv_mode == 1 && g.conv_z == nullptr) || (reinterpret_cast<uintptr_t>(g.conv_src) & 15u) || (g.conv_ld % 4) ||
                (g.conv_sb % 4) || (g.conv_cols % 8) || g.conv_rows != (g.a_mn ? g.K : g.M) || g.conv_cols != (g.a_mn ? g.M : g.K) || !g.b_mn) {
                set_error("gemm_tc_grouped: bad conv operand (alignment / extents; the B operand must be MN-major)");
                return MB_ERR_INVALID;
            }
            // the fp32 score matrix, loaded as plain row-major tiles (no swizzle) into the raw ring
            cuuint64_t dims[3] = {(cuuint64_t)g.conv_cols, (cuuint64_t)g.conv_rows, (cuuint64_t)g.batches};
            cuuint64_t strides[2] = {(cuuint64_t)g.conv_ld * 4, (cuuint64_t)(g.batches == 1 ? (int64_t)g.conv_rows * g.conv_ld : g.conv_sb) * 4};
            cuuint32_t box[3] = {g.a_mn ? (cuuint32_t)BLOCK_M : bk, g.a_mn ? bk : (cuuint32_t)BLOCK_M, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            // (TS kernel, K-major: rows of 32 floats = 128 bytes, SWIZZLE_128B, so that a thread per row reads them without bank conflicts)
            CUresult r = encode_fn()(&maps.m[i][0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.conv_src), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, (ts && !g.a_mn) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled (score matrix) failed with CUresult " + std::to_string((int)r));
                return MB_ERR_CUDA;
            }
        } else if (!g.a_mn) {
            MB_TRY(bf16_map(&maps.m[i][0], g.A_hi, g.K, g.M, g.batches, g.lda, g.sAb, BLOCK_M));
            MB_TRY(bf16_map(&maps.m[i][1], lo ? g.A_lo : g.A_hi, g.K, g.M, g.batches, g.lda, g.sAb, BLOCK_M));
        } else {
            MB_TRY(bf16_map(&maps.m[i][0], g.A_hi, g.M, g.K, g.batches, g.lda, g.sAb, bk));
            MB_TRY(bf16_map(&maps.m[i][1], lo ? g.A_lo : g.A_hi, g.M, g.K, g.batches, g.lda, g.sAb, bk));
        }
        if (!g.b_mn) {
            MB_TRY(bf16_map(&maps.m[i][2], g.B_hi, g.K, g.N, g.batches, g.ldb, g.sBb, GHALF_N));
            MB_TRY(bf16_map(&maps.m[i][3], lo ? g.B_lo : g.B_hi, g.K, g.N, g.batches, g.ldb, g.sBb, GHALF_N));
        } else {
            MB_TRY(bf16_map(&maps.m[i][2], g.B_hi, g.N, g.K, g.batches, g.ldb, g.sBb, bk));
            MB_TRY(bf16_map(&maps.m[i][3], lo ? g.B_lo : g.B_hi, g.N, g.K, g.batches, g.ldb, g.sBb, bk));
        }
        bool tma_store = ((reinterpret_cast<uintptr_t>(g.D) & 15u) == 0) && (g.ldd % 4 == 0) && (g.batches == 1 || g.sDb % 4 == 0);
        if (tma_store) {
            cuuint64_t dims[3] = {(cuuint64_t)g.N, (cuuint64_t)g.M, (cuuint64_t)g.batches};
            cuuint64_t strides[2] = {(cuuint64_t)g.ldd * 4, (cuuint64_t)(g.batches == 1 ? (int64_t)g.M * g.ldd : g.sDb) * 4};
            cuuint32_t box[3] = {32, 32, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = encode_fn()(&maps.m[i][4], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g.D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) tma_store = false;
        }
        if (!tma_store) maps.m[i][4] = maps.m[i][2];
        if (conv_i) maps.m[i][1] = maps.m[i][2];  // (never used: a valid descriptor for prefetch.tensormap)
        GProblem& q = p.prob[i];
        q.D = g.D;
        q.ldd = g.ldd;
        q.sDb = g.sDb;
        q.M = g.M;
        q.N = g.N;
        q.K = g.K;
        q.batches = g.batches;
        q.a_mn = g.a_mn;
        q.b_mn = g.b_mn;
        q.tma_store = tma_store ? 1 : 0;
        q.stats = conv_i ? nullptr : g.stats;
        q.stat_slots = g.stat_slots;
        q.conv_src = g.conv_src;
        q.conv_z = g.conv_z;
        q.conv_ld = g.conv_ld;
        q.conv_sb = g.conv_sb;
        q.conv_rows = g.conv_rows;
        q.conv_cols = g.conv_cols;
        q.conv_mode = g.conv_mode;
        if (q.stats != nullptr && (g.stat_slots < (g.N + kTcStatSlotCols - 1) / kTcStatSlotCols)) {
            set_error("gemm_tc_grouped: stat_slots too small");
            return MB_ERR_INVALID;
        }
        with_stats = with_stats || q.stats != nullptr;
        key.M[i] = g.M;
        key.N[i] = g.N;
        key.K[i] = g.K;
        key.batches[i] = g.batches;
        total_tiles += (int64_t)((g.M + GTILE_M - 1) / GTILE_M) * (ts ? 1 : (g.N + GTILE_N - 1) / GTILE_N) * g.batches;
    }
    if (n == 1) {
        p.prob[1] = p.prob[0];
        for (int j = 0; j < 5; j++) maps.m[1][j] = maps.m[0][j];
    }
    int clusters = (int)std::min<int64_t>(clusters_max, total_tiles);
    key.clusters = clusters;
    key.slot64 = ts ? 2 : (with_stats ? 1 : 0);  // statistics slots are 64 columns wide: column pieces of the tail tiles must start on multiples of 64
    TableVal tv;
    {
        std::lock_guard<std::mutex> lk(table_mutex());
        auto it = table_cache().find(key);
        if (it == table_cache().end()) {
            // Tiles in batch-major order (both problems of a batch next to each other: they read the same G / Neg / A blocks, which
            // then come from L2 instead of HBM), handed to the least-loaded CTA pair (list scheduling).  The last tiles of the list
            // are split into narrower column pieces so the tail is filled to within one small piece; the split is chosen by
            // simulating the schedule with the cost model below.
            struct Tile {
                int prob, b, m0, n0, width;
                int64_t cost;
            };
            int ksteps[2] = {0, 0}, max_batches = 0;
            for (int i = 0; i < n; i++) {
                ksteps[i] = (probs[i].K + 15) / 16;
                max_batches = std::max(max_batches, probs[i].batches);
            }
            auto cost_of = [&](int prob, int width) {  // MMA time ~ n_eff per k-step; + pipeline fill / epilogue drain per tile
                const int w0 = ts ? std::min(width, (int)GeoTS::PIECE0) : width;
                return (int64_t)(((w0 + 31) / 32) * 32 + ((width - w0 + 31) / 32) * 32) * ksteps[prob] + 1024;
            };
            std::vector<Tile> base;
            for (int b = 0; b < max_batches; b++)
                for (int i = 0; i < n; i++) {
                    const TcGroupProblem& g = probs[i];
                    if (b >= g.batches) continue;
                    const int tile_n = ts ? (int)GeoTS::ACC_COLS : GTILE_N;  // TS kernel: one tile covers all columns
                    for (int m0 = 0; m0 < g.M; m0 += GTILE_M)
                        for (int n0 = 0; n0 < g.N; n0 += tile_n) {
                            const int w = std::min(tile_n, g.N - n0);
                            base.push_back({i, b, m0, n0, w, cost_of(i, w)});
                        }
                }
            auto build = [&](int tail, int piece) {
                std::vector<Tile> out;
                const size_t keep = base.size() - std::min<size_t>(base.size(), (size_t)tail);
                for (size_t t = 0; t < base.size(); t++) {
                    const Tile& x = base[t];
                    if (t < keep || x.width <= piece) {
                        out.push_back(x);
                        continue;
                    }
                    for (int o = 0; o < x.width; o += piece) {
                        const int w = std::min(piece, x.width - o);
                        out.push_back({x.prob, x.b, x.m0, x.n0 + o, w, cost_of(x.prob, w)});
                    }
                }
                return out;
            };
            auto schedule = [&](const std::vector<Tile>& tiles, std::vector<std::vector<int>>* per_pair) {
                std::vector<int64_t> load((size_t)clusters, 0);
                if (per_pair) per_pair->assign((size_t)clusters, {});
                for (size_t t = 0; t < tiles.size(); t++) {
                    int best = 0;
                    for (int c = 1; c < clusters; c++)
                        if (load[c] < load[best]) best = c;
                    load[best] += tiles[t].cost;
                    if (per_pair) (*per_pair)[best].push_back((int)t);
                }
                return *std::max_element(load.begin(), load.end());
            };
            int best_tail = 0, best_piece = ts ? (int)GeoTS::ACC_COLS : GTILE_N;
            int64_t best_span = schedule(base, nullptr);
            for (int piece : {192, 128, 96, 64}) {
                if (ts) break;  // (full-width tiles are not split)
                if (with_stats && piece % kTcStatSlotCols != 0) continue;
                for (int k = 1; k <= 16; k++) {
                    const int tail = clusters * k / 4;
                    const int64_t span = schedule(build(tail, piece), nullptr);
                    if (span < best_span) {
                        best_span = span;
                        best_tail = tail;
                        best_piece = piece;
                    }
                }
            }
            const std::vector<Tile> tiles = build(best_tail, best_piece);
            std::vector<std::vector<int>> per_pair;
            schedule(tiles, &per_pair);
            int rounds = 0;
            for (auto& v : per_pair) rounds = std::max(rounds, (int)v.size());
            std::vector<int4> host((size_t)rounds * clusters, make_int4(-1, 0, 0, 0));
            for (int c = 0; c < clusters; c++)
                for (size_t r = 0; r < per_pair[c].size(); r++) {
                    const Tile& x = tiles[per_pair[c][r]];
                    host[r * clusters + c] = make_int4(x.prob | (x.width << 8), x.b, x.m0, x.n0);
                }
            int4* dptr = nullptr;
            MB_CUDA_TRY(cudaMalloc(&dptr, host.size() * sizeof(int4)));
            MB_CUDA_TRY(cudaMemcpy(dptr, host.data(), host.size() * sizeof(int4), cudaMemcpyHostToDevice));
            tv = {dptr, rounds};
            table_cache()[key] = tv;
        } else {
            tv = it->second;
        }
    }
    p.table = tv.dev_ptr;
    p.rounds = tv.rounds;
    // function attributes are per device (one process may drive several: the reference's device_models_): opt in once on each
    static std::atomic<bool> attr_set[64];
    if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
        MB_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<false>::SMEM_TOTAL));
        MB_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<true>::SMEM_TOTAL));
        MB_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GeoTS::SMEM_TOTAL));
        if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
    }
    if (ts)
        gemm_tc_ts_kernel<<<2 * clusters, kConvThreads, GeoTS::SMEM_TOTAL, st>>>(maps, p);
    else if (conv)
        gemm_tc_group_kernel<true><<<2 * clusters, kConvThreads, Geo<true>::SMEM_TOTAL, st>>>(maps, p);
    else
        gemm_tc_group_kernel<false><<<2 * clusters, kTcThreads, Geo<false>::SMEM_TOTAL, st>>>(maps, p);
    MB_LAUNCH_CHECK();
    return MB_OK;
}

}  // namespace mb
