// c_api.cu -- the extern "C" boundary (include/marius_b200.h) and the per-batch orchestration of the hot path.
#include <cuda_bf16.h>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.h"

namespace mb {

static thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { g_error = msg; }

// Process teardown: once exit() has started the CUDA runtime may already be unloading, and destroying streams / events / graphs then
// is undefined.  The handler is registered after the first CUDA call of mb_create, so it runs BEFORE the runtime's own exit hooks
// (atexit is LIFO); from then on mb_destroy only drops the host object and leaves device resources to the driver.
static std::atomic<bool> g_exiting{false};
static void on_process_exit() { g_exiting.store(true); }

// MB_SEGV_TRACE=1: print a native backtrace on SIGSEGV / SIGABRT (debugging aid, off by default)
static void segv_trace(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "[marius_b200] fatal signal, native backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

// bump allocator over the context's workspace; run once with base == nullptr to size, once to place
struct Arena {
    char* base;
    size_t off = 0;
    explicit Arena(char* b) : base(b) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

}  // namespace mb

struct mb_context {
    int device = 0;
    char* ws = nullptr;
    size_t ws_bytes = 0;
    // device staging for mb_train_step_host
    // host-buffer steps: the index tensors of step i+1 cross PCIe on the copy stream, into one of two device slots, while step i runs
    int64_t* pf[2] = {nullptr, nullptr};
    size_t pf_cap[2] = {0, 0};
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    float* h_loss = nullptr;
    int* d_flag = nullptr;           // persistent device error flag (score-filter range check): checked once per evaluate call
    float* h_loss_pinned = nullptr;  // two pinned host landing slots of the step's loss, used alternately (mb_train_step_host_async keeps
                                     // one step in flight while the caller reads the previous step's loss)
    int loss_slot = 0;
    cudaEvent_t ev_loss[2] = {nullptr, nullptr};
    std::vector<void*> ipc_mappings;      // peer shards opened with mb_ipc_import (closed in mb_destroy)
    cudaStream_t side = nullptr;          // index plans (slot / relation sorts) overlap the forward pass here
    cudaStream_t side2 = nullptr;         // the dNeg contraction runs here, concurrently with dA + edge_backward
    cudaStream_t gstream = nullptr;       // graphs are captured / replayed here (the caller's stream may be the legacy default stream,
                                          // which cannot be captured); ordered against the caller's stream with ev_in / ev_out
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
    cudaEvent_t ev_sfetch_fork = nullptr, ev_sfetch_join = nullptr, ev_efetch_join = nullptr;  // sharded step: remote state rows are fetched on side2 during the contractions
    cudaEvent_t ev_slot = nullptr;        // the node-slot plan is ready (the node update waits on this, not on the relation plan)
    // CUDA-graph replay of the fused step (mb_train_step / mb_train_step_host): one captured graph per call signature
    struct StepGraph {
        bool valid = false;
        int warm = 0;            // eager runs seen with this key (the first run is never captured: it may allocate / set attributes)
        std::vector<uint64_t> key;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t n_uniq = nullptr, n_edges = nullptr, n_dneg = nullptr, n_sneg = nullptr, n_loss = nullptr;
    };
    StepGraph sg;
    int64_t* raw = nullptr;      // raw-edge steps: device staging of (global edges | dst_negs | src_negs | unique ids | local edges | local negs | count)
    size_t raw_cap = 0;
    int graphs_enabled = -1;     // MB_GRAPH env (default on)
    int64_t* g_uniq = nullptr;   // index staging with fixed addresses (graph kernels read these)
    int64_t* g_edges = nullptr;
    int64_t* g_dneg = nullptr;
    int64_t* g_sneg = nullptr;
    size_t g_uniq_cap = 0, g_edges_cap = 0, g_dneg_cap = 0, g_sneg_cap = 0;
    // optional per-stage CUDA-event timing (mb_profile_*): events are recorded on the caller's stream
    int profiling = 0;  // 0 off, 1 stage timing with the side streams folded into the caller's stream, 2 timeline (streams stay concurrent)
    struct Span {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

namespace mb {

static void drop_graph(mb_context* ctx) {
    if (ctx->sg.exec) cudaGraphExecDestroy(ctx->sg.exec);
    if (ctx->sg.graph) cudaGraphDestroy(ctx->sg.graph);
    ctx->sg = mb_context::StepGraph();
}

static mb_status ensure_ws(mb_context* ctx, size_t bytes, cudaStream_t st) {
    if (bytes <= ctx->ws_bytes) return MB_OK;
    MB_CUDA_TRY(cudaStreamSynchronize(st));
    drop_graph(ctx);  // captured kernels hold pointers into the old workspace
    if (ctx->ws) MB_CUDA_TRY(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = bytes + bytes / 8 + (1 << 20);
    cudaError_t e = cudaMalloc(&ctx->ws, want);
    if (e != cudaSuccess) {
        set_error("workspace allocation of " + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
        return MB_ERR_NOMEM;
    }
    ctx->ws_bytes = want;
    return MB_OK;
}

enum Stage { ST_GATHER = 0, ST_SORT, ST_REL_SORT, ST_PREP, ST_GEMM_FWD, ST_LOSS, ST_GEMM_DA, ST_GEMM_DNEG, ST_EDGE_BWD, ST_UPDATE, ST_REL_GRAD, ST_EXCHANGE, ST_SAMPLE, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"gather_rows", "slot_sort", "rel_sort", "edge_prep+neg_gather", "gemm_scores", "loss_grad", "gemm_dA",
                                            "gemm_dNeg", "edge_backward", "segment_reduce+adagrad_update", "rel_grad_reduce", "shard_barriers+owner_apply", "negative_sampling+unique_mapping"};

struct StageTimer {
    mb_context* ctx;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    int stage;
    static cudaEvent_t get_event(mb_context* c) {
        if (!c->event_pool.empty()) {
            cudaEvent_t e = c->event_pool.back();
            c->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    StageTimer(mb_context* c, int stage_, cudaStream_t s) : ctx(c), st(s), stage(stage_) {
        if (ctx && ctx->profiling) {
            a = get_event(ctx);
            b = get_event(ctx);
            cudaEventRecord(a, st);
        }
    }
    ~StageTimer() {
        if (a) {
            cudaEventRecord(b, st);
            ctx->spans.push_back({stage, a, b});
        }
    }
};

static int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) b++;
    return b;
}

// Everything one batch needs, carved out of the workspace.
struct Plan {
    // dims
    int64_t U, d, B, Bc, Bp, R, CN, n_slots;
    int stat_slots;  // 64-column statistics slots per score row (tensor-core path: the loss is fused into the contractions)
    int C, N, sides, cols;
    bool has_rel, use_tc;
    // buffers
    float *emb_u, *A, *pos, *gpos, *row_loss, *NegE, *S, *dA, *gcat, *drel;
    int64_t* bounds = nullptr;         // sharded table: [world + 1] owner bounds of the sorted unique-id list
    const float** row_ptrs = nullptr;  // sharded table: address of every unique row (own HBM or the fetched copy in emb_u)
    bool sharded = false;
    __nv_bfloat16 *A_hl, *Neg_hl;
    float2* stats = nullptr;  // [sides*Bp][stat_slots] (max, sum exp) per score-row slot, written by the forward contraction's epilogue
    float* zw = nullptr;      // [sides*Bp] log-partition of every score row (minus log of the loss weight): G = exp(S - zw)
    float* block_loss = nullptr;  // [ceil(sides*Bp / 16)] per-block sums of the row losses (loss_merge_kernel)
    uint32_t *keys_a, *keys_b, *vals_a, *vals_b, *offsets, *hist;
    uint32_t *rkeys_a, *rkeys_b, *rvals_a, *rvals_b, *roffsets, *rhist;

    void layout(Arena& ar, bool need_emb_u, bool training, bool own_scores) {
        emb_u = need_emb_u ? ar.take<float>(U * d) : nullptr;
        row_ptrs = sharded ? ar.take<const float*>(U) : nullptr;
        bounds = sharded ? ar.take<int64_t>(16) : nullptr;
        A = ar.take<float>(sides * Bp * d);
        pos = ar.take<float>(sides * Bp);
        NegE = use_tc ? nullptr : ar.take<float>(sides * CN * d);
        S = own_scores ? ar.take<float>(sides * Bp * N) : nullptr;
        A_hl = use_tc ? ar.take<__nv_bfloat16>(2 * sides * Bp * d) : nullptr;
        Neg_hl = use_tc ? ar.take<__nv_bfloat16>(2 * sides * CN * d) : nullptr;
        gpos = row_loss = dA = gcat = drel = nullptr;
        stats = nullptr;
        zw = nullptr;
        keys_a = keys_b = vals_a = vals_b = offsets = hist = nullptr;
        rkeys_a = rkeys_b = rvals_a = rvals_b = roffsets = rhist = nullptr;
        if (training) {
            gpos = ar.take<float>(sides * Bp);
            row_loss = ar.take<float>(sides * Bp);
            dA = ar.take<float>(sides * Bp * d);
            gcat = ar.take<float>(n_slots * d);
            stats = use_tc ? ar.take<float2>(sides * Bp * stat_slots) : nullptr;
            zw = use_tc ? ar.take<float>(sides * Bp) : nullptr;
            block_loss = use_tc ? ar.take<float>(loss_merge_blocks(sides * Bp) + 1) : nullptr;
            keys_a = ar.take<uint32_t>(n_slots);
            keys_b = ar.take<uint32_t>(n_slots);
            vals_a = ar.take<uint32_t>(n_slots);
            vals_b = ar.take<uint32_t>(n_slots);
            offsets = ar.take<uint32_t>(U + 2);
            hist = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(n_slots)));
            if (has_rel) {
                drel = ar.take<float>(sides * B * d);
                rkeys_a = ar.take<uint32_t>(B);
                rkeys_b = ar.take<uint32_t>(B);
                rvals_a = ar.take<uint32_t>(B);
                rvals_b = ar.take<uint32_t>(B);
                roffsets = ar.take<uint32_t>(R + 2);
                rhist = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(B)));
            }
        }
    }
};

static mb_status validate_batch(const mb_batch* b) {
    MB_REQUIRE(b != nullptr, "batch is null");
    MB_REQUIRE(b->decoder >= MB_DECODER_DOT && b->decoder <= MB_DECODER_COMPLEX, "unknown decoder kind");
    MB_REQUIRE(b->U >= 0 && b->d > 0 && b->B >= 0 && b->C > 0 && b->N > 0, "bad batch dimensions");
    MB_REQUIRE(b->edge_cols == 2 || b->edge_cols == 3, "Edge list must be a 3 or 2 column tensor");  // decoder_methods.cpp:66-72
    MB_REQUIRE(b->edges != nullptr || b->B == 0, "edges is null");
    MB_REQUIRE(b->dst_negs != nullptr, "dst_negs is null");
    if (b->decoder != MB_DECODER_DOT && b->edge_cols == 3) MB_REQUIRE(b->rel != nullptr && b->R > 0, "relations required for DistMult/ComplEx");
    if (b->decoder == MB_DECODER_COMPLEX) MB_REQUIRE(b->d % 2 == 0, "ComplEx needs an even embedding dimension");
    MB_REQUIRE(b->U < ((int64_t)1 << 31), "U too large");
    return MB_OK;
}

static void fill_plan_dims(Plan& p, const mb_batch* b, int precision) {
    p.U = b->U;
    p.d = b->d;
    p.B = b->B;
    p.C = b->C;
    p.N = b->N;
    p.R = b->R;
    p.cols = b->edge_cols;
    p.Bc = (b->B + b->C - 1) / b->C;  // ceil(B / C)   comparators.cpp:9
    if (p.Bc == 0) p.Bc = 0;
    p.Bp = p.Bc * b->C;
    p.CN = (int64_t)b->C * b->N;
    p.has_rel = (b->edge_cols == 3) && (b->decoder != MB_DECODER_DOT) && b->rel != nullptr;
    p.sides = (p.has_rel && b->inv_rel != nullptr && b->src_negs != nullptr) ? 2 : 1;  // use_inverse_relations_ (decoder_methods.cpp:90)
    p.n_slots = 2 * p.B + 2 * p.CN;
    // tensor-core contractions take their bf16 hi/lo operands from the vectorised row kernels (d <= 512); wider rows use the general-d
    // row kernels with the fp32 FFMA contraction
    p.use_tc = (precision != MB_PREC_FP32) && gemm_tc_supported(p.d, p.N) && p.Bc > 0 && p.d <= 512;
    p.stat_slots = (p.N + kTcStatSlotCols - 1) / kTcStatSlotCols;
}

static mb_status tc_contract(const void* A_hi, const void* A_lo, int64_t lda, int64_t sAb, bool a_mn, const void* B_hi, const void* B_lo, int64_t ldb,
                             int64_t sBb, bool b_mn, float* D, int64_t ldd, int64_t sDb, int M, int N, int K, int batches, int passes, cudaStream_t st,
                             float2* stats = nullptr, int stat_slots = 0) {
    TcGroupProblem g{A_hi, A_lo, lda, sAb, a_mn ? 1 : 0, B_hi, B_lo, ldb, sBb, b_mn ? 1 : 0, D, ldd, sDb, M, N, K, batches};
    g.stats = stats;
    g.stat_slots = stat_slots;
    return gemm_tc_grouped(&g, 1, passes, st);
}

// forward: adjusted rows A, positive scores, negative rows, score GEMM.  S0/S1 are the score outputs per side ([Bp,N] each).
static mb_status run_forward(mb_context* ctx, const Plan& p, const mb_batch* b, const float* emb, int64_t emb_ld, const int64_t* row_map, int precision,
                             float* pos, float* S0, float* S1, bool uniform_S, cudaStream_t st, bool skip_scores = false, const float* const* row_ptrs = nullptr,
                             const mb_shards* sh = nullptr, cudaEvent_t rows_fetched = nullptr, float2* stats = nullptr) {
    const int d = (int)p.d;
    const int64_t a_half = p.sides * p.Bp * d;  // hi block then lo block, each [sides][Bp][d]
    const int64_t n_half = p.sides * p.CN * d;
    {
        StageTimer tm(ctx, ST_PREP, st);
        // the fp32 adjusted rows are only read by the SIMT GEMM and by the scalar (general-d) backward kernel
        const bool need_A = !p.use_tc || !decoder_vec_ok(emb, emb_ld, d, p.has_rel, b->rel, p.sides == 2 ? b->inv_rel : nullptr, p.sides);
        MB_TRY(launch_prep(sh, rows_fetched, row_ptrs, emb, emb_ld, row_map, b->edges, p.cols, b->rel, p.sides == 2 ? b->inv_rel : nullptr, p.B, p.Bp, p.CN, d, b->decoder, p.sides,
                           b->dst_negs, p.sides == 2 ? b->src_negs : nullptr, need_A ? p.A : nullptr, pos, p.use_tc ? (void*)p.A_hl : nullptr,
                           p.use_tc ? (void*)(p.A_hl + a_half) : nullptr, p.NegE, p.use_tc ? (void*)p.Neg_hl : nullptr,
                           p.use_tc ? (void*)(p.Neg_hl + n_half) : nullptr, st));
    }
    if (p.Bc == 0 || skip_scores) return MB_OK;
    const int passes = precision == MB_PREC_BF16 ? 1 : 3;
    // fused loss: the epilogue of the score contraction leaves (max, sum exp) of every 64-column slot of every score row in `stats`
    if (stats != nullptr && p.use_tc && uniform_S) MB_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(float2) * p.sides * p.Bp * p.stat_slots, st));
    StageTimer tm_fwd(ctx, ST_GEMM_FWD, st);
    // scores[side][chunk] = A[side][chunk] . Neg[side][chunk]^T     (comparators.cpp:69-72)
    int launches = uniform_S ? 1 : p.sides;
    for (int l = 0; l < launches; l++) {
        int batches = uniform_S ? p.sides * p.C : p.C;
        float* S = l == 0 ? S0 : S1;
        int64_t aoff = (int64_t)l * p.Bp * d, noff = (int64_t)l * p.CN * d;
        if (p.use_tc) {
            MB_TRY(tc_contract(p.A_hl + aoff, p.A_hl + a_half + aoff, d, p.Bc * d, false, p.Neg_hl + noff, p.Neg_hl + n_half + noff, d,
                               (int64_t)p.N * d, false, S, p.N, p.Bc * p.N, (int)p.Bc, p.N, d, batches, passes, st, uniform_S ? stats : nullptr,
                               p.stat_slots));
        } else {
            MB_TRY(gemm_simt(p.A + aoff, d, 1, p.Bc * d, p.NegE + noff, 1, d, (int64_t)p.N * d, S, p.N, p.Bc * p.N, (int)p.Bc, p.N, d, batches, st));
        }
    }
    return MB_OK;
}

enum class UpdateMode { kBatchLocal, kFusedTable };

// The duplicate-accumulation plans (slot list sorted by node id, edges sorted by relation id) only depend on the batch's
// index tensors, so they run on the context's side stream concurrently with gather / prep / the score GEMM.
static mb_status run_index_plans(mb_context* ctx, const Plan& p, const mb_batch* b, bool need_rel, uint32_t** svals, uint32_t** rvals, cudaStream_t st,
                                 cudaEvent_t slot_plan_ready = nullptr) {
    uint32_t* skeys = nullptr;
    {
        StageTimer tm(ctx, ST_SORT, st);
        MB_TRY(launch_slot_keys(b->edges, p.cols, p.B, b->dst_negs, p.sides == 2 ? b->src_negs : nullptr, p.CN, p.keys_a, st));
        MB_TRY(radix_sort_pairs<uint32_t>(p.keys_a, p.keys_b, p.vals_a, p.vals_b, p.n_slots,
                                          p.sides == 2 ? bits_for((uint64_t)std::max<int64_t>(p.U, 1)) : 32, p.hist, &skeys, svals, st));
        MB_TRY(segment_offsets_u32(skeys, p.n_slots, p.U, p.offsets, st));
    }
    if (slot_plan_ready) MB_CUDA_TRY(cudaEventRecord(slot_plan_ready, st));
    if (need_rel) {
        StageTimer tm(ctx, ST_REL_SORT, st);
        uint32_t* rk = nullptr;
        MB_TRY(launch_rel_keys(b->edges, p.cols, p.B, p.rkeys_a, st));
        MB_TRY(radix_sort_pairs<uint32_t>(p.rkeys_a, p.rkeys_b, p.rvals_a, p.rvals_b, p.B, bits_for((uint64_t)p.R), p.rhist, &rk, rvals, st));
        MB_TRY(segment_offsets_u32(rk, p.B, p.R, p.roffsets, st));
    }
    return MB_OK;
}

static mb_status run_train(mb_context* ctx, const mb_batch* b, const float* emb_in, int64_t emb_ld, const float* state, int64_t state_ld, float* table,
                           float* state_table, int64_t ld, const int64_t* unique_ids, float lr, int reduction, int precision, float* loss,
                           float* grad, float* delta_e, float* delta_s, float* rel_grad, float* inv_rel_grad, UpdateMode mode, cudaStream_t st,
                           const float* const* ext = nullptr /* {gpos, gneg, ginv_pos, ginv_neg}: upstream gradients instead of the fused loss */,
                           const mb_shards* sh = nullptr /* table sharded over peer GPUs: unique_ids are global rows */) {
    MB_TRY(validate_batch(b));
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    MB_REQUIRE(reduction == MB_REDUCTION_MEAN || reduction == MB_REDUCTION_SUM, "unknown reduction");
    Plan p;
    fill_plan_dims(p, b, precision);
    const bool fused = mode == UpdateMode::kFusedTable;
    p.sharded = fused && sh != nullptr && sh->world > 1;
    {
        Arena sizing(nullptr);
        p.layout(sizing, fused, true, true);
        MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
        Arena place(ctx->ws);
        p.layout(place, fused, true, true);
    }
    const int d = (int)p.d;
    const bool need_rel = p.has_rel && (rel_grad != nullptr || inv_rel_grad != nullptr) && p.R > 0;

    // ---- fork: index plans on the side stream (inline when profiling so that stage times stay meaningful)
    uint32_t *svals = nullptr, *rvals = nullptr;
    const bool overlap = ctx->profiling != 1 && ctx->side != nullptr;
    if (overlap) {
        MB_CUDA_TRY(cudaEventRecord(ctx->ev_fork, st));
        MB_CUDA_TRY(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        MB_TRY(run_index_plans(ctx, p, b, need_rel, &svals, &rvals, ctx->side, ctx->ev_slot));
    } else {
        MB_TRY(run_index_plans(ctx, p, b, need_rel, &svals, &rvals, st));
    }

    const float* emb = emb_in;
    const int64_t* row_map = nullptr;
    bool emb_fetch_forked = false;
    if (fused) {
        // DataLoader::loadGPUParameters (dataloader.cpp:529-548).  With the vector kernels the gather is fused away: prep / backward read
        // table[unique_ids[local id]] directly (the table is not modified until the update at the end of the step, so this is the same
        // snapshot the reference's gathered copy holds); the state rows are read in place by the update kernel.
        if (decoder_vec_ok(table, ld, d, p.has_rel, b->rel, p.sides == 2 ? b->inv_rel : nullptr, p.sides)) {
            emb = table;
            emb_ld = ld;
            row_map = unique_ids;
            if (p.sharded) {
                // barrier 1: every rank has applied the gradient rows of the previous step, so the rows fetched below are the tables after
                // step t-1 everywhere; then the owner bounds of this batch (and the row counts the owners will receive)
                {
                    StageTimer tm(ctx, ST_EXCHANGE, st);
                    MB_TRY(launch_shard_barrier(sh, st));
                    MB_TRY(launch_owner_bounds(sh, unique_ids, p.U, d, p.bounds, st));
                }
                // remote rows are copied once into the batch cache; prep / backward then address every row through row_ptrs.  The fetch
                // runs on the second side stream while the (local) negative rows are prepared.
                cudaStream_t fs = (overlap && ctx->side2 != nullptr) ? ctx->side2 : st;
                if (fs != st) {
                    MB_CUDA_TRY(cudaEventRecord(ctx->ev_sfetch_fork, st));
                    MB_CUDA_TRY(cudaStreamWaitEvent(fs, ctx->ev_sfetch_fork, 0));
                }
                {
                    StageTimer tm(ctx, ST_GATHER, fs);
                    MB_TRY(launch_fetch_remote_rows(sh, unique_ids, p.U, ld, d, p.emb_u, p.row_ptrs, false, fs));
                }
                if (fs != st) {
                    MB_CUDA_TRY(cudaEventRecord(ctx->ev_efetch_join, fs));
                    emb_fetch_forked = true;
                }
            }
        } else if (sh != nullptr && sh->world > 1) {
            set_error("the sharded step needs the vector kernels (d % 8 == 0, 16-byte aligned tables)");
            return MB_ERR_UNSUPPORTED;
        } else {
            StageTimer tm(ctx, ST_GATHER, st);
            MB_TRY(gather_rows(table, ld, d, unique_ids, p.U, p.emb_u, d, st));
            emb = p.emb_u;
            emb_ld = d;
        }
    }

    MB_TRY(run_forward(ctx, p, b, emb, emb_ld, row_map, precision, p.pos, p.S, p.S + p.Bp * p.N, true, st, ext != nullptr, p.row_ptrs, p.sharded ? sh : nullptr,
                       emb_fetch_forked ? ctx->ev_efetch_join : nullptr, ext == nullptr ? p.stats : nullptr));

    // SoftmaxCrossEntropy forward + gradient (loss.cpp:50-67); both sides in one launch (rows = sides*Bp)
    const int64_t rows = p.sides * p.Bp;
    const float w = reduction == MB_REDUCTION_SUM ? 1.0f : (p.Bp > 0 ? 1.0f / (float)p.Bp : 0.f);
    bool merged_loss = false;
    if (rows > 0 && ext != nullptr) {
        // generic autograd path: the caller's loss produced d loss / d (pos, neg, inv_pos, inv_neg); the backward contractions read the
        // fp32 gradient matrix directly (converter warps, conv_mode 2)
        for (int sd = 0; sd < p.sides; sd++) {
            MB_REQUIRE(ext[2 * sd] != nullptr && ext[2 * sd + 1] != nullptr, "upstream gradients missing");
            MB_CUDA_TRY(cudaMemcpyAsync(p.gpos + sd * p.Bp, ext[2 * sd], sizeof(float) * p.Bp, cudaMemcpyDeviceToDevice, st));
            MB_CUDA_TRY(cudaMemcpyAsync(p.S + sd * p.Bp * p.N, ext[2 * sd + 1], sizeof(float) * p.Bp * p.N, cudaMemcpyDeviceToDevice, st));
        }
    } else if (rows > 0 && p.use_tc) {
        // SoftmaxCrossEntropy fused into the contractions: the forward epilogue produced the row statistics, this merges them into the
        // log-partition z of every row (-> row loss, d loss / d pos); the gradient matrix G = exp(S - z) itself is produced inside the
        // backward contractions and never written to memory
        StageTimer tm(ctx, ST_LOSS, st);
        MB_TRY(launch_loss_merge(p.stats, p.stat_slots, p.pos, p.gpos, p.row_loss, p.zw, rows, w, p.block_loss, st));
        merged_loss = true;
    } else if (rows > 0) {
        StageTimer tm(ctx, ST_LOSS, st);
        MB_TRY(launch_loss(p.S, p.S, p.pos, p.gpos, p.row_loss, nullptr, nullptr, rows, p.N, w, st, p.N));
    }
    if (loss && ext == nullptr) {
        if (rows > 0 && merged_loss)
            MB_TRY(launch_loss_reduce(p.block_loss, loss_merge_blocks(rows), loss, st));  // (rows / 16 terms, fixed order)
        else if (rows > 0)
            MB_TRY(launch_loss_reduce(p.row_loss, rows, loss, st));
        else
            MB_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
    }
    const int passes = precision == MB_PREC_BF16 ? 1 : 3;
    const int batches = p.sides * p.C;
    bool dneg_forked = false;
    float* gneg = p.gcat + 2 * p.B * d;  // d dst_negs | d src_negs, [sides][C][N][d]
    if (p.Bc > 0 && p.use_tc) {
        // dA = G . Neg and dNeg = G^T . A in ONE grouped, cost-balanced persistent launch (gemm_tc_group.cu)
        const int64_t a_half = p.sides * p.Bp * d, n_half = p.sides * p.CN * d;
        StageTimer tm(ctx, ST_GEMM_DA, st);
        // A operand of both problems = G, produced from the fp32 scores (or the upstream gradient matrix) inside the kernel
        TcGroupProblem g[2] = {
            {nullptr, nullptr, 0, 0, 0, p.Neg_hl, p.Neg_hl + n_half, d, (int64_t)p.N * d, 1, p.dA, d, p.Bc * d, (int)p.Bc, d, p.N, batches},
            {nullptr, nullptr, 0, 0, 1, p.A_hl, p.A_hl + a_half, d, p.Bc * d, 1, gneg, d, (int64_t)p.N * d, p.N, d, (int)p.Bc, batches}};
        for (auto& q : g) {
            q.conv_src = p.S;
            q.conv_z = ext == nullptr ? p.zw : nullptr;
            q.conv_ld = p.N;
            q.conv_sb = p.Bc * (int64_t)p.N;
            q.conv_rows = (int)p.Bc;
            q.conv_cols = p.N;
            q.conv_mode = ext == nullptr ? 1 : 2;
        }
        MB_TRY(gemm_tc_grouped(g, 2, passes, st));
    } else if (p.Bc > 0) {
        {
            // fp32 FFMA path.  dNeg = G^T . A is independent of dA / edge_backward, so it runs on the second side stream
            cudaStream_t s2 = overlap ? ctx->side2 : st;
            if (overlap) {
                MB_CUDA_TRY(cudaEventRecord(ctx->ev_fork2, st));
                MB_CUDA_TRY(cudaStreamWaitEvent(s2, ctx->ev_fork2, 0));
                dneg_forked = true;
            }
            StageTimer tm(ctx, ST_GEMM_DNEG, s2);
            MB_TRY(gemm_simt(p.S, 1, p.N, p.Bc * p.N, p.A, d, 1, p.Bc * d, gneg, d, (int64_t)p.N * d, p.N, d, (int)p.Bc, batches, s2));
            if (overlap) MB_CUDA_TRY(cudaEventRecord(ctx->ev_join2, s2));
        }
        {
            StageTimer tm(ctx, ST_GEMM_DA, st);  // dA = G . Neg
            MB_TRY(gemm_simt(p.S, p.N, 1, p.Bc * p.N, p.NegE, d, 1, (int64_t)p.N * d, p.dA, d, p.Bc * d, (int)p.Bc, d, p.N, batches, st));
        }
    } else {
        MB_CUDA_TRY(cudaMemsetAsync(gneg, 0, sizeof(float) * 2 * p.CN * d, st));
    }
    {
        StageTimer tm(ctx, ST_EDGE_BWD, st);
        MB_TRY(launch_edge_bwd(p.row_ptrs, emb, emb_ld, row_map, b->edges, p.cols, b->rel, p.sides == 2 ? b->inv_rel : nullptr, p.B, p.Bp, d, b->decoder, p.sides, p.A, p.dA,
                               p.gpos, p.gcat, p.has_rel ? p.drel : nullptr, st));
    }
    // ---- join: the slot / relation plans and the negative-row gradients are needed from here on
    if (overlap) MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_slot, 0));
    if (dneg_forked) MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
    // relation gradients (segmented sum of per-edge gradients by relation id) touch nothing the node update touches: they run on the
    // side stream next to it
    const bool rel_forked = need_rel && overlap;
    if (rel_forked) {
        MB_CUDA_TRY(cudaEventRecord(ctx->ev_fork, st));
        MB_CUDA_TRY(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    }
    if (need_rel) {
        cudaStream_t rs = rel_forked ? ctx->side : st;
        StageTimer tm(ctx, ST_REL_GRAD, rs);
        MB_TRY(launch_rel_reduce(p.drel, p.sides == 2 ? p.drel + p.B * d : nullptr, rel_grad, p.sides == 2 ? inv_rel_grad : nullptr, rvals, p.roffsets, p.R,
                                 d, rs));
        if (rel_forked) MB_CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->side));
    }
    {
        // node gradients: segmented sum over sorted slots (+ Adagrad)
        StageTimer tm(ctx, ST_UPDATE, st);
        if (fused) {
            // barrier 2 (sharded): every rank has fetched the rows it needs -- from here on the tables may change
            if (p.sharded) MB_TRY(launch_shard_barrier(sh, st));
            if (p.sharded && overlap && ctx->side2 != nullptr) {
                // the gradient rows of remote rows cross NVLink (link-bound) on the second side stream while the rows this rank owns
                // are updated (HBM-bound) on the main stream
                MB_CUDA_TRY(cudaEventRecord(ctx->ev_fork2, st));
                MB_CUDA_TRY(cudaStreamWaitEvent(ctx->side2, ctx->ev_fork2, 0));
                MB_TRY(launch_seg_reduce(sh, 2, p.gcat, svals, p.offsets, p.U, d, nullptr, 0, nullptr, 0, nullptr, nullptr, table, state_table, ld, unique_ids, lr,
                                         ctx->side2, p.bounds, 2));
                MB_CUDA_TRY(cudaEventRecord(ctx->ev_join2, ctx->side2));
                MB_TRY(launch_seg_reduce(sh, 2, p.gcat, svals, p.offsets, p.U, d, nullptr, 0, nullptr, 0, nullptr, nullptr, table, state_table, ld, unique_ids, lr, st,
                                         p.bounds, 1));
                MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
            } else {
                MB_TRY(launch_seg_reduce(sh, 2, p.gcat, svals, p.offsets, p.U, d, nullptr, 0, nullptr, 0, nullptr, nullptr, table, state_table, ld, unique_ids, lr, st,
                                         p.sharded ? p.bounds : nullptr));
            }
        } else if (delta_e != nullptr || delta_s != nullptr) {
            MB_REQUIRE(state != nullptr && delta_e != nullptr && delta_s != nullptr, "delta_e/delta_s need state and both outputs");
            MB_TRY(launch_seg_reduce(nullptr, 1, p.gcat, svals, p.offsets, p.U, d, grad, d, state, state_ld, delta_e, delta_s, nullptr, nullptr, 0, nullptr, lr, st));
        } else if (grad != nullptr) {
            MB_TRY(launch_seg_reduce(nullptr, 0, p.gcat, svals, p.offsets, p.U, d, grad, d, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, lr, st));
        }
    }
    if (p.sharded) {
        // barrier 3: every sender's gradient rows are in the owners' inboxes; each owner applies them sender by sender in rank order,
        // every sender's rows as one sparse-Adagrad step on the owner's own HBM (one-process tests: this call serves every owner)
        StageTimer tm(ctx, ST_EXCHANGE, st);
        MB_TRY(launch_shard_barrier(sh, st));
        for (int i = 0; i < sh->world; i++) {
            if (i == sh->rank) continue;
            if (sh->single_process)
                MB_TRY(launch_inbox_apply(sh, i, sh->rank, ld, d, lr, p.U, st));
            else
                MB_TRY(launch_inbox_apply(sh, sh->rank, i, ld, d, lr, sh->exchange_rows, st));
        }
    }
    if (rel_forked) MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" {

const char* mb_last_error(void) { return g_error.c_str(); }
int mb_version(void) { return MB_VERSION; }
uint64_t mb_launch_count(void) { return g_launches.load(); }
const char* mb_build_info(void) { return "sm_100a tcgen05+TMA (bf16x3 split GEMM), fp32 SIMT fallback"; }

mb_status mb_create(int device, mb_context** out) {
    MB_REQUIRE(out != nullptr, "out is null");
    int count = 0;
    MB_CUDA_TRY(cudaGetDeviceCount(&count));
    if (count <= 0) {
        set_error("no CUDA device: the marius_b200 hot path has no CPU fallback");
        return MB_ERR_CUDA;
    }
    MB_REQUIRE(device >= 0 && device < count, "device out of range");
    MB_CUDA_TRY(cudaSetDevice(device));
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFree(nullptr);  // force runtime initialisation so that its exit hooks are registered before ours
        std::atexit(on_process_exit);
        const char* e = getenv("MB_SEGV_TRACE");
        if (e && atoi(e) != 0) {
            signal(SIGSEGV, segv_trace);
            signal(SIGABRT, segv_trace);
        }
    });
    mb_context* c = new mb_context();
    c->device = device;
    cudaError_t e = cudaMalloc(&c->h_loss, sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_loss_pinned, 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copied[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copied[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_loss[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_loss[1], cudaEventDisableTiming);
    // The index plans are chains of small kernels that must finish before the node update.  They share the GPU with grid-filling row
    // kernels and persistent contractions, so their blocks are given scheduling priority (a low-priority side stream was observed to
    // finish its 60 us of sorts 400 us late, stalling the update).
    int prio_lo = 0, prio_hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_slot, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_sfetch_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_sfetch_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_efetch_join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete c;
        set_error(std::string("context creation failed: ") + cudaGetErrorString(e));
        return MB_ERR_CUDA;
    }
    *out = c;
    return MB_OK;
}

void mb_destroy(mb_context* ctx) {
    if (!ctx) return;
    if (g_exiting.load()) {  // see on_process_exit
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    for (int i = 0; i < 2; i++) {
        if (ctx->pf[i]) cudaFree(ctx->pf[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    if (ctx->h_loss) cudaFree(ctx->h_loss);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->h_loss_pinned) cudaFreeHost(ctx->h_loss_pinned);
    for (auto& ev : ctx->ev_loss)
        if (ev) cudaEventDestroy(ev);
    drop_graph(ctx);
    for (void* m : ctx->ipc_mappings) cudaIpcCloseMemHandle(m);
    if (ctx->raw) cudaFree(ctx->raw);
    if (ctx->g_uniq) cudaFree(ctx->g_uniq);
    if (ctx->g_edges) cudaFree(ctx->g_edges);
    if (ctx->g_dneg) cudaFree(ctx->g_dneg);
    if (ctx->g_sneg) cudaFree(ctx->g_sneg);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->side2) cudaStreamDestroy(ctx->side2);
    if (ctx->gstream) cudaStreamDestroy(ctx->gstream);
    if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
    if (ctx->ev_out) cudaEventDestroy(ctx->ev_out);
    if (ctx->ev_fork2) cudaEventDestroy(ctx->ev_fork2);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_slot) cudaEventDestroy(ctx->ev_slot);
    if (ctx->ev_sfetch_fork) cudaEventDestroy(ctx->ev_sfetch_fork);
    if (ctx->ev_sfetch_join) cudaEventDestroy(ctx->ev_sfetch_join);
    if (ctx->ev_efetch_join) cudaEventDestroy(ctx->ev_efetch_join);
    for (auto& sp : ctx->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    delete ctx;
}

size_t mb_workspace_bytes(const mb_context* ctx) { return ctx ? ctx->ws_bytes : 0; }

mb_status mb_enable_peer_access(mb_context* ctx, int peer_device) {
    MB_REQUIRE(ctx != nullptr && peer_device >= 0, "bad arguments");
    if (peer_device == ctx->device) return MB_OK;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    int can = 0;
    MB_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) {
        set_error("device " + std::to_string(ctx->device) + " cannot access peer " + std::to_string(peer_device));
        return MB_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return MB_OK;
    }
    MB_CUDA_TRY(e);
    return MB_OK;
}

// ---- CUDA IPC for the peer-sharded table -------------------------------------------------------------------------------------
// The importer must open the handle with ITS OWN device current (the mapping is created for the current device; memory imported
// under another device's context is not reachable through cudaDeviceEnablePeerAccess), so this is done here rather than through
// torch's tensor sharing, which opens handles under the exporting device's index.
mb_status mb_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
    MB_REQUIRE(dev_ptr != nullptr && handle_out != nullptr && offset_out != nullptr, "null argument");
    typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
    static GetRangeFn get_range = [] {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fp = nullptr;
        return reinterpret_cast<GetRangeFn>(fp);
    }();
    unsigned long long base = (unsigned long long)(uintptr_t)dev_ptr;
    size_t size = 0;
    if (get_range != nullptr) {
        unsigned long long b = 0;
        if (get_range(&b, &size, (unsigned long long)(uintptr_t)dev_ptr) == 0 && b != 0) base = b;
    }
    cudaIpcMemHandle_t h;
    MB_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(handle_out, &h, sizeof(h));
    *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return MB_OK;
}

mb_status mb_ipc_import(mb_context* ctx, const void* handle, int64_t offset, void** ptr_out) {
    MB_REQUIRE(ctx != nullptr && handle != nullptr && ptr_out != nullptr && offset >= 0, "bad arguments");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    MB_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->ipc_mappings.push_back(base);
    *ptr_out = static_cast<char*>(base) + offset;
    return MB_OK;
}

int64_t mb_shard_exchange_bytes(int world, int64_t exchange_rows, int64_t d) {
    if (world < 1 || world > 8 || exchange_rows < 0 || d <= 0) return -1;
    return shard_exchange_bytes(world, exchange_rows, d);
}

mb_status mb_shard_error(mb_context* ctx, const mb_shards* shards, int* error_out) {
    MB_REQUIRE(ctx != nullptr && shards != nullptr && error_out != nullptr, "null argument");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    *error_out = 0;
    if (shards->world <= 1 || shards->exchange[shards->rank] == nullptr) return MB_OK;
    return shard_error_flag(shards, error_out);
}

mb_status mb_graph_enable(mb_context* ctx, int on) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    ctx->graphs_enabled = on ? 1 : 0;
    return MB_OK;
}

mb_status mb_profile_enable(mb_context* ctx, int on) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    ctx->profiling = on == 2 ? 2 : (on != 0 ? 1 : 0);
    return MB_OK;
}

int mb_profile_num_stages(void) { return ST_COUNT; }
const char* mb_profile_stage_name(int stage) { return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : ""; }

mb_status mb_profile_read(mb_context* ctx, float* total_ms, int* counts) {
    MB_REQUIRE(ctx != nullptr && total_ms != nullptr && counts != nullptr, "bad arguments");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    MB_CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < ST_COUNT; i++) {
        total_ms[i] = 0.f;
        counts[i] = 0;
    }
    for (auto& sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            total_ms[sp.stage] += ms;
            counts[sp.stage] += 1;
        }
        ctx->event_pool.push_back(sp.a);
        ctx->event_pool.push_back(sp.b);
    }
    ctx->spans.clear();
    return MB_OK;
}

mb_status mb_profile_timeline(mb_context* ctx, int cap, int* stages, float* start_ms, float* end_ms, int* n) {
    MB_REQUIRE(ctx != nullptr && stages != nullptr && start_ms != nullptr && end_ms != nullptr && n != nullptr && cap >= 0, "bad arguments");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    MB_CUDA_TRY(cudaDeviceSynchronize());
    *n = 0;
    if (ctx->spans.empty()) return MB_OK;
    // time origin: the earliest start event (elapsed time is signed, so any span's start serves as a provisional origin)
    cudaEvent_t origin = ctx->spans[0].a;
    float min_off = 0.f;
    for (auto& sp : ctx->spans) {
        float off = 0.f;
        if (cudaEventElapsedTime(&off, origin, sp.a) == cudaSuccess && off < min_off) min_off = off;
    }
    for (auto& sp : ctx->spans) {
        float a = 0.f, b = 0.f;
        if (*n < cap && cudaEventElapsedTime(&a, origin, sp.a) == cudaSuccess && cudaEventElapsedTime(&b, origin, sp.b) == cudaSuccess) {
            stages[*n] = sp.stage;
            start_ms[*n] = a - min_off;
            end_ms[*n] = b - min_off;
            (*n)++;
        }
        ctx->event_pool.push_back(sp.a);
        ctx->event_pool.push_back(sp.b);
    }
    ctx->spans.clear();
    return MB_OK;
}

mb_status mb_gather_rows(const float* table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n, float* out, int64_t out_ld,
                         void* stream) {
    MB_REQUIRE(n >= 0 && d >= 0 && num_rows >= 0, "negative size");
    MB_REQUIRE(n == 0 || (table && idx && out), "null pointer");
    MB_REQUIRE(ld >= d && out_ld >= d, "leading dimension smaller than row length");
    return gather_rows(table, ld, d, idx, n, out, out_ld, (cudaStream_t)stream);
}

mb_status mb_scatter_add_rows(float* table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* vals, int64_t vals_ld,
                              void* stream) {
    MB_REQUIRE(n >= 0 && d >= 0 && num_rows >= 0, "negative size");
    MB_REQUIRE(n == 0 || (table && idx && vals), "null pointer");  // storage.cpp:652-655 (!values.defined())
    MB_REQUIRE(ld >= d && vals_ld >= d, "leading dimension smaller than row length");
    return scatter_rows(table, ld, d, idx, n, vals, vals_ld, true, (cudaStream_t)stream);
}

mb_status mb_scatter_put_rows(float* table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* vals, int64_t vals_ld,
                              void* stream) {
    MB_REQUIRE(n >= 0 && d >= 0 && num_rows >= 0, "negative size");
    MB_REQUIRE(n == 0 || (table && idx && vals), "null pointer");
    MB_REQUIRE(ld >= d && vals_ld >= d, "leading dimension smaller than row length");
    return scatter_rows(table, ld, d, idx, n, vals, vals_ld, false, (cudaStream_t)stream);
}

mb_status mb_global_to_local_map(int64_t* map, int64_t total_rows, int64_t partition_size, const int32_t* partition_ids, const int32_t* buffer_slots,
                                 int n_resident, void* stream) {
    MB_REQUIRE(map != nullptr && total_rows >= 0 && partition_size > 0 && n_resident >= 0, "bad arguments");
    MB_REQUIRE(n_resident == 0 || (partition_ids && buffer_slots), "null pointer");
    return global_to_local_map(map, total_rows, partition_size, partition_ids, buffer_slots, n_resident, (cudaStream_t)stream);
}

mb_status mb_adagrad_deltas(const float* grad, const float* state, int64_t n, int64_t d, int64_t ld, float lr, float* delta_e, float* delta_s,
                            void* stream) {
    MB_REQUIRE(n >= 0 && d >= 0 && ld >= d, "bad sizes");
    MB_REQUIRE(n == 0 || (grad && state && delta_e && delta_s), "null pointer");
    return adagrad_deltas(grad, state, n, d, ld, lr, delta_e, delta_s, (cudaStream_t)stream);
}

mb_status mb_adagrad_update_rows(float* table, float* state_table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n,
                                 const float* grad, int64_t grad_ld, float lr, void* stream) {
    MB_REQUIRE(n >= 0 && d >= 0 && num_rows >= 0 && ld >= d && grad_ld >= d, "bad sizes");
    MB_REQUIRE(n == 0 || (table && state_table && idx && grad), "null pointer");
    return adagrad_update_rows(table, state_table, ld, d, idx, n, grad, grad_ld, lr, (cudaStream_t)stream);
}

mb_status mb_dense_adagrad_step(float* param, float* state_sum, const float* grad, int64_t n, float lr, float eps, void* stream) {
    MB_REQUIRE(n >= 0 && (n == 0 || (param && state_sum && grad)), "bad arguments");
    return dense_adagrad_step(param, state_sum, grad, n, lr, eps, (cudaStream_t)stream);
}

mb_status mb_map_tensors(mb_context* ctx, const int64_t* all_ids, int64_t n, int64_t max_id, int64_t* unique_out, int64_t* mapped_out,
                         int64_t* num_unique_dev, void* stream) {
    MB_REQUIRE(ctx != nullptr && n >= 0 && max_id >= 0, "bad arguments");
    MB_REQUIRE(n == 0 || (all_ids && unique_out && mapped_out), "null pointer");
    MB_REQUIRE(num_unique_dev != nullptr, "num_unique_dev is null");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    Arena sizing(nullptr);
    auto lay = [&](Arena& ar, uint64_t*& ka, uint64_t*& kb, uint32_t*& va, uint32_t*& vb, uint32_t*& flags, uint32_t*& hist, uint32_t*& total) {
        ka = ar.take<uint64_t>(n);
        kb = ar.take<uint64_t>(n);
        va = ar.take<uint32_t>(n);
        vb = ar.take<uint32_t>(n);
        flags = ar.take<uint32_t>(n + 1);
        hist = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(n)));
        total = ar.take<uint32_t>(1);
    };
    uint64_t *ka, *kb;
    uint32_t *va, *vb, *flags, *hist, *total;
    lay(sizing, ka, kb, va, vb, flags, hist, total);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    lay(place, ka, kb, va, vb, flags, hist, total);
    return map_tensors_device(all_ids, n, bits_for((uint64_t)max_id), ka, kb, va, vb, flags, hist, total, unique_out, mapped_out, num_unique_dev, st);
}

mb_status mb_sample_negatives(int64_t num_nodes, int C, int N, float degree_fraction, const int64_t* edges, int64_t B, int edge_cols, int inverse,
                              uint64_t seed, uint32_t batch_index, int64_t* out, void* stream) {
    MB_REQUIRE(num_nodes > 0 && C >= 0 && N >= 0, "bad dimensions");
    MB_REQUIRE(degree_fraction >= 0.f && degree_fraction <= 1.f, "degree_fraction must be in [0, 1]");
    MB_REQUIRE((int64_t)C * N == 0 || out != nullptr, "out is null");
    const int num_batch = (int)(N * degree_fraction);  // negative.cpp:334
    if (num_batch > 0) {
        MB_REQUIRE(edges != nullptr && B > 0, "degree-based negatives need the batch's edges");
        MB_REQUIRE(edge_cols == 2 || edge_cols == 3, "Edge list must be a 3 or 2 column tensor");
    }
    return launch_sample_negatives(num_nodes, C, N, num_batch, edges, B, edge_cols, inverse != 0, seed, batch_index, out, (cudaStream_t)stream);
}

mb_status mb_edge_sample(mb_context* ctx, const int64_t* edges, int64_t B, int edge_cols, const int64_t* src_negs, const int64_t* dst_negs, int C, int N,
                         int64_t max_id, int64_t* unique_out, int64_t* edges_local, int64_t* src_negs_local, int64_t* dst_negs_local,
                         int64_t* num_unique_dev, void* stream) {
    MB_REQUIRE(ctx != nullptr && B >= 0 && C >= 0 && N >= 0 && max_id >= 0, "bad arguments");
    MB_REQUIRE(edge_cols == 2 || edge_cols == 3, "Edge list must be a 3 or 2 column tensor");  // dataloader.cpp:463-469
    const int64_t CN = (int64_t)C * N, n_s = src_negs ? CN : 0, n = 2 * B + n_s + CN;
    MB_REQUIRE(num_unique_dev != nullptr, "num_unique_dev is null");
    MB_REQUIRE(B == 0 || (edges && edges_local), "null edges");
    MB_REQUIRE(CN == 0 || (dst_negs && dst_negs_local), "null dst_negs");
    MB_REQUIRE(n_s == 0 || src_negs_local != nullptr, "null src_negs_local");
    MB_REQUIRE(n == 0 || unique_out != nullptr, "null unique_out");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t *ka, *kb;
    uint32_t *va, *vb, *flags, *hist, *total;
    int64_t *all_ids, *mapped;
    auto lay = [&](Arena& ar) {
        all_ids = ar.take<int64_t>(n);
        mapped = ar.take<int64_t>(n);
        ka = ar.take<uint64_t>(n);
        kb = ar.take<uint64_t>(n);
        va = ar.take<uint32_t>(n);
        vb = ar.take<uint32_t>(n);
        flags = ar.take<uint32_t>(n + 1);
        hist = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(n)));
        total = ar.take<uint32_t>(1);
    };
    Arena sizing(nullptr);
    lay(sizing);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    lay(place);
    MB_TRY(launch_concat_ids(edges, B, edge_cols, src_negs, dst_negs, CN, all_ids, st));
    MB_TRY(map_tensors_device(all_ids, n, bits_for((uint64_t)max_id), ka, kb, va, vb, flags, hist, total, unique_out, mapped, num_unique_dev, st));
    MB_TRY(launch_split_mapped(mapped, edges, B, edge_cols, src_negs != nullptr, CN, edges_local, src_negs_local, dst_negs_local, st));
    return MB_OK;
}

mb_status mb_reduce_rows_by_key(mb_context* ctx, const int64_t* ids, const float* rows, int64_t n, int64_t d, int64_t max_id, int64_t* unique_out,
                                float* rows_out, int64_t* num_unique_dev, void* stream) {
    MB_REQUIRE(ctx != nullptr && n >= 0 && d > 0 && max_id >= 0, "bad arguments");
    MB_REQUIRE(n == 0 || (ids && rows && unique_out && rows_out), "null pointer");
    MB_REQUIRE(num_unique_dev != nullptr, "num_unique_dev is null");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (n == 0) {
        MB_CUDA_TRY(cudaMemsetAsync(num_unique_dev, 0, sizeof(int64_t), st));
        return MB_OK;
    }
    uint64_t *ka, *kb;
    uint32_t *va, *vb, *flags, *hist, *total;
    int64_t* mapped;
    uint32_t *k32a, *k32b, *v32a, *v32b, *offsets, *hist2;
    auto lay = [&](Arena& ar) {
        ka = ar.take<uint64_t>(n);
        kb = ar.take<uint64_t>(n);
        va = ar.take<uint32_t>(n);
        vb = ar.take<uint32_t>(n);
        flags = ar.take<uint32_t>(n + 1);
        hist = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(n)));
        total = ar.take<uint32_t>(1);
        mapped = ar.take<int64_t>(n);
        k32a = ar.take<uint32_t>(n);
        k32b = ar.take<uint32_t>(n);
        v32a = ar.take<uint32_t>(n);
        v32b = ar.take<uint32_t>(n);
        offsets = ar.take<uint32_t>(n + 2);
        hist2 = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(n)));
    };
    Arena sizing(nullptr);
    lay(sizing);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    lay(place);
    // (1) unique ids + position of every input row in the unique list (== map_tensors)
    MB_TRY(map_tensors_device(ids, n, bits_for((uint64_t)max_id), ka, kb, va, vb, flags, hist, total, unique_out, mapped, num_unique_dev, st));
    // (2) slots sorted by unique position -> segmented sum in slot order (deterministic)
    MB_TRY(launch_i64_to_u32(mapped, k32a, n, st));
    uint32_t *sk = nullptr, *sv = nullptr;
    MB_TRY(radix_sort_pairs<uint32_t>(k32a, k32b, v32a, v32b, n, bits_for((uint64_t)n), hist2, &sk, &sv, st));
    MB_TRY(segment_offsets_u32(sk, n, n, offsets, st));  // segments beyond num_unique are empty and write zeros into unused output rows
    return launch_seg_reduce(nullptr, 0, rows, sv, offsets, n, (int)d, rows_out, d, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0.f, st);
}

mb_status mb_decoder_forward(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, int precision, float* pos, float* neg,
                             float* inv_pos, float* inv_neg, void* stream) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    MB_TRY(validate_batch(batch));
    MB_REQUIRE(emb != nullptr && pos != nullptr && neg != nullptr, "UndefinedTensor");  // comparators.cpp:63-65
    MB_REQUIRE(emb_ld >= batch->d, "emb_ld < d");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    Plan p;
    fill_plan_dims(p, batch, precision);
    if (p.sides == 2) MB_REQUIRE(inv_pos != nullptr && inv_neg != nullptr, "inverse outputs required when inverse relations are used");
    {
        Arena sizing(nullptr);
        p.layout(sizing, false, false, false);
        MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
        Arena place(ctx->ws);
        p.layout(place, false, false, false);
    }
    // pos / inv_pos are separate user buffers: stage them in the workspace layout [sides][Bp] and copy out
    MB_TRY(run_forward(ctx, p, batch, emb, emb_ld, nullptr, precision, p.pos, neg, inv_neg, false, st));
    MB_CUDA_TRY(cudaMemcpyAsync(pos, p.pos, sizeof(float) * p.Bp, cudaMemcpyDeviceToDevice, st));
    if (p.sides == 2) MB_CUDA_TRY(cudaMemcpyAsync(inv_pos, p.pos + p.Bp, sizeof(float) * p.Bp, cudaMemcpyDeviceToDevice, st));
    return MB_OK;
}

static mb_status filter_checked(float* scores, int64_t rows, int64_t N, int64_t ld, const int64_t* filter, int64_t F, cudaStream_t st) {
    if (F == 0) return MB_OK;
    MB_REQUIRE(filter != nullptr && scores != nullptr, "null filter / scores");
    int* flag = nullptr;
    MB_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&flag), sizeof(int), st));
    MB_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
    mb_status rs = launch_score_filter(scores, rows, N, ld, filter, F, flag, st);
    int host_flag = 0;
    cudaError_t e = cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(flag, st);
    MB_TRY(rs);
    MB_CUDA_TRY(e);
    MB_REQUIRE(host_flag == 0, "score filter index out of range");
    return MB_OK;
}

mb_status mb_apply_score_filter(float* scores, int64_t rows, int64_t N, int64_t ld, const int64_t* filter, int64_t F, void* stream) {
    MB_REQUIRE(rows >= 0 && N >= 0 && F >= 0 && ld >= N, "bad dimensions");
    return filter_checked(scores, rows, N, ld, filter, F, (cudaStream_t)stream);
}

mb_status mb_compute_ranks(const float* pos, const float* neg, int64_t rows, int64_t N, int64_t ld, int64_t* ranks, void* stream) {
    MB_REQUIRE(rows >= 0 && N >= 0 && ld >= N, "bad dimensions");
    MB_REQUIRE(rows == 0 || (pos != nullptr && neg != nullptr && ranks != nullptr), "UndefinedTensor");
    return launch_ranks(pos, neg, rows, N, ld, ranks, (cudaStream_t)stream);
}

mb_status mb_evaluate_batch(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, int precision, const int64_t* dst_filter,
                            int64_t Fd, const int64_t* src_filter, int64_t Fs, int64_t* ranks, int64_t* inv_ranks, float* pos, float* inv_pos,
                            void* stream) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    MB_TRY(validate_batch(batch));
    MB_REQUIRE(emb != nullptr && ranks != nullptr, "UndefinedTensor");
    MB_REQUIRE(emb_ld >= batch->d, "emb_ld < d");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    MB_REQUIRE(Fd >= 0 && Fs >= 0, "negative filter size");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    Plan p;
    fill_plan_dims(p, batch, precision);
    if (p.sides == 2) MB_REQUIRE(inv_ranks != nullptr, "inverse outputs required when inverse relations are used");
    {
        Arena sizing(nullptr);
        p.layout(sizing, false, false, true);
        MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
        Arena place(ctx->ws);
        p.layout(place, false, false, true);
    }
    float* S1 = p.S + p.Bp * p.N;
    MB_TRY(run_forward(ctx, p, batch, emb, emb_ld, nullptr, precision, p.pos, p.S, S1, true, st));
    if (p.Bc == 0) {  // no positives: nothing to rank
        return MB_OK;
    }
    // both filters report out-of-range indices through the context's persistent flag: one read-back per call, no allocation
    const bool any_filter = Fd > 0 || (p.sides == 2 && Fs > 0);
    if (any_filter) MB_CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st));
    if (Fd > 0) {
        MB_REQUIRE(dst_filter != nullptr, "null filter");
        MB_TRY(launch_score_filter(p.S, p.Bp, p.N, p.N, dst_filter, Fd, ctx->d_flag, st));
    }
    MB_TRY(launch_ranks(p.pos, p.S, p.Bp, p.N, p.N, ranks, st));
    if (p.sides == 2) {
        if (Fs > 0) {
            MB_REQUIRE(src_filter != nullptr, "null filter");
            MB_TRY(launch_score_filter(S1, p.Bp, p.N, p.N, src_filter, Fs, ctx->d_flag, st));
        }
        MB_TRY(launch_ranks(p.pos + p.Bp, S1, p.Bp, p.N, p.N, inv_ranks, st));
    }
    if (any_filter) {
        int host_flag = 0;
        MB_CUDA_TRY(cudaMemcpyAsync(&host_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        MB_CUDA_TRY(cudaStreamSynchronize(st));
        MB_REQUIRE(host_flag == 0, "score filter index out of range");
    }
    if (pos) MB_CUDA_TRY(cudaMemcpyAsync(pos, p.pos, sizeof(float) * p.Bp, cudaMemcpyDeviceToDevice, st));
    if (inv_pos && p.sides == 2) MB_CUDA_TRY(cudaMemcpyAsync(inv_pos, p.pos + p.Bp, sizeof(float) * p.Bp, cudaMemcpyDeviceToDevice, st));
    return MB_OK;
}

// ---- score-filter construction + streaming all-node evaluation ------------------------------------------------
mb_status mb_filter_sort_edges(mb_context* ctx, const int64_t* graph_edges, int64_t E, int edge_cols, int inverse, int64_t max_id, int64_t* sorted_out,
                               void* stream) {
    MB_REQUIRE(ctx != nullptr && E >= 0 && max_id >= 0, "bad arguments");
    MB_REQUIRE(edge_cols == 2 || edge_cols == 3, "Edge list must be a 3 or 2 column tensor");
    MB_REQUIRE(E == 0 || (graph_edges != nullptr && sorted_out != nullptr), "null pointer");
    MB_REQUIRE(E < ((int64_t)1 << 32), "too many edges for 32-bit sort payloads");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t *ka, *kb;
    uint32_t *va, *vb, *hist;
    auto lay = [&](Arena& ar) {
        ka = ar.take<uint64_t>(E);
        kb = ar.take<uint64_t>(E);
        va = ar.take<uint32_t>(E);
        vb = ar.take<uint32_t>(E);
        hist = reinterpret_cast<uint32_t*>(ar.take<char>(sort_scratch_bytes(E)));
    };
    Arena sizing(nullptr);
    lay(sizing);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    lay(place);
    // the kept endpoint: the source for destination corruption, the destination for source corruption (negative.cpp:70-78)
    const int key_col = inverse ? edge_cols - 1 : 0;
    return launch_filter_sort(graph_edges, E, edge_cols, key_col, ka, kb, va, vb, hist, bits_for((uint64_t)max_id), sorted_out, st);
}

mb_status mb_compute_filter(mb_context* ctx, const int64_t* sorted_edges, int64_t E, int edge_cols, int inverse, const int64_t* batch_edges, int64_t B,
                            int64_t* filter_out, int64_t cap, int64_t* count_dev, void* stream) {
    MB_REQUIRE(ctx != nullptr && E >= 0 && B >= 0 && cap >= 0 && count_dev != nullptr, "bad arguments");
    MB_REQUIRE(edge_cols == 2 || edge_cols == 3, "Edge list must be a 3 or 2 column tensor");
    MB_REQUIRE((E == 0 || sorted_edges != nullptr) && (B == 0 || batch_edges != nullptr) && (cap == 0 || filter_out != nullptr), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    int64_t *counts, *offsets;
    auto lay = [&](Arena& ar) {
        counts = ar.take<int64_t>(B + 1);
        offsets = ar.take<int64_t>(B + 1);
    };
    Arena sizing(nullptr);
    lay(sizing);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    lay(place);
    const int key_col = inverse ? edge_cols - 1 : 0, cor_col = inverse ? 0 : edge_cols - 1;
    return launch_filter_match(sorted_edges, E, edge_cols, key_col, cor_col, batch_edges, B, counts, offsets, filter_out, cap, count_dev, st);
}

mb_status mb_evaluate_all_nodes(mb_context* ctx, int decoder, const float* table, int64_t num_nodes, int64_t ld, int64_t d, const int64_t* edges, int64_t B,
                                int edge_cols, const float* rel, const float* inv_rel, int64_t R, const int64_t* dst_filter, int64_t Fd,
                                const int64_t* src_filter, int64_t Fs, int precision, int64_t tile_rows, int64_t* ranks, int64_t* inv_ranks, float* pos_out,
                                float* inv_pos_out, void* stream) {
    MB_REQUIRE(ctx != nullptr && table != nullptr && ranks != nullptr, "UndefinedTensor");
    MB_REQUIRE(num_nodes > 0 && d > 0 && ld == d && B >= 0 && Fd >= 0 && Fs >= 0, "bad dimensions (the table must be dense: ld == d)");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    if (B == 0) return MB_OK;
    // the positives as one chunk (C = 1); the "negatives" are the table itself, tile by tile
    mb_batch b;
    std::memset(&b, 0, sizeof(b));
    b.decoder = decoder;
    b.U = num_nodes;
    b.d = d;
    b.B = B;
    b.R = R;
    b.C = 1;
    b.N = 8;
    b.edges = edges;
    b.edge_cols = edge_cols;
    b.dst_negs = edges;  // (placeholder: no negative rows are gathered, CN = 0 below)
    b.src_negs = inv_rel != nullptr ? edges : nullptr;
    b.rel = rel;
    b.inv_rel = inv_rel;
    MB_TRY(validate_batch(&b));
    Plan p;
    fill_plan_dims(p, &b, precision);
    const int sides = p.sides;
    if (sides == 2) MB_REQUIRE(inv_ranks != nullptr, "inverse outputs required when inverse relations are used");
    MB_REQUIRE(p.use_tc || precision == MB_PREC_FP32, "all-node evaluation needs d % 8 == 0 for the tensor-core path");
    int64_t T = tile_rows > 0 ? tile_rows : 32768;
    T = std::min<int64_t>((T + 7) / 8 * 8, (num_nodes + 7) / 8 * 8);
    float *A, *pos, *S, *NegF;
    __nv_bfloat16 *A_hl, *Neg_hl;
    auto lay = [&](Arena& ar) {
        A = ar.take<float>(sides * B * d);
        pos = ar.take<float>(sides * B);
        A_hl = p.use_tc ? ar.take<__nv_bfloat16>(2 * sides * B * d) : nullptr;
        Neg_hl = p.use_tc ? ar.take<__nv_bfloat16>(2 * T * d) : nullptr;
        NegF = nullptr;
        S = ar.take<float>(sides * B * T);
    };
    Arena sizing(nullptr);
    lay(sizing);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    lay(place);
    const int64_t a_half = sides * B * d;
    const bool need_A = !p.use_tc || !decoder_vec_ok(table, ld, (int)d, p.has_rel, rel, sides == 2 ? inv_rel : nullptr, sides);
    MB_TRY(launch_prep(nullptr, nullptr, nullptr, table, ld, nullptr, edges, edge_cols, rel, sides == 2 ? inv_rel : nullptr, B, B, 0, (int)d, decoder, sides, nullptr,
                       nullptr, need_A ? A : nullptr, pos, p.use_tc ? (void*)A_hl : nullptr, p.use_tc ? (void*)(A_hl + a_half) : nullptr, nullptr, nullptr,
                       nullptr, st));
    MB_TRY(launch_fill_i64(ranks, B, 1, st));
    if (sides == 2) MB_TRY(launch_fill_i64(inv_ranks, B, 1, st));
    const bool any_filter = Fd > 0 || (sides == 2 && Fs > 0);
    if (any_filter) MB_CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st));
    const int passes = precision == MB_PREC_BF16 ? 1 : 3;
    for (int64_t t0 = 0; t0 < num_nodes; t0 += T) {
        const int64_t Tc = std::min<int64_t>(T, num_nodes - t0);
        if (p.use_tc) {
            MB_TRY(launch_split(table + t0 * ld, Tc * d, Neg_hl, Neg_hl + T * d, st));
            // S[side] = A[side] . tile^T : both sides score against the same tile (one launch per side, or both sides grouped in one)
            TcGroupProblem g[2];
            for (int sd = 0; sd < sides; sd++)
                g[sd] = TcGroupProblem{A_hl + sd * B * d, A_hl + a_half + sd * B * d, d, B * d, 0, Neg_hl, Neg_hl + T * d, d, T * d, 0, S + sd * B * T, T, B * T,
                                       (int)B, (int)Tc, (int)d, 1};
            MB_TRY(gemm_tc_grouped(g, sides, passes, st));
        } else {
            MB_TRY(gemm_simt(A, d, 1, B * d, table + t0 * ld, 1, d, 0, S, T, B * T, (int)B, (int)Tc, (int)d, sides, st));
        }
        if (Fd > 0) MB_TRY(launch_filter_tile(S, B, T, t0, Tc, dst_filter, Fd, num_nodes, ctx->d_flag, st));
        MB_TRY(launch_rank_accumulate(pos, S, B, Tc, T, ranks, st));
        if (sides == 2) {
            if (Fs > 0) MB_TRY(launch_filter_tile(S + B * T, B, T, t0, Tc, src_filter, Fs, num_nodes, ctx->d_flag, st));
            MB_TRY(launch_rank_accumulate(pos + B, S + B * T, B, Tc, T, inv_ranks, st));
        }
    }
    if (pos_out) MB_CUDA_TRY(cudaMemcpyAsync(pos_out, pos, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
    if (inv_pos_out && sides == 2) MB_CUDA_TRY(cudaMemcpyAsync(inv_pos_out, pos + B, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
    if (any_filter) {
        int host_flag = 0;
        MB_CUDA_TRY(cudaMemcpyAsync(&host_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        MB_CUDA_TRY(cudaStreamSynchronize(st));
        MB_REQUIRE(host_flag == 0, "score filter index out of range");
    }
    (void)NegF;
    return MB_OK;
}

mb_status mb_train_batch(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, const float* state, int64_t state_ld, float lr,
                         int reduction, int precision, float* loss, float* grad, float* delta_e, float* delta_s, float* rel_grad, float* inv_rel_grad,
                         void* stream) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    MB_REQUIRE(emb != nullptr, "UndefinedTensor: node embeddings");
    MB_REQUIRE(batch == nullptr || emb_ld >= batch->d, "emb_ld < d");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    return run_train(ctx, batch, emb, emb_ld, state, state_ld, nullptr, nullptr, 0, nullptr, lr, reduction, precision, loss, grad, delta_e, delta_s,
                     rel_grad, inv_rel_grad, UpdateMode::kBatchLocal, (cudaStream_t)stream);
}

mb_status mb_decoder_backward(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, int precision, const float* gpos,
                              const float* gneg, const float* ginv_pos, const float* ginv_neg, float* grad, float* rel_grad, float* inv_rel_grad,
                              void* stream) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    MB_REQUIRE(emb != nullptr && gpos != nullptr && gneg != nullptr && grad != nullptr, "UndefinedTensor");
    MB_REQUIRE(batch == nullptr || emb_ld >= batch->d, "emb_ld < d");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    const float* ext[4] = {gpos, gneg, ginv_pos, ginv_neg};
    return run_train(ctx, batch, emb, emb_ld, nullptr, 0, nullptr, nullptr, 0, nullptr, 0.f, MB_REDUCTION_SUM, precision, nullptr, grad, nullptr, nullptr,
                     rel_grad, inv_rel_grad, UpdateMode::kBatchLocal, (cudaStream_t)stream, ext);
}

// ---- CUDA-graph replay of the fused step ------------------------------------------------------------------------------------------
// The step is ~28 dependent launches on three streams; replaying it as one graph removes the inter-kernel launch gaps and most of
// the per-step host time.  What changes from batch to batch is only (a) the contents of the index tensors and (b) U, the number of
// unique rows.  (a): the graph reads fixed staging buffers, refreshed by four memcpy nodes whose source pointer / size are patched
// before each launch (cudaGraphExecMemcpyNodeSetParams1D).  (b): the graph runs with U = capacity (2B + 2CN >= U); the extra
// segments are empty and the update kernel skips empty segments, so they touch no memory.
static mb_status grow_i64(int64_t** p, size_t* cap, size_t need);

static bool graphs_on(mb_context* ctx) {
    if (ctx->graphs_enabled < 0) {
        const char* e = getenv("MB_GRAPH");
        ctx->graphs_enabled = e ? (atoi(e) != 0) : 1;
    }
    return ctx->graphs_enabled == 1 && !ctx->profiling;
}

static mb_status train_step_any(mb_context* ctx, const mb_batch* ub, bool host_inputs, float* table, float* state_table, int64_t ld,
                                const int64_t* unique_ids, float lr, int reduction, int precision, float* loss_dev, float* loss_host, float* rel_grad,
                                float* inv_rel_grad, cudaStream_t st, const mb_shards* sh = nullptr) {
    MB_TRY(validate_batch(ub));
    const int64_t n_e = ub->B * ub->edge_cols, n_n = (int64_t)ub->C * ub->N, cap_u = 2 * ub->B + 2 * n_n;
    Plan probe;
    fill_plan_dims(probe, ub, precision);
    const bool vec = decoder_vec_ok(table, ld, (int)ub->d, probe.has_rel, ub->rel, probe.sides == 2 ? ub->inv_rel : nullptr, probe.sides);
    const bool eligible = graphs_on(ctx) && vec && probe.use_tc && ub->B > 0 && ub->U <= cap_u;
    if (!eligible && !host_inputs && loss_host == nullptr) {
        return run_train(ctx, ub, nullptr, 0, nullptr, 0, table, state_table, ld, unique_ids, lr, reduction, precision, loss_dev, nullptr, nullptr, nullptr,
                         rel_grad, inv_rel_grad, UpdateMode::kFusedTable, st, nullptr, sh);
    }
    // fixed-address staging of the batch's index tensors
    if ((size_t)cap_u > ctx->g_uniq_cap || (size_t)n_e > ctx->g_edges_cap || (size_t)n_n > ctx->g_dneg_cap || (size_t)n_n > ctx->g_sneg_cap) {
        MB_CUDA_TRY(cudaStreamSynchronize(st));
        drop_graph(ctx);
        MB_TRY(grow_i64(&ctx->g_uniq, &ctx->g_uniq_cap, (size_t)cap_u));
        MB_TRY(grow_i64(&ctx->g_edges, &ctx->g_edges_cap, (size_t)n_e));
        MB_TRY(grow_i64(&ctx->g_dneg, &ctx->g_dneg_cap, (size_t)n_n));
        MB_TRY(grow_i64(&ctx->g_sneg, &ctx->g_sneg_cap, (size_t)n_n));
    }
    const cudaMemcpyKind kind = host_inputs ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const bool has_sneg = ub->src_negs != nullptr;
    float* loss_target = loss_dev ? loss_dev : ctx->h_loss;
    mb_batch db = *ub;
    db.edges = ctx->g_edges;
    db.dst_negs = ctx->g_dneg;
    db.src_negs = has_sneg ? ctx->g_sneg : nullptr;

    auto stage_inputs = [&]() -> mb_status {  // Batch::to (batch.cpp:21-60) when the inputs are host tensors
        MB_CUDA_TRY(cudaMemcpyAsync(ctx->g_uniq, unique_ids, sizeof(int64_t) * ub->U, kind, st));
        MB_CUDA_TRY(cudaMemcpyAsync(ctx->g_edges, ub->edges, sizeof(int64_t) * n_e, kind, st));
        MB_CUDA_TRY(cudaMemcpyAsync(ctx->g_dneg, ub->dst_negs, sizeof(int64_t) * n_n, kind, st));
        if (has_sneg) MB_CUDA_TRY(cudaMemcpyAsync(ctx->g_sneg, ub->src_negs, sizeof(int64_t) * n_n, kind, st));
        return MB_OK;
    };

    if (!eligible) {  // host inputs without graph replay
        MB_TRY(stage_inputs());
        MB_TRY(run_train(ctx, &db, nullptr, 0, nullptr, 0, table, state_table, ld, ctx->g_uniq, lr, reduction, precision, loss_target, nullptr, nullptr,
                         nullptr, rel_grad, inv_rel_grad, UpdateMode::kFusedTable, st, nullptr, sh));
        if (loss_host) MB_CUDA_TRY(cudaMemcpyAsync(loss_host, loss_target, sizeof(float), cudaMemcpyDeviceToHost, st));
        return MB_OK;
    }

    std::vector<uint64_t> key = {(uint64_t)ub->decoder, (uint64_t)ub->d, (uint64_t)ub->B, (uint64_t)ub->R, (uint64_t)ub->C, (uint64_t)ub->N,
                                 (uint64_t)ub->edge_cols, (uint64_t)(uintptr_t)ub->rel, (uint64_t)(uintptr_t)ub->inv_rel, (uint64_t)has_sneg,
                                 (uint64_t)(uintptr_t)table, (uint64_t)(uintptr_t)state_table, (uint64_t)ld, (uint64_t)reduction, (uint64_t)precision,
                                 (uint64_t)(uintptr_t)loss_target, (uint64_t)(uintptr_t)rel_grad, (uint64_t)(uintptr_t)inv_rel_grad, (uint64_t)host_inputs,
                                 (uint64_t)(loss_host != nullptr), (uint64_t)(uintptr_t)st, 0};
    std::memcpy(&key.back(), &lr, sizeof(float));
    if (sh != nullptr) {
        key.push_back((uint64_t)sh->world);
        key.push_back((uint64_t)sh->rank);
        key.push_back((uint64_t)sh->rows_per_rank);
        for (int i = 0; i < sh->world && i < 8; i++) {
            key.push_back((uint64_t)(uintptr_t)sh->tables[i]);
            key.push_back((uint64_t)(uintptr_t)sh->states[i]);
            key.push_back((uint64_t)(uintptr_t)sh->exchange[i]);
        }
        key.push_back((uint64_t)sh->exchange_rows);
        key.push_back((uint64_t)sh->single_process);
    }
    if (ctx->sg.key != key) {
        if (ctx->sg.valid || ctx->sg.exec) {
            MB_CUDA_TRY(cudaStreamSynchronize(st));
            drop_graph(ctx);
        }
        ctx->sg.key = key;
        ctx->sg.warm = 0;
    }
    mb_batch gb = db;
    gb.U = cap_u;  // capacity: see the header comment
    if (!ctx->sg.valid) {
        if (ctx->sg.warm == 0) {
            // first sight of this signature: run eagerly (allocations, cudaFuncSetAttribute, tile tables happen here, outside capture)
            ctx->sg.warm = 1;
            MB_CUDA_TRY(cudaMemsetAsync(ctx->g_uniq, 0xFF, sizeof(int64_t) * cap_u, st));  // ids beyond U = -1 (padding)
            MB_TRY(stage_inputs());
            MB_TRY(run_train(ctx, &gb, nullptr, 0, nullptr, 0, table, state_table, ld, ctx->g_uniq, lr, reduction, precision, loss_target, nullptr,
                             nullptr, nullptr, rel_grad, inv_rel_grad, UpdateMode::kFusedTable, st, nullptr, sh));
            if (loss_host) MB_CUDA_TRY(cudaMemcpyAsync(loss_host, loss_target, sizeof(float), cudaMemcpyDeviceToHost, st));
            return MB_OK;
        }
        // capture (on the context's own stream)
        cudaStream_t gs = ctx->gstream;
        MB_CUDA_TRY(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
        mb_status rs = MB_OK;
        {
            // the graph runs with U = capacity: every replay first resets the id staging buffer to -1, so the entries beyond this batch's
            // U are padding (never a stale id of an earlier batch, which the sharded step would fetch over NVLink for nothing)
            cudaError_t c0 = cudaMemsetAsync(ctx->g_uniq, 0xFF, sizeof(int64_t) * cap_u, gs);
            cudaError_t c1 = cudaMemcpyAsync(ctx->g_uniq, unique_ids, sizeof(int64_t) * ub->U, kind, gs);
            cudaError_t c2 = cudaMemcpyAsync(ctx->g_edges, ub->edges, sizeof(int64_t) * n_e, kind, gs);
            cudaError_t c3 = cudaMemcpyAsync(ctx->g_dneg, ub->dst_negs, sizeof(int64_t) * n_n, kind, gs);
            cudaError_t c4 = has_sneg ? cudaMemcpyAsync(ctx->g_sneg, ub->src_negs, sizeof(int64_t) * n_n, kind, gs) : cudaSuccess;
            if (c0 != cudaSuccess || c1 != cudaSuccess || c2 != cudaSuccess || c3 != cudaSuccess || c4 != cudaSuccess) rs = MB_ERR_CUDA;
        }
        if (rs == MB_OK)
            rs = run_train(ctx, &gb, nullptr, 0, nullptr, 0, table, state_table, ld, ctx->g_uniq, lr, reduction, precision, loss_target, nullptr, nullptr,
                           nullptr, rel_grad, inv_rel_grad, UpdateMode::kFusedTable, gs, nullptr, sh);
        if (rs == MB_OK && loss_host && cudaMemcpyAsync(loss_host, loss_target, sizeof(float), cudaMemcpyDeviceToHost, gs) != cudaSuccess)
            rs = MB_ERR_CUDA;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        if (rs != MB_OK || ce != cudaSuccess || graph == nullptr) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ctx->graphs_enabled = 0;  // fall back to eager launches for the rest of this context's life
            set_error("graph capture of the step failed; continuing with eager launches");
            return train_step_any(ctx, ub, host_inputs, table, state_table, ld, unique_ids, lr, reduction, precision, loss_dev, loss_host, rel_grad,
                                  inv_rel_grad, st, sh);
        }
        cudaGraphExec_t exec = nullptr;
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGraphDestroy(graph);
            cudaGetLastError();
            ctx->graphs_enabled = 0;
            return train_step_any(ctx, ub, host_inputs, table, state_table, ld, unique_ids, lr, reduction, precision, loss_dev, loss_host, rel_grad,
                                  inv_rel_grad, st, sh);
        }
        // locate the input-staging memcpy nodes by their destination
        size_t nn = 0;
        MB_CUDA_TRY(cudaGraphGetNodes(graph, nullptr, &nn));
        std::vector<cudaGraphNode_t> nodes(nn);
        MB_CUDA_TRY(cudaGraphGetNodes(graph, nodes.data(), &nn));
        for (auto nd : nodes) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeMemcpy) continue;
            cudaMemcpy3DParms mp;
            if (cudaGraphMemcpyNodeGetParams(nd, &mp) != cudaSuccess) continue;
            void* dst = mp.dstPtr.ptr;
            if (dst == ctx->g_uniq) ctx->sg.n_uniq = nd;
            else if (dst == ctx->g_edges) ctx->sg.n_edges = nd;
            else if (dst == ctx->g_dneg) ctx->sg.n_dneg = nd;
            else if (dst == ctx->g_sneg) ctx->sg.n_sneg = nd;
            else if (loss_host && dst == loss_host) ctx->sg.n_loss = nd;
        }
        if (!ctx->sg.n_uniq || !ctx->sg.n_edges || !ctx->sg.n_dneg || (has_sneg && !ctx->sg.n_sneg) || (loss_host && !ctx->sg.n_loss)) {
            cudaGraphExecDestroy(exec);
            cudaGraphDestroy(graph);
            ctx->graphs_enabled = 0;
            return train_step_any(ctx, ub, host_inputs, table, state_table, ld, unique_ids, lr, reduction, precision, loss_dev, loss_host, rel_grad,
                                  inv_rel_grad, st, sh);
        }
        ctx->sg.graph = graph;
        ctx->sg.exec = exec;
        ctx->sg.valid = true;
    }
    // patch the four input copies and replay
    MB_CUDA_TRY(cudaGraphExecMemcpyNodeSetParams1D(ctx->sg.exec, ctx->sg.n_uniq, ctx->g_uniq, unique_ids, sizeof(int64_t) * ub->U, kind));
    MB_CUDA_TRY(cudaGraphExecMemcpyNodeSetParams1D(ctx->sg.exec, ctx->sg.n_edges, ctx->g_edges, ub->edges, sizeof(int64_t) * n_e, kind));
    MB_CUDA_TRY(cudaGraphExecMemcpyNodeSetParams1D(ctx->sg.exec, ctx->sg.n_dneg, ctx->g_dneg, ub->dst_negs, sizeof(int64_t) * n_n, kind));
    if (has_sneg) MB_CUDA_TRY(cudaGraphExecMemcpyNodeSetParams1D(ctx->sg.exec, ctx->sg.n_sneg, ctx->g_sneg, ub->src_negs, sizeof(int64_t) * n_n, kind));
    if (loss_host)  // the loss lands in this call's slot
        MB_CUDA_TRY(cudaGraphExecMemcpyNodeSetParams1D(ctx->sg.exec, ctx->sg.n_loss, loss_host, loss_target, sizeof(float), cudaMemcpyDeviceToHost));
    MB_CUDA_TRY(cudaEventRecord(ctx->ev_in, st));
    MB_CUDA_TRY(cudaStreamWaitEvent(ctx->gstream, ctx->ev_in, 0));
    MB_CUDA_TRY(cudaGraphLaunch(ctx->sg.exec, ctx->gstream));
    MB_CUDA_TRY(cudaEventRecord(ctx->ev_out, ctx->gstream));
    MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_out, 0));
    // kernels inside the replayed graph (mb_launch_count stays an honest kernel count); sharded: + fetch, bounds, 3 barriers, world-1 applies
    count_launch(sh != nullptr && sh->world > 1 ? 29 + 3 + (sh->single_process ? 0 : 3) + (sh->world - 1) : 29);
    return MB_OK;
}

mb_status mb_train_step(mb_context* ctx, const mb_batch* batch, float* table, float* state_table, int64_t num_rows, int64_t ld,
                        const int64_t* unique_ids, float lr, int reduction, int precision, float* loss, float* rel_grad, float* inv_rel_grad,
                        void* stream) {
    MB_REQUIRE(ctx != nullptr, "context is null");
    MB_REQUIRE(table != nullptr && state_table != nullptr && unique_ids != nullptr, "null table / ids");
    MB_REQUIRE(batch == nullptr || ld >= batch->d, "ld < d");
    MB_REQUIRE(num_rows >= 0, "num_rows < 0");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    MB_REQUIRE(reduction == MB_REDUCTION_MEAN || reduction == MB_REDUCTION_SUM, "unknown reduction");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    return train_step_any(ctx, batch, false, table, state_table, ld, unique_ids, lr, reduction, precision, loss, nullptr, rel_grad, inv_rel_grad,
                          (cudaStream_t)stream);
}

static mb_status check_shards(const mb_shards* sh) {
    MB_REQUIRE(sh != nullptr && sh->world >= 1 && sh->world <= 8 && sh->rows_per_rank > 0, "bad shard description");
    MB_REQUIRE(sh->rank >= 0 && sh->rank < sh->world, "shard rank out of range");
    for (int i = 0; i < sh->world; i++) MB_REQUIRE(sh->tables[i] != nullptr && sh->states[i] != nullptr, "null shard pointer");
    if (sh->world > 1) {
        for (int i = 0; i < sh->world; i++) MB_REQUIRE(sh->exchange[i] != nullptr, "null exchange area (mb_shard_exchange_bytes)");
        MB_REQUIRE(sh->exchange_rows > 0, "exchange_rows must be positive");
    }
    return MB_OK;
}

static mb_status check_shard_capacity(const mb_shards* sh, const mb_batch* b) {
    if (sh != nullptr && sh->world > 1 && b != nullptr)
        MB_REQUIRE(2 * b->B + 2 * (int64_t)b->C * b->N <= sh->exchange_rows, "batch larger than the shards' exchange_rows (2B + 2CN gradient rows per inbox)");
    return MB_OK;
}

mb_status mb_train_step_sharded(mb_context* ctx, const mb_batch* batch, const mb_shards* shards, int64_t ld, const int64_t* unique_ids, float lr,
                                int reduction, int precision, float* loss, float* rel_grad, float* inv_rel_grad, void* stream) {
    MB_REQUIRE(ctx != nullptr && unique_ids != nullptr, "null context / ids");
    MB_TRY(check_shards(shards));
    MB_TRY(check_shard_capacity(shards, batch));
    MB_REQUIRE(batch == nullptr || ld >= batch->d, "ld < d");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    MB_REQUIRE(reduction == MB_REDUCTION_MEAN || reduction == MB_REDUCTION_SUM, "unknown reduction");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    return train_step_any(ctx, batch, false, shards->tables[shards->rank], shards->states[shards->rank], ld, unique_ids, lr, reduction, precision, loss, nullptr, rel_grad,
                          inv_rel_grad, (cudaStream_t)stream, shards);
}

// Host-buffer steps.  Batch::to (batch.cpp:21-60): the index tensors go host -> device on the compute stream and the loss comes back into
// one of the context's two pinned slots.  The asynchronous form returns once the step is enqueued (ticket = the slot); the caller keeps
// the host buffers alive until mb_train_step_host_wait(ticket) has returned, and has at most two steps in flight.
static mb_status host_step_enqueue(mb_context* ctx, const mb_batch* hb, const mb_shards* shards, float* table, float* state_table, int64_t ld,
                                   const int64_t* unique_ids_host, float lr, int reduction, int precision, float* rel_grad, float* inv_rel_grad,
                                   cudaStream_t st, int* ticket) {
    MB_REQUIRE(ctx != nullptr && unique_ids_host != nullptr, "null context / ids");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    MB_REQUIRE(reduction == MB_REDUCTION_MEAN || reduction == MB_REDUCTION_SUM, "unknown reduction");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    const int slot = ctx->loss_slot;
    static const bool prefetch = [] { const char* e = getenv("MB_H2D_PREFETCH"); return e ? atoi(e) != 0 : true; }();
    if (prefetch && hb != nullptr && hb->B > 0 && hb->edges != nullptr && hb->dst_negs != nullptr) {
        // Batch::to on its own stream (batch.cpp:21-60 uses a pool stream + record_stream the same way): the four index tensors go into
        // device slot `slot` over the copy stream, so the PCIe transfer of this batch runs while the previous batch computes; the step
        // itself then takes device inputs (its graph starts with device-to-device copies into the fixed staging buffers).
        const int64_t n_u = hb->U, n_e = hb->B * hb->edge_cols, n_n = (int64_t)hb->C * hb->N;
        const bool has_s = hb->src_negs != nullptr;
        const size_t need = (size_t)(n_u + n_e + n_n + (has_s ? n_n : 0));
        if (need > ctx->pf_cap[slot]) {
            MB_CUDA_TRY(cudaEventSynchronize(ctx->ev_loss[slot]));  // the slot's previous step has consumed it
            if (ctx->pf[slot]) MB_CUDA_TRY(cudaFree(ctx->pf[slot]));
            ctx->pf[slot] = nullptr;
            ctx->pf_cap[slot] = 0;
            MB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->pf[slot]), sizeof(int64_t) * (need + need / 8 + 1024)));
            ctx->pf_cap[slot] = need + need / 8 + 1024;
        }
        int64_t* d_u = ctx->pf[slot];
        int64_t* d_e = d_u + n_u;
        int64_t* d_dn = d_e + n_e;
        int64_t* d_sn = d_dn + n_n;
        MB_CUDA_TRY(cudaStreamWaitEvent(ctx->copy, ctx->ev_loss[slot], 0));  // (a no-op unless the caller broke the two-in-flight rule)
        MB_CUDA_TRY(cudaMemcpyAsync(d_u, unique_ids_host, sizeof(int64_t) * n_u, cudaMemcpyHostToDevice, ctx->copy));
        MB_CUDA_TRY(cudaMemcpyAsync(d_e, hb->edges, sizeof(int64_t) * n_e, cudaMemcpyHostToDevice, ctx->copy));
        MB_CUDA_TRY(cudaMemcpyAsync(d_dn, hb->dst_negs, sizeof(int64_t) * n_n, cudaMemcpyHostToDevice, ctx->copy));
        if (has_s) MB_CUDA_TRY(cudaMemcpyAsync(d_sn, hb->src_negs, sizeof(int64_t) * n_n, cudaMemcpyHostToDevice, ctx->copy));
        MB_CUDA_TRY(cudaEventRecord(ctx->ev_copied[slot], ctx->copy));
        MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_copied[slot], 0));
        mb_batch db = *hb;
        db.edges = d_e;
        db.dst_negs = d_dn;
        db.src_negs = has_s ? d_sn : nullptr;
        MB_TRY(train_step_any(ctx, &db, false, table, state_table, ld, d_u, lr, reduction, precision, nullptr, ctx->h_loss_pinned + slot, rel_grad,
                              inv_rel_grad, st, shards));
    } else {
        MB_TRY(train_step_any(ctx, hb, true, table, state_table, ld, unique_ids_host, lr, reduction, precision, nullptr, ctx->h_loss_pinned + slot, rel_grad,
                              inv_rel_grad, st, shards));
    }
    MB_CUDA_TRY(cudaEventRecord(ctx->ev_loss[slot], st));
    ctx->loss_slot = slot ^ 1;
    *ticket = slot;
    return MB_OK;
}

// Steps from RAW edges: DataLoader::getBatch -> edgeSample (dataloader.cpp:389-471) happens on the device.  The host hands over only the
// batch's positive edges with GLOBAL node ids ([B, 3] int64, 24 B per edge); negatives are drawn on the device (mb_sample_negatives), the
// unique-id mapping is the device radix sort (mb_edge_sample), and the fused step runs on the result with U = capacity (the unique list
// is padded with -1).  No host-side unique, 1.2 MB instead of 3.6 MB over PCIe per 50 000-edge step.
mb_status mb_train_step_edges_host_async(mb_context* ctx, int decoder, const int64_t* edges_host, int64_t B, int64_t num_nodes, int C, int N, uint64_t seed,
                                         uint32_t batch_index, const float* rel, const float* inv_rel, int64_t R, int64_t d, float* table,
                                         float* state_table, int64_t num_rows, int64_t ld, float lr, int reduction, int precision, float* rel_grad,
                                         float* inv_rel_grad, int* ticket, void* stream) {
    MB_REQUIRE(ctx != nullptr && edges_host != nullptr && ticket != nullptr, "null argument");
    MB_REQUIRE(table != nullptr && state_table != nullptr, "null table");
    MB_REQUIRE(B > 0 && C > 0 && N > 0 && num_nodes > 0 && num_nodes <= num_rows && d > 0, "bad dimensions");
    MB_REQUIRE(precision >= MB_PREC_FP32 && precision <= MB_PREC_BF16, "unknown precision");
    MB_REQUIRE(reduction == MB_REDUCTION_MEAN || reduction == MB_REDUCTION_SUM, "unknown reduction");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    const bool inverse = inv_rel != nullptr;
    const int64_t CN = (int64_t)C * N, n_e = 3 * B, cap_u = 2 * B + 2 * CN;
    // layout of the staging block (int64 units)
    const size_t need = (size_t)(n_e + 2 * CN + cap_u + n_e + 2 * CN + 2);
    if (need > ctx->raw_cap) {
        MB_CUDA_TRY(cudaStreamSynchronize(st));
        MB_CUDA_TRY(cudaStreamSynchronize(ctx->copy));
        drop_graph(ctx);
        if (ctx->raw) MB_CUDA_TRY(cudaFree(ctx->raw));
        ctx->raw = nullptr;
        ctx->raw_cap = 0;
        MB_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ctx->raw), sizeof(int64_t) * (need + need / 8)));
        ctx->raw_cap = need + need / 8;
    }
    int64_t* e_glob = ctx->raw;
    int64_t* dn_glob = e_glob + n_e;
    int64_t* sn_glob = dn_glob + CN;
    int64_t* uniq = sn_glob + CN;
    int64_t* e_loc = uniq + cap_u;
    int64_t* dn_loc = e_loc + n_e;
    int64_t* sn_loc = dn_loc + CN;
    int64_t* num = sn_loc + CN;
    const int slot = ctx->loss_slot;
    // Batch::to of the raw edges on the copy stream (ordered behind the previous step's use of the staging block)
    MB_CUDA_TRY(cudaEventRecord(ctx->ev_in, st));
    MB_CUDA_TRY(cudaStreamWaitEvent(ctx->copy, ctx->ev_in, 0));
    MB_CUDA_TRY(cudaMemcpyAsync(e_glob, edges_host, sizeof(int64_t) * n_e, cudaMemcpyHostToDevice, ctx->copy));
    MB_CUDA_TRY(cudaEventRecord(ctx->ev_copied[slot], ctx->copy));
    MB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_copied[slot], 0));
    {
        StageTimer tm(ctx, ST_SAMPLE, st);
        MB_TRY(launch_sample_negatives(num_nodes, C, N, 0, e_glob, B, 3, false, seed, batch_index, dn_glob, st));
        if (inverse) MB_TRY(launch_sample_negatives(num_nodes, C, N, 0, e_glob, B, 3, true, seed, batch_index, sn_glob, st));
        MB_TRY(mb_edge_sample(ctx, e_glob, B, 3, inverse ? sn_glob : nullptr, dn_glob, C, N, num_nodes - 1, uniq, e_loc, inverse ? sn_loc : nullptr, dn_loc, num, st));
    }
    mb_batch db;
    std::memset(&db, 0, sizeof(db));
    db.decoder = decoder;
    db.U = cap_u;  // capacity: entries past the batch's unique count are -1 (padding segments are empty and skipped)
    db.d = d;
    db.B = B;
    db.R = R;
    db.C = C;
    db.N = N;
    db.edges = e_loc;
    db.edge_cols = 3;
    db.dst_negs = dn_loc;
    db.src_negs = inverse ? sn_loc : nullptr;
    db.rel = rel;
    db.inv_rel = inv_rel;
    MB_TRY(train_step_any(ctx, &db, false, table, state_table, ld, uniq, lr, reduction, precision, nullptr, ctx->h_loss_pinned + slot, rel_grad, inv_rel_grad, st));
    MB_CUDA_TRY(cudaEventRecord(ctx->ev_loss[slot], st));
    ctx->loss_slot = slot ^ 1;
    *ticket = slot;
    return MB_OK;
}

mb_status mb_train_step_host_wait(mb_context* ctx, int ticket, float* loss_host) {
    MB_REQUIRE(ctx != nullptr && (ticket == 0 || ticket == 1), "bad ticket");
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    MB_CUDA_TRY(cudaEventSynchronize(ctx->ev_loss[ticket]));
    if (loss_host) *loss_host = ctx->h_loss_pinned[ticket];
    return MB_OK;
}

mb_status mb_train_step_sharded_host_async(mb_context* ctx, const mb_batch* hb, const mb_shards* shards, int64_t ld, const int64_t* unique_ids_host,
                                           float lr, int reduction, int precision, float* rel_grad, float* inv_rel_grad, int* ticket, void* stream) {
    MB_REQUIRE(ticket != nullptr, "ticket is null");
    MB_TRY(check_shards(shards));
    MB_TRY(check_shard_capacity(shards, hb));
    return host_step_enqueue(ctx, hb, shards, shards->tables[shards->rank], shards->states[shards->rank], ld, unique_ids_host, lr, reduction, precision,
                             rel_grad, inv_rel_grad, (cudaStream_t)stream, ticket);
}

mb_status mb_train_step_sharded_host(mb_context* ctx, const mb_batch* hb, const mb_shards* shards, int64_t ld, const int64_t* unique_ids_host, float lr,
                                     int reduction, int precision, float* loss_host, float* rel_grad, float* inv_rel_grad, void* stream) {
    int ticket = 0;
    MB_TRY(mb_train_step_sharded_host_async(ctx, hb, shards, ld, unique_ids_host, lr, reduction, precision, rel_grad, inv_rel_grad, &ticket, stream));
    return mb_train_step_host_wait(ctx, ticket, loss_host);
}

// Diagnostic: D[b] = A[b] . B[b] over K through the contraction kernels, fp32 in / fp32 out.
//   a_mn == 0: A is [batches][M][K] (K contiguous)      a_mn == 1: A is [batches][K][M]
//   b_mn == 0: B is [batches][N][K]                     b_mn == 1: B is [batches][K][N]
mb_status mb_debug_gemm(mb_context* ctx, const float* A, int a_mn, const float* B, int b_mn, float* D, int M, int N, int K, int batches, int precision,
                        int block_n, void* stream) {
    MB_REQUIRE(ctx && A && B && D && M > 0 && N > 0 && K > 0 && batches > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t na = (int64_t)batches * M * K, nb = (int64_t)batches * N * K;
    if (precision == MB_PREC_FP32) {
        return gemm_simt(A, a_mn ? 1 : K, a_mn ? M : 1, (int64_t)M * K, B, b_mn ? N : 1, b_mn ? 1 : K, (int64_t)N * K, D, N, (int64_t)M * N, M, N, K,
                         batches, st);
    }
    MB_REQUIRE(gemm_tc_supported(a_mn ? M : K, b_mn ? N : K), "tcgen05 path needs inner extents that are multiples of 8");
    Arena sizing(nullptr);
    sizing.take<__nv_bfloat16>(2 * na);
    sizing.take<__nv_bfloat16>(2 * nb);
    MB_TRY(ensure_ws(ctx, sizing.off + 256, st));
    Arena place(ctx->ws);
    __nv_bfloat16* a_hl = place.take<__nv_bfloat16>(2 * na);
    __nv_bfloat16* b_hl = place.take<__nv_bfloat16>(2 * nb);
    MB_TRY(launch_split(B, nb, b_hl, b_hl + nb, st));
    if (block_n == 2 || block_n == 3) {
        // diagnostics of the backward kernels: the A operand is converted inside the kernel from the fp32 matrix (conv_mode 2: identity);
        // block_n == 3 additionally forces the shared-memory-A kernel
        MB_REQUIRE(b_mn, "the converting kernels take an MN-major B operand");
        TcGroupProblem g{nullptr, nullptr, 0, 0, a_mn ? 1 : 0, b_hl, b_hl + nb, N, (int64_t)N * K, 1, D, N, (int64_t)M * N, M, N, K, batches};
        g.conv_src = A;
        g.conv_ld = a_mn ? M : K;
        g.conv_sb = (int64_t)M * K;
        g.conv_rows = a_mn ? K : M;
        g.conv_cols = a_mn ? M : K;
        g.conv_mode = 2;
        return gemm_tc_grouped(&g, 1, precision == MB_PREC_BF16 ? 1 : 3, st, block_n == 3);
    }
    if (block_n == 4 || block_n == 5) {
        // diagnostics: BOTH backward problems in one grouped launch, from one square fp32 matrix S [batches][M][M] (K == M):
        //   D[b] = f(S[b]) . B[b]   and   D[batches + b] = f(S[b])^T . B[batches + b],   f = identity (4) or exp (5, zero shifts)
        // B is [2 * batches][K][N] (MN-major), D is [2 * batches][M][N].
        MB_REQUIRE(b_mn && !a_mn && M == K, "grouped diagnostics: square A, MN-major B");
        Arena sz2(nullptr);
        sz2.take<__nv_bfloat16>(4 * nb);
        sz2.take<float>((int64_t)batches * M);
        MB_TRY(ensure_ws(ctx, sz2.off + 256, st));
        Arena pl2(ctx->ws);
        __nv_bfloat16* b2 = pl2.take<__nv_bfloat16>(4 * nb);
        float* z = pl2.take<float>((int64_t)batches * M);
        MB_TRY(launch_split(B, 2 * nb, b2, b2 + 2 * nb, st));
        MB_CUDA_TRY(cudaMemsetAsync(z, 0, sizeof(float) * batches * M, st));
        TcGroupProblem g[2];
        for (int q = 0; q < 2; q++) {
            g[q] = TcGroupProblem{nullptr, nullptr, 0, 0, q, b2 + q * nb, b2 + 2 * nb + q * nb, N, (int64_t)N * K, 1, D + (int64_t)q * batches * M * N, N, (int64_t)M * N, M, N, K, batches};
            g[q].conv_src = A;
            g[q].conv_ld = M;
            g[q].conv_sb = (int64_t)M * M;
            g[q].conv_rows = M;
            g[q].conv_cols = M;
            g[q].conv_mode = block_n == 5 ? 1 : 2;
            g[q].conv_z = z;
        }
        return gemm_tc_grouped(g, 2, precision == MB_PREC_BF16 ? 1 : 3, st);
    }
    MB_TRY(launch_split(A, na, a_hl, a_hl + na, st));
    return tc_contract(a_hl, a_hl + na, a_mn ? M : K, (int64_t)M * K, a_mn != 0, b_hl, b_hl + nb, b_mn ? N : K, (int64_t)N * K, b_mn != 0, D, N,
                       (int64_t)M * N, M, N, K, batches, precision == MB_PREC_BF16 ? 1 : 3, st);
}

int mb_debug_wait_log(uint64_t* out, int cap) { return gemm_tc_wait_log(reinterpret_cast<unsigned long long*>(out), cap); }

static mb_status grow_i64(int64_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return MB_OK;
    if (*p) MB_CUDA_TRY(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = need + need / 4 + 1024;
    MB_CUDA_TRY(cudaMalloc(p, want * sizeof(int64_t)));
    *cap = want;
    return MB_OK;
}

mb_status mb_train_step_host_async(mb_context* ctx, const mb_batch* hb, float* table, float* state_table, int64_t num_rows, int64_t ld,
                                   const int64_t* unique_ids_host, float lr, int reduction, int precision, float* rel_grad, float* inv_rel_grad,
                                   int* ticket, void* stream) {
    MB_REQUIRE(table != nullptr && state_table != nullptr, "null table");
    MB_REQUIRE(ticket != nullptr, "ticket is null");
    (void)num_rows;
    return host_step_enqueue(ctx, hb, nullptr, table, state_table, ld, unique_ids_host, lr, reduction, precision, rel_grad, inv_rel_grad,
                             (cudaStream_t)stream, ticket);
}

mb_status mb_train_step_host(mb_context* ctx, const mb_batch* hb, float* table, float* state_table, int64_t num_rows, int64_t ld,
                             const int64_t* unique_ids_host, float lr, int reduction, int precision, float* loss_host, float* rel_grad,
                             float* inv_rel_grad, void* stream) {
    int ticket = 0;
    MB_TRY(mb_train_step_host_async(ctx, hb, table, state_table, num_rows, ld, unique_ids_host, lr, reduction, precision, rel_grad, inv_rel_grad, &ticket,
                                    stream));
    return mb_train_step_host_wait(ctx, ticket, loss_host);
}

}  // extern "C"
