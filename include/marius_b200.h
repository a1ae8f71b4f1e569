/* marius_b200.h -- C ABI of the B200-native Marius embedding hot path.
 *
 * One shared library (marius_b200/lib/libmarius_b200.so, hand-written sm_100a CUDA) behind plain
 * pointers and sizes: no torch types, no C++ types, no exceptions.  Every entry point returns an
 * mb_status; mb_last_error() gives the message for the calling thread.
 *
 * The reference (marius-team/marius) has no plugin/FFI ABI; its seam is a set of C++ virtual
 * interfaces that pass torch::Tensor.  Each entry point below names the reference
 * call site(s) it replaces (paths relative to /root/reference/src/cpp).  The C++/libtorch adapters
 * that keep the reference's class surface (Storage, PartitionBuffer, EdgeDecoder, Batch, Model) live
 * in marius_b200/csrc/host/ and only ever call this header.  INTEGRATION.md shows the binding a
 * Marius maintainer would add.
 *
 * Conventions
 *   - ids are int64, values fp32, matrices row-major with an explicit leading dimension (elements).
 *   - every device pointer must belong to the device the context was created on.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless stated.  Calls are re-entrant across contexts; one context = one in-flight batch.
 *   - "unique ids within one call" is the only exclusivity the update kernels assume
 *     (storage/buffer.cpp:459), exactly as the reference.
 */
#ifndef MARIUS_B200_H_
#define MARIUS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_VERSION 100

typedef enum mb_status {
    MB_OK = 0,
    MB_ERR_INVALID = 1,     /* bad shapes / null pointers: the reference throws std::runtime_error (storage.cpp:607-610,652-655) */
    MB_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed */
    MB_ERR_UNSUPPORTED = 3, /* valid in the reference, not implemented here (documented in DESIGN.md) */
    MB_ERR_NOMEM = 4
} mb_status;

/* decoder kinds: the DotCompare family of include/configuration/options.h:60 */
typedef enum mb_decoder {
    MB_DECODER_DOT = 0,      /* 2-column edges, no relation operator  (decoder_methods.cpp:99-101, comparators.cpp:62-72) */
    MB_DECODER_DISTMULT = 1, /* HadamardOperator        + DotCompare  (distmult.cpp:7-19, relation_operators.cpp:7-12)  */
    MB_DECODER_COMPLEX = 2   /* ComplexHadamardOperator + DotCompare  (complex.cpp:7-19,  relation_operators.cpp:14-35) */
} mb_decoder;

typedef enum mb_reduction { MB_REDUCTION_MEAN = 0, MB_REDUCTION_SUM = 1 } mb_reduction; /* options.h:24 */

/* arithmetic of the dense (batch x dim).(neg x dim)^T contractions (comparators.cpp:69-72 and their autograd) */
typedef enum mb_precision {
    MB_PREC_FP32 = 0,   /* fp32 FFMA on the SIMT pipes: reference-exact arithmetic, any d                          */
    MB_PREC_BF16X3 = 1, /* tcgen05 kind::f16, operands split hi+lo in bf16, 3 products, fp32 TMEM accumulation:
                           |err| <= ~2^-16 relative per product -- meets the 1e-4 parity bar   */
    MB_PREC_BF16 = 2    /* single bf16 product (fast, does NOT meet the parity bar; never the default)           */
} mb_precision;

typedef struct mb_context mb_context; /* device, SM count, workspace arena, TMA descriptors, sort scratch */

/* ---- context ----------------------------------------------------------------------------------------- */
mb_status mb_create(int device, mb_context** out);
void mb_destroy(mb_context* ctx);
const char* mb_last_error(void);
int mb_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches counter) */
uint64_t mb_launch_count(void);
/* name of the largest tensor-core path compiled in ("tcgen05" when built for sm_100a) */
const char* mb_build_info(void);

/* ---- storage: gather / scatter-add ---------------------------------------------------------------------
 * mb_gather_rows        InMemory::indexRead CUDA branch  storage/storage.cpp:613-614 (data_.index_select)
 *                       PartitionBuffer::indexRead       storage/buffer.cpp:441-455 (HBM-resident slab)
 * mb_scatter_add_rows   InMemory::indexAdd CUDA branch   storage/storage.cpp:656-657 (data_.index_add_)
 *                       PartitionBuffer::indexAdd        storage/buffer.cpp:459-480
 * mb_scatter_put_rows   InMemory::indexPut               storage/storage.cpp:675-690
 * ids must be in [0, num_rows); this is checked on the device only when `check` != 0 (the reference
 * does not check either); a violation makes the call return MB_ERR_INVALID after a stream sync.        */
mb_status mb_gather_rows(const float* table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n, float* out, int64_t out_ld,
                         void* stream);
mb_status mb_scatter_add_rows(float* table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* vals,
                              int64_t vals_ld, void* stream);
mb_status mb_scatter_put_rows(float* table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* vals,
                              int64_t vals_ld, void* stream);
/* PartitionBuffer::getGlobalToLocalMap (buffer.cpp:581-633): map[g] = slot*partition_size + (g - p*partition_size) for every
 * resident partition p (given as parallel arrays partition id -> buffer slot, host memory), -1 elsewhere.  `map` is device memory. */
mb_status mb_global_to_local_map(int64_t* map, int64_t total_rows, int64_t partition_size, const int32_t* partition_ids,
                                 const int32_t* buffer_slots, int n_resident, void* stream);

/* ---- sparse Adagrad ------------------------------------------------------------------------------------
 * mb_adagrad_deltas       Batch::accumulateGradients  data/batch.cpp:62-79 (returns delta_e, delta_s; op-for-op fp32)
 * mb_adagrad_update_rows  = accumulateGradients + DataLoader::updateEmbeddings (dataloader.cpp:550-564 ->
 *                           graph_storage.cpp:289,319-323 -> Storage::indexAdd x2) fused: one read-modify-write
 *                           of each unique row of the table and of the state table.                     */
mb_status mb_adagrad_deltas(const float* grad, const float* state, int64_t n, int64_t d, int64_t ld, float lr, float* delta_e, float* delta_s,
                            void* stream);
mb_status mb_adagrad_update_rows(float* table, float* state_table, int64_t num_rows, int64_t ld, int64_t d, const int64_t* idx, int64_t n,
                                 const float* grad, int64_t grad_ld, float lr, void* stream);

/* ---- unique-id mapping ----------------------------------------------------------------------------------
 * mb_map_tensors  map_tensors  common/util.cpp:180-205 (cat + torch::_unique2(sorted, inverse)), as used by
 *                 DataLoader::edgeSample (dataloader.cpp:399-409,447-461).  `all_ids` [n] global ids ->
 *                 `unique_out` [<= n] sorted unique ids, `mapped_out` [n] position of each input id, *num_unique
 *                 (device int64).  Stable LSD radix sort on the device; bit-identical to the reference.  Entries of
 *                 unique_out past *num_unique are set to -1.                                                  */
mb_status mb_map_tensors(mb_context* ctx, const int64_t* all_ids, int64_t n, int64_t max_id, int64_t* unique_out, int64_t* mapped_out,
                         int64_t* num_unique_dev, void* stream);

/* ---- batch assembly ----------------------------------------------------------------------------------------
 * mb_sample_negatives  CorruptNodeNegativeSampler::getNegatives (data/samplers/negative.cpp:328-366): out [C,N] int64 (device).  Per
 *   chunk the first (int)(N * degree_fraction) ids are an endpoint of a random edge of the batch (batch_sample, negative.cpp:7-19:
 *   the source column when inverse != 0, else the destination column), the rest are uniform in [0, num_nodes).  The ids are a pure
 *   function of (seed, batch_index, inverse, position): Philox4x32-10, restated by oracle/marius_oracle.py::sample_negatives -- the
 *   reference's own stream (libtorch's global generator) is not reproducible on a device, see DESIGN.md.
 * mb_edge_sample  DataLoader::edgeSample without a neighbour sampler (data/dataloader.cpp:389-471): all_ids = cat(src, dst,
 *   src_negs, dst_negs) -> map_tensors -> unique_out (sorted unique global ids, entries past *num_unique_dev are -1; capacity
 *   2B + CN (+ CN)), edges_local [B,cols] = (mapped src, rel, mapped dst), src_negs_local / dst_negs_local [C,N].  src_negs may be NULL
 *   (no inverse side).  Bit-identical to the reference for the same global ids. */
mb_status mb_sample_negatives(int64_t num_nodes, int C, int N, float degree_fraction, const int64_t* edges, int64_t B, int edge_cols, int inverse,
                              uint64_t seed, uint32_t batch_index, int64_t* out, void* stream);
mb_status mb_edge_sample(mb_context* ctx, const int64_t* edges, int64_t B, int edge_cols, const int64_t* src_negs, const int64_t* dst_negs, int C, int N,
                         int64_t max_id, int64_t* unique_out, int64_t* edges_local, int64_t* src_negs_local, int64_t* dst_negs_local,
                         int64_t* num_unique_dev, void* stream);

/* mb_reduce_rows_by_key: rows_out[u,:] = sum of rows[i,:] over all i with ids[i] == unique_out[u]; unique_out sorted ascending,
 * *num_unique (device).  The owner-side merge of the multi-GPU row exchange: gradient rows for the same table
 * row arriving from different ranks are summed (fixed order, no atomics) before the Adagrad update.  unique_out and rows_out hold n
 * entries; entries past *num_unique are -1 / zero rows, and mb_adagrad_update_rows skips negative ids, so the pair can be fed to
 * the update without a host round trip. */
mb_status mb_reduce_rows_by_key(mb_context* ctx, const int64_t* ids, const float* rows, int64_t n, int64_t d, int64_t max_id, int64_t* unique_out,
                                float* rows_out, int64_t* num_unique_dev, void* stream);

/* ---- decoder + training step ---------------------------------------------------------------------------
 * Batch-local problem, exactly the tensors the reference Batch carries (data/batch.h:49-75):
 *   emb   [U,d]  node_embeddings_            edges [B,3] (or [B,2] when kind == DOT): local src, rel, local dst
 *   dst_negs / src_negs [C,N]  dst_neg_indices_mapping_ / src_neg_indices_mapping_ (local ids)
 *   rel / inv_rel [R,d]  EdgeDecoder::relations_ / inverse_relations_ (edge_decoder.h:17-18); inv_rel == NULL or
 *   src_negs == NULL -> use_inverse_relations_ = false.
 * Bp = C * ceil(B / C): rows are zero-padded like pad_and_reshape (comparators.cpp:7-20) and pos scores are
 * zero-padded to Bp (decoder_methods.cpp:103-111).                                                          */
typedef struct mb_batch {
    int decoder;      /* mb_decoder */
    int64_t U, d, B, R;
    int C, N;
    const int64_t* edges;    /* [B,3] or [B,2] */
    int edge_cols;           /* 3 or 2 */
    const int64_t* dst_negs; /* [C,N] */
    const int64_t* src_negs; /* [C,N] or NULL */
    const float* rel;        /* [R,d] or NULL */
    const float* inv_rel;    /* [R,d] or NULL */
} mb_batch;

/* Model::forward_lp -> node_corrupt_forward (nn/model.cpp:252-288, decoders/edge/decoder_methods.cpp:57-114):
 * pos [Bp], neg [Bp,N], inv_pos [Bp], inv_neg [Bp,N] (inverse outputs may be NULL).                       */
mb_status mb_decoder_forward(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, int precision, float* pos, float* neg,
                             float* inv_pos, float* inv_neg, void* stream);

/* ---- evaluation ----------------------------------------------------------------------------------------
 * apply_score_filter (data/samplers/negative.cpp:306-311, called by Model::forward_lp, nn/model.cpp:279-285):
 * scores[filter[i,0], filter[i,1]] = -1e9 for the F (row, column) pairs of `filter` ([F,2] int64, device).  Like index_put_, a
 * negative index wraps once and an out-of-range pair is an error (MB_ERR_INVALID; reported after a stream synchronisation). */
mb_status mb_apply_score_filter(float* scores, int64_t rows, int64_t N, int64_t ld, const int64_t* filter, int64_t F, void* stream);

/* LinkPredictionReporter::computeRanks (reporting/reporting.cpp:56-58): ranks[i] = 1 + #{ j : neg[i,j] >= pos[i] }, int64. */
mb_status mb_compute_ranks(const float* pos, const float* neg, int64_t rows, int64_t N, int64_t ld, int64_t* ranks, void* stream);

/* Model::evaluate_batch for link prediction (nn/model.cpp:335-349): forward_lp with the batch's score filters
 * (dst_filter on the dst-corruption scores, src_filter on the src-corruption scores; NULL / 0 = none), then one computeRanks per
 * side.  ranks [Bp] and inv_ranks [Bp] (inverse side; may be NULL when the batch has none) are int64 device arrays in the order
 * LinkPredictionReporter::addResult receives them; padding rows (B <= i < Bp) rank N + 1 exactly as in the reference.  The score
 * matrices stay in the context's workspace; pos / inv_pos [Bp] are optional outputs (NULL to skip). */
mb_status mb_evaluate_batch(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, int precision, const int64_t* dst_filter,
                            int64_t Fd, const int64_t* src_filter, int64_t Fs, int64_t* ranks, int64_t* inv_ranks, float* pos, float* inv_pos,
                            void* stream);

/* Backward of node_corrupt_forward for ARBITRARY upstream gradients (what libtorch autograd computes under loss.backward(),
 * model.cpp:324, for any of the reference's loss functions, nn/loss.cpp:50-175): gpos [Bp], gneg [Bp,N] (+ inverse side) ->
 * grad [U,d] = d/d node_embeddings_, rel_grad / inv_rel_grad [R,d].  Used by the autograd::Function of the C++ adapter. */
mb_status mb_decoder_backward(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, int precision, const float* gpos,
                              const float* gneg, const float* ginv_pos, const float* ginv_neg, float* grad, float* rel_grad, float* inv_rel_grad,
                              void* stream);

/* Model::train_batch for link prediction (nn/model.cpp:290-333) on batch-local tensors: forward_lp, SoftmaxCrossEntropy on
 * both sides (nn/loss.cpp:50-67, model.cpp:309-312), backward, Batch::accumulateGradients (batch.cpp:62-79).
 * Outputs (device; any may be NULL): loss [1]; grad [U,d] = d loss / d node_embeddings_;
 * delta_e, delta_s [U,d] (node_gradients_, node_state_update_); rel_grad, inv_rel_grad [R,d] (overwritten).  */
mb_status mb_train_batch(mb_context* ctx, const mb_batch* batch, const float* emb, int64_t emb_ld, const float* state, int64_t state_ld,
                         float lr, int reduction, int precision, float* loss, float* grad, float* delta_e, float* delta_s, float* rel_grad,
                         float* inv_rel_grad, void* stream);

/* The whole per-batch path on a device-resident table (SynchronousTrainer::train loop body, pipeline/trainer.cpp:106-138
 * == ComputeWorkerGPU::run, pipeline/pipeline_gpu.cpp:49-91):
 *   loadGPUParameters (gather emb; the state row is read in place)  dataloader.cpp:529-548
 *   Model::train_batch                                               model.cpp:290-333
 *   updateEmbeddings(batch, gpu=true) (scatter-add emb + state)     dataloader.cpp:550-557
 * `unique_ids` [U] are table rows (global ids for InMemory, buffer-local ids for the partition buffer).
 * rel_grad / inv_rel_grad [R,d] receive the dense relation gradients (the reference's dense optimizer, nn/optim.cpp,
 * consumes them; mb_dense_adagrad_step below is that optimizer).  loss [1] device.                       */
mb_status mb_train_step(mb_context* ctx, const mb_batch* batch, float* table, float* state_table, int64_t num_rows, int64_t ld,
                        const int64_t* unique_ids, float lr, int reduction, int precision, float* loss, float* rel_grad, float* inv_rel_grad,
                        void* stream);

/* The same fused step on a table SHARDED BY NODE PARTITION over the GPUs of one box (replaces the reference's
 * replicas-plus-shared-host-table scheme, nn/model.cpp:136-159, pipeline/pipeline_gpu.cpp:23).  Rank o owns global rows
 * [o * rows_per_rank, (o+1) * rows_per_rank); tables[o] / states[o] / exchange[o] are the owners' base pointers as seen from THIS
 * process: its own allocations for o == rank, CUDA-IPC mappings of the peers' allocations otherwise (peer access over NVLink).
 * `unique_ids` are GLOBAL row ids, sorted ascending (map_tensors order).  Per step and rank:
 *   1. remote embedding rows are fetched once into a batch-local cache by a copy kernel that keeps many rows per SM in flight
 *      (NVLink latency is ~5x HBM latency); the decoder kernels then read local memory only;
 *   2. the update kernel applies Adagrad to the rows this rank owns and SHIPS THE GRADIENT ROW of every remote row into the owner's
 *      inbox (plain 128-bit stores over NVLink: one row out per remote row, the Adagrad state never crosses the link);
 *   3. every owner applies the gradient rows it received, sender by sender in rank order, each as one sparse-Adagrad step
 *      (batch.cpp:62-79) on its own HBM.
 * Three device-side flag barriers per step over the exchange areas (epoch counters written with st.release.sys, polled with
 * ld.acquire.sys; no host involvement, so the step still replays as one CUDA graph) order the phases: all fetches of step t see the
 * tables after every apply of step t-1; no row is updated while a peer may still fetch it; an inbox is applied only when complete.
 * The result is therefore deterministic and defined: pre-step rows + the owner's own Adagrad step + one Adagrad step per sender in
 * rank order.  All ranks must call the sharded step the same number of times (a barrier that is not met within ~2 s sets the
 * context's error flag instead of hanging the GPU).  Cross-GPU traffic: 1 row in + 1 row out per remote row.  Up to 8 shards.
 * single_process != 0: all shards are driven by this one call (tests; one GPU holding several shards): barriers are skipped and this
 * call also runs step 3 for every owner. */
typedef struct mb_shards {
    float* tables[8];
    float* states[8];
    int world;
    int64_t rows_per_rank;
    int rank; /* which shard is this process's own HBM (tables[rank] is local memory, the others are peer mappings) */
    void* exchange[8];     /* per-owner exchange area of mb_shard_exchange_bytes(world, exchange_rows, d) bytes, zero-initialised */
    int64_t exchange_rows; /* capacity (gradient rows) of every (owner, sender) inbox: >= 2B + 2CN of the largest batch */
    int single_process;
} mb_shards;
/* bytes of one rank's exchange area: barrier flags + `world` inboxes of (count, ids[rows], gradient rows[rows][d]) */
int64_t mb_shard_exchange_bytes(int world, int64_t exchange_rows, int64_t d);
/* non-zero once a barrier of the sharded step timed out on this context's device (sticky) */
mb_status mb_shard_error(mb_context* ctx, const mb_shards* shards, int* error_out);
/* CUDA IPC plumbing for one-process-per-GPU deployments: mb_ipc_export returns the 64-byte handle of the allocation containing
 * `dev_ptr` and the offset of `dev_ptr` inside it; mb_ipc_import opens a peer's handle with the context's device current (lazy peer
 * access over NVLink) and returns the peer pointer usable in mb_shards.  Mappings are closed by mb_destroy. */
mb_status mb_ipc_export(const void* dev_ptr, void* handle_out /* 64 bytes */, int64_t* offset_out);
mb_status mb_ipc_import(mb_context* ctx, const void* handle /* 64 bytes */, int64_t offset, void** ptr_out);
/* cudaDeviceEnablePeerAccess from the context's device to `peer_device` (idempotent): required before kernels dereference shards of that peer */
mb_status mb_enable_peer_access(mb_context* ctx, int peer_device);
mb_status mb_train_step_sharded(mb_context* ctx, const mb_batch* batch, const mb_shards* shards, int64_t ld, const int64_t* unique_ids, float lr,
                                int reduction, int precision, float* loss, float* rel_grad, float* inv_rel_grad, void* stream);
mb_status mb_train_step_sharded_host(mb_context* ctx, const mb_batch* host_batch, const mb_shards* shards, int64_t ld,
                                     const int64_t* unique_ids_host, float lr, int reduction, int precision, float* loss_host, float* rel_grad,
                                     float* inv_rel_grad, void* stream);

/* Same call with HOST index buffers (pinned or pageable) -- what a reference-side caller holding a CPU Batch would pass
 * (Batch::to, data/batch.cpp:21-60, moves exactly these tensors): copies unique_ids / edges / negatives to the device (on the
 * context's copy stream, into one of two device slots, as Batch::to uses a pool stream; MB_H2D_PREFETCH=0 keeps the copies on
 * `stream`), runs mb_train_step, copies the loss back into *loss_host and waits for the step.                                   */
mb_status mb_train_step_host(mb_context* ctx, const mb_batch* host_batch, float* table, float* state_table, int64_t num_rows, int64_t ld,
                             const int64_t* unique_ids_host, float lr, int reduction, int precision, float* loss_host, float* rel_grad,
                             float* inv_rel_grad, void* stream);
/* The same step without the final wait: returns once it is enqueued on `stream`; *ticket names the pinned slot the loss will land in.
 * mb_train_step_host_wait(ticket) blocks until that step has finished and returns its loss.  The host index buffers must stay valid
 * until then, and at most two steps may be in flight (two slots).  This is the reference's pipeline shape -- the transfer / compute
 * stages run ahead of the consumer of the result (pipeline/pipeline_gpu.cpp:10-145) -- and lets a caller enqueue batch i+1 while the
 * GPU still runs batch i. */
mb_status mb_train_step_host_async(mb_context* ctx, const mb_batch* host_batch, float* table, float* state_table, int64_t num_rows, int64_t ld,
                                   const int64_t* unique_ids_host, float lr, int reduction, int precision, float* rel_grad, float* inv_rel_grad,
                                   int* ticket, void* stream);
mb_status mb_train_step_sharded_host_async(mb_context* ctx, const mb_batch* host_batch, const mb_shards* shards, int64_t ld,
                                           const int64_t* unique_ids_host, float lr, int reduction, int precision, float* rel_grad,
                                           float* inv_rel_grad, int* ticket, void* stream);
mb_status mb_train_step_host_wait(mb_context* ctx, int ticket, float* loss_host);
/* The step from RAW positive edges (DataLoader::getBatch -> edgeSample -> loadGPUParameters -> train_batch -> updateEmbeddings,
 * dataloader.cpp:389-564 with the table device-resident): `edges_host` [B,3] int64 (src, rel, dst) with GLOBAL node ids, pinned or
 * pageable.  Negatives are sampled on the device (mb_sample_negatives, uniform, keyed by (seed, batch_index)), the unique-id mapping
 * is mb_edge_sample, then the fused step.  Asynchronous like mb_train_step_host_async (ticket / mb_train_step_host_wait); at most one
 * raw-edge step should be in flight per context (the staging block is single-buffered behind stream order).  inv_rel == NULL: no
 * inverse side. */
mb_status mb_train_step_edges_host_async(mb_context* ctx, int decoder, const int64_t* edges_host, int64_t B, int64_t num_nodes, int C, int N,
                                         uint64_t seed, uint32_t batch_index, const float* rel, const float* inv_rel, int64_t R, int64_t d,
                                         float* table, float* state_table, int64_t num_rows, int64_t ld, float lr, int reduction, int precision,
                                         float* rel_grad, float* inv_rel_grad, int* ticket, void* stream);


/* ---- filtered evaluation over ALL nodes ----------------------------------------------------------------------
 * mb_filter_sort_edges   the graph's edge list ([E, cols] int64, device) sorted stably by the endpoint a corruption KEEPS (source for
 *   destination corruption, destination when inverse != 0): the pool compute_filter_corruption searches (negative.cpp:62-112).  Done
 *   once per evaluation, not per batch.
 * mb_compute_filter      compute_filter_corruption with all nodes as negatives (negative.cpp:152-163, `filtered` evaluation): for batch
 *   edge i every pool edge with the same kept endpoint and relation yields (i, id of its other endpoint).  filter_out [cap, 2] receives
 *   the pairs ordered by batch edge, then pool order; *count_dev (device) the number of pairs (pairs beyond cap are not written).
 * mb_evaluate_all_nodes  Model::evaluate_batch with negatives = arange(num_nodes) (negative.cpp:321-325,355) WITHOUT materialising the
 *   [B, num_nodes] score matrix: the table streams through the score contraction in tiles of `tile_rows` rows, each tile is filtered
 *   and its (score >= positive) counts are added to the ranks.  `edges` hold GLOBAL node ids (rows of `table`), `table` is dense
 *   (ld == d).  ranks / inv_ranks [B] int64, pos_out / inv_pos_out [B] optional. */
mb_status mb_filter_sort_edges(mb_context* ctx, const int64_t* graph_edges, int64_t E, int edge_cols, int inverse, int64_t max_id, int64_t* sorted_out,
                               void* stream);
mb_status mb_compute_filter(mb_context* ctx, const int64_t* sorted_edges, int64_t E, int edge_cols, int inverse, const int64_t* batch_edges, int64_t B,
                            int64_t* filter_out, int64_t cap, int64_t* count_dev, void* stream);
mb_status mb_evaluate_all_nodes(mb_context* ctx, int decoder, const float* table, int64_t num_nodes, int64_t ld, int64_t d, const int64_t* edges, int64_t B,
                                int edge_cols, const float* rel, const float* inv_rel, int64_t R, const int64_t* dst_filter, int64_t Fd,
                                const int64_t* src_filter, int64_t Fs, int precision, int64_t tile_rows, int64_t* ranks, int64_t* inv_ranks, float* pos_out,
                                float* inv_pos_out, void* stream);

/* AdagradOptimizer::step on a dense parameter (nn/optim.cpp:114-145): state += g*g ; p -= lr * g / (sqrt(state) + eps) */
mb_status mb_dense_adagrad_step(float* param, float* state_sum, const float* grad, int64_t n, float lr, float eps, void* stream);

/* Diagnostic entry point for the contraction kernels (tests/, profiling): D[b] (MxN) = A[b] . B[b] over K, fp32 in/out.
 * a_mn == 0: A is [batches][M][K]; a_mn == 1: A is [batches][K][M].  b_mn likewise with N.  block_n: 2 = run the backward
 * contraction kernel (A operand converted in the kernel from the fp32 matrix, through tensor memory; b_mn must be 1), 3 = the same with
 * the A operand staged in shared memory; any other value = the plain kernel on pre-split bf16 operands. */
mb_status mb_debug_gemm(mb_context* ctx, const float* A, int a_mn, const float* B, int b_mn, float* D, int M, int N, int K, int batches,
                        int precision, int block_n, void* stream);
/* Diagnostics (MB_TC_WAITLOG=1 in the environment): the contraction kernels bound every barrier wait and trap instead of hanging; with the
 * log enabled a waiter that gives up first leaves (block << 32 | thread, barrier shared address << 32 | parity) in host-mapped memory.
 * Copies up to cap / 2 records into out and returns their number; readable even after the failed launch poisoned the context. */
int mb_debug_wait_log(uint64_t* out, int cap);

/* mb_train_step / mb_train_step_host replay the step as one CUDA graph from the second call with the same signature on (same shapes,
 * tables, relation tables, output pointers, stream); index tensors and the unique-row count may change freely.  MB_GRAPH=0 in the
 * environment or mb_graph_enable(ctx, 0) selects plain stream launches. */
mb_status mb_graph_enable(mb_context* ctx, int on);

/* Per-stage device timing of the step (bench.py's roofline): when enabled, every stage of mb_train_step / mb_train_batch is
 * bracketed by CUDA events recorded on the caller's stream.  mb_profile_read synchronises the device, fills total_ms[stage] and
 * counts[stage] (arrays of mb_profile_num_stages() entries) with the time accumulated since the last read, and resets them. */
mb_status mb_profile_enable(mb_context* ctx, int on);
int mb_profile_num_stages(void);
const char* mb_profile_stage_name(int stage);
mb_status mb_profile_read(mb_context* ctx, float* total_ms, int* counts);
/* mb_profile_enable(ctx, 2) keeps the step's side streams concurrent (mode 1 folds them into the caller's stream so that stage
 * times add up); mb_profile_timeline then returns, for up to `cap` recorded stage launches, the stage id and its start / end in ms
 * from the earliest recorded start -- the device timeline of the overlapped step -- and resets the recording. */
mb_status mb_profile_timeline(mb_context* ctx, int cap, int* stages, float* start_ms, float* end_ms, int* n);

/* bytes of device workspace a context currently holds (grows on demand, reused across batches) */
size_t mb_workspace_bytes(const mb_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MARIUS_B200_H_ */
