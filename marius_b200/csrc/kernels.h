// kernels.h -- internal launch interfaces between the translation units of libmarius_b200.so
#pragma once
#include "common.cuh"

namespace mb {

// storage_kernels.cu
mb_status gather_rows(const float* table, int64_t ld, int64_t d, const int64_t* idx, int64_t n, float* out, int64_t out_ld, cudaStream_t st);
mb_status scatter_rows(float* table, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* vals, int64_t vals_ld, bool add, cudaStream_t st);
mb_status adagrad_deltas(const float* grad, const float* state, int64_t n, int64_t d, int64_t ld, float lr, float* de, float* ds, cudaStream_t st);
mb_status adagrad_update_rows(float* table, float* state_table, int64_t ld, int64_t d, const int64_t* idx, int64_t n, const float* grad, int64_t grad_ld,
                              float lr, cudaStream_t st);
mb_status global_to_local_map(int64_t* map, int64_t total_rows, int64_t psize, const int32_t* part_ids, const int32_t* slots, int n_res, cudaStream_t st);
mb_status dense_adagrad_step(float* p, float* sum, const float* g, int64_t n, float lr, float eps, cudaStream_t st);

// radix_sort.cu
size_t sort_scratch_bytes(int64_t n);
template <typename K>
mb_status radix_sort_pairs(K* keys_a, K* keys_b, uint32_t* vals_a, uint32_t* vals_b, int64_t n, int key_bits, uint32_t* hist_scratch, K** keys_sorted,
                           uint32_t** vals_sorted, cudaStream_t st);
mb_status launch_i64_to_u32(const int64_t* in, uint32_t* out, int64_t n, cudaStream_t st);
mb_status segment_offsets_u32(const uint32_t* sorted_keys, int64_t n, int64_t num_keys, uint32_t* offsets, cudaStream_t st);
mb_status map_tensors_device(const int64_t* all_ids, int64_t n, int key_bits, uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                             uint32_t* flags, uint32_t* hist_scratch, uint32_t* total_scratch, int64_t* unique_out, int64_t* mapped_out,
                             int64_t* num_unique_dev, cudaStream_t st);

// decoder_kernels.cu
mb_status launch_edge_prep(const float* emb, int64_t emb_ld, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int64_t Bp,
                           int d, int decoder, float* A0, float* A1, float* pos0, float* pos1, void* A0_hi, void* A0_lo, void* A1_hi, void* A1_lo,
                           cudaStream_t st);
mb_status launch_gather_split(const float* emb, int64_t emb_ld, const int64_t* idx, int64_t n, int d, float* out, void* hi, void* lo, cudaStream_t st);
mb_status launch_split(const float* x, int64_t n, void* hi, void* lo, cudaStream_t st, int64_t cols = 0, int64_t ld_out = 0);
mb_status launch_loss_grad(float* S, const float* pos, float* gpos, float* row_loss, void* G_hi, void* G_lo, int64_t rows, int N, float w, cudaStream_t st,
                           int64_t ldg = 0);
mb_status launch_loss_reduce(const float* row_loss, int64_t n, float* loss, cudaStream_t st);
int64_t loss_merge_blocks(int64_t rows);
mb_status launch_loss_merge(const float2* stats, int slots, const float* pos, float* gpos, float* row_loss, float* zw, int64_t rows, float w,
                            float* block_loss, cudaStream_t st);
mb_status launch_edge_backward(const float* emb, int64_t emb_ld, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int d,
                               int decoder, const float* A0, const float* A1, const float* dA0, const float* dA1, const float* gpos0,
                               const float* gpos1, float* gcat, float* drel0, float* drel1, cudaStream_t st);
mb_status launch_slot_keys(const int64_t* edges, int cols, int64_t B, const int64_t* dst_negs, const int64_t* src_negs, int64_t CN, uint32_t* keys,
                           cudaStream_t st);
mb_status launch_rel_keys(const int64_t* edges, int cols, int64_t B, uint32_t* keys, cudaStream_t st);
mb_status launch_segment_reduce(int mode, const float* rows, const uint32_t* slots, const uint32_t* offsets, int64_t n_seg, int d, float* out, int64_t out_ld,
                                const float* state, int64_t state_ld, float* delta_e, float* delta_s, float* table, float* state_table, int64_t ld,
                                const int64_t* ids, float lr, cudaStream_t st);

bool decoder_vec_ok(const float* emb, int64_t emb_ld, int d, bool has_rel, const float* rel, const float* inv_rel, int sides);
mb_status launch_prep(const mb_shards* sh, cudaEvent_t rows_fetched, const float* const* row_ptrs, const float* emb, int64_t emb_ld, const int64_t* row_map, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int64_t Bp,
                      int64_t CN, int d, int decoder, int sides, const int64_t* dst_negs, const int64_t* src_negs, float* A, float* pos, void* A_hi,
                      void* A_lo, float* Neg, void* Neg_hi, void* Neg_lo, cudaStream_t st);
mb_status launch_loss(const float* S, float* G, const float* pos, float* gpos, float* row_loss, void* G_hi, void* G_lo, int64_t rows, int N, float w,
                      cudaStream_t st, int64_t ldg = 0);
mb_status launch_edge_bwd(const float* const* row_ptrs, const float* emb, int64_t emb_ld, const int64_t* row_map, const int64_t* edges, int cols, const float* rel, const float* inv_rel, int64_t B, int64_t Bp,
                          int d, int decoder, int sides, const float* A, const float* dA, const float* gpos, float* gcat, float* drel, cudaStream_t st);
mb_status launch_fetch_remote_rows(const mb_shards* sh, const int64_t* ids, int64_t U, int64_t ld, int d, float* cache, const float** row_ptrs,
                                   bool state_rows, cudaStream_t st);
mb_status launch_seg_reduce(const mb_shards* sh, int mode, const float* rows, const uint32_t* slots, const uint32_t* offsets, int64_t n_seg, int d, float* out, int64_t out_ld,
                            const float* state, int64_t state_ld, float* delta_e, float* delta_s, float* table, float* state_table, int64_t ld,
                            const int64_t* ids, float lr, cudaStream_t st, const int64_t* owner_bounds = nullptr, int part = 0);

mb_status launch_rel_reduce(const float* drel0, const float* drel1, float* out0, float* out1, const uint32_t* slots, const uint32_t* offsets, int64_t R,
                            int d, cudaStream_t st);

// sample_kernels.cu
mb_status launch_sample_negatives(int64_t num_nodes, int C, int N, int num_batch, const int64_t* edges, int64_t B, int cols, bool inverse, uint64_t seed,
                                  uint32_t batch, int64_t* out, cudaStream_t st);
mb_status launch_concat_ids(const int64_t* edges, int64_t B, int cols, const int64_t* src_negs, const int64_t* dst_negs, int64_t CN, int64_t* all_ids,
                            cudaStream_t st);
mb_status launch_split_mapped(const int64_t* mapped, const int64_t* edges, int64_t B, int cols, bool has_src_negs, int64_t CN, int64_t* edges_local,
                              int64_t* src_negs_local, int64_t* dst_negs_local, cudaStream_t st);

// eval_kernels.cu
mb_status launch_score_filter(float* scores, int64_t rows, int64_t N, int64_t ld, const int64_t* filter, int64_t F, int* bad_flag, cudaStream_t st);
mb_status launch_ranks(const float* pos, const float* neg, int64_t rows, int64_t N, int64_t ld, int64_t* ranks, cudaStream_t st);
mb_status launch_filter_sort(const int64_t* edges, int64_t E, int cols, int key_col, uint64_t* ka, uint64_t* kb, uint32_t* va, uint32_t* vb, uint32_t* hist,
                             int key_bits, int64_t* sorted_out, cudaStream_t st);
mb_status launch_filter_match(const int64_t* pool, int64_t E, int cols, int key_col, int cor_col, const int64_t* batch, int64_t B, int64_t* counts,
                              int64_t* offsets, int64_t* out, int64_t cap, int64_t* total_dev, cudaStream_t st);
mb_status launch_filter_tile(float* scores, int64_t rows, int64_t ld, int64_t t0, int64_t T, const int64_t* filter, int64_t F, int64_t num_nodes, int* bad,
                             cudaStream_t st);
mb_status launch_rank_accumulate(const float* pos, const float* neg, int64_t rows, int64_t T, int64_t ld, int64_t* ranks, cudaStream_t st);
mb_status launch_fill_i64(int64_t* p, int64_t n, int64_t v, cudaStream_t st);

// shard_kernels.cu : exchange step of the sharded table (flag barriers, owner bounds, owner-side apply of received gradient rows)
int64_t shard_exchange_bytes(int world, int64_t rows, int64_t d);
mb_status launch_shard_barrier(const mb_shards* sh, cudaStream_t st);
mb_status launch_owner_bounds(const mb_shards* sh, const int64_t* ids, int64_t n, int64_t d, int64_t* bounds, cudaStream_t st);
void shard_inbox_ptrs(const mb_shards* sh, int64_t d, int64_t** ids_out, float** rows_out);
mb_status launch_inbox_apply(const mb_shards* sh, int owner, int sender, int64_t ld, int d, float lr, int64_t max_rows, cudaStream_t st);
mb_status shard_error_flag(const mb_shards* sh, int* out);

// gemm_simt.cu
mb_status gemm_simt(const float* A, int64_t sAm, int64_t sAk, int64_t sAb, const float* B, int64_t sBk, int64_t sBn, int64_t sBb, float* C, int64_t ldc,
                    int64_t sCb, int M, int N, int K, int batches, cudaStream_t st);

// gemm_tc_group.cu : up to two contractions in one table-scheduled persistent 2-CTA launch
bool gemm_tc_supported(int64_t a_inner, int64_t b_inner);
int gemm_tc_wait_log(unsigned long long* out, int cap);  // diagnostics (MB_TC_WAITLOG=1): records of bounded waits that gave up
struct TcGroupProblem {
    const void *A_hi, *A_lo;
    int64_t lda, sAb;
    int a_mn;
    const void *B_hi, *B_lo;
    int64_t ldb, sBb;
    int b_mn;
    float* D;
    int64_t ldd, sDb;
    int M, N, K, batches;
    // optional, forward contraction: softmax statistics of every output row, one (max, sum exp(x - max)) pair per 64-column slot
    // [batches * M][stat_slots]; the buffer must be zeroed before the launch (a slot no tile wrote has sum == 0)
    float2* stats = nullptr;
    int stat_slots = 0;
    // optional, backward contractions: the A operand is not read from A_hi / A_lo but produced on the fly from the fp32 matrix
    // conv_src [batches][conv_rows][conv_cols] (leading dimension conv_ld, batch stride conv_sb):
    //   conv_mode 1: A = exp2(conv_src * log2 e - conv_z[row])  (conv_z pre-scaled by log2 e)   conv_mode 2: A = conv_src
    // a_mn == 0: A[m][k] = f(conv_src[m][k]) ; a_mn == 1: A[m][k] = f(conv_src[k][m]).  All problems of a launch must agree on conv_mode != 0.
    const float* conv_src = nullptr;
    const float* conv_z = nullptr;
    int64_t conv_ld = 0, conv_sb = 0;
    int conv_rows = 0, conv_cols = 0, conv_mode = 0;
};
constexpr int kTcStatSlotCols = 64;
mb_status gemm_tc_grouped(const TcGroupProblem* probs, int n, int passes, cudaStream_t st, bool force_smem_a = false);

}  // namespace mb
