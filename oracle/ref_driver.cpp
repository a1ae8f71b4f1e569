// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin C-ABI driver around the *unmodified* Marius reference classes (linked
// from the objects oracle/Makefile compiles out of /root/reference).  Nothing in
// the product path (marius_b200/) links or loads this; only tests/, bench.py's
// cpu_baseline / `--impl reference` arm and __graft_entry__.smoke() may use it,
// and only as the checker / CPU baseline.
//
// Every entry point drives the reference's own call sequence for the hot path
// (SURVEY.md 8a):
//   InMemory::indexRead / indexAdd                  storage.cpp:606-673
//   PartitionBuffer::indexRead / indexAdd / map     buffer.cpp:441-480,581-633
//   Model::forward_lp -> node_corrupt_forward       model.cpp:252-288, decoder_methods.cpp:57-114
//   Model::train_batch (loss, backward, Adagrad)    model.cpp:290-333, batch.cpp:62-79
//   map_tensors (unique-id mapping)                 util.cpp:180-205
// All pointers are host memory; int64 ids; fp32 values; row-major.

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "common/util.h"
#include "data/batch.h"
#include "data/graph.h"
#include "data/samplers/negative.h"
#include "nn/decoders/edge/complex.h"
#include "nn/decoders/edge/decoder_methods.h"
#include "nn/decoders/edge/distmult.h"
#include "nn/decoders/edge/transe.h"
#include "nn/encoders/encoder.h"
#include "nn/layers/embedding/embedding.h"
#include "nn/loss.h"
#include "nn/model.h"
#include "reporting/reporting.h"
#include "storage/buffer.h"
#include "storage/storage.h"

namespace {

thread_local std::string g_last_error;

torch::Tensor f32(const float *p, std::vector<int64_t> sizes) {
    return torch::from_blob(const_cast<float *>(p), sizes, torch::TensorOptions().dtype(torch::kFloat32));
}
torch::Tensor i64(const int64_t *p, std::vector<int64_t> sizes) {
    return torch::from_blob(const_cast<int64_t *>(p), sizes, torch::TensorOptions().dtype(torch::kInt64));
}
void copy_out(float *dst, const torch::Tensor &t) {
    if (dst == nullptr || !t.defined()) return;
    torch::Tensor c = t.detach().contiguous().to(torch::kFloat32);
    std::memcpy(dst, c.data_ptr<float>(), sizeof(float) * c.numel());
}

// Build the reference Model for a pure-embedding link-prediction task the way
// test/python/bindings/integration/test_nn.py:37-75 does: one EmbeddingLayer stage,
// a DistMult/ComplEx/TransE decoder in CORRUPT_NODE ("train") mode, SoftmaxCE loss.
std::shared_ptr<Model> make_model(int decoder_type, int d, int num_rel, bool inverse, int reduction, const float *rel, const float *inv_rel) {
    auto layer_config = std::make_shared<LayerConfig>();
    layer_config->type = LayerType::EMBEDDING;
    layer_config->input_dim = -1;
    layer_config->output_dim = d;
    layer_config->bias = false;
    layer_config->activation = ActivationFunction::NONE;
    auto emb_layer = std::make_shared<EmbeddingLayer>(layer_config, torch::Device(torch::kCPU));
    std::vector<std::vector<shared_ptr<Layer>>> layers = {{emb_layer}};
    auto encoder = std::make_shared<GeneralEncoder>(layers);

    auto opts = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCPU);
    shared_ptr<EdgeDecoder> decoder;
    if (decoder_type == 0) {
        decoder = std::make_shared<DistMult>(num_rel, d, opts, inverse, EdgeDecoderMethod::CORRUPT_NODE);
    } else if (decoder_type == 1) {
        decoder = std::make_shared<ComplEx>(num_rel, d, opts, inverse, EdgeDecoderMethod::CORRUPT_NODE);
    } else {
        decoder = std::make_shared<TransE>(num_rel, d, opts, inverse, EdgeDecoderMethod::CORRUPT_NODE);
    }
    {
        torch::NoGradGuard ng;
        if (rel != nullptr) decoder->relations_.copy_(f32(rel, {num_rel, d}));
        if (inverse && inv_rel != nullptr) decoder->inverse_relations_.copy_(f32(inv_rel, {num_rel, d}));
    }

    auto loss_opts = std::make_shared<LossOptions>();
    loss_opts->loss_reduction = reduction == 0 ? LossReduction::MEAN : LossReduction::SUM;
    auto loss = std::make_shared<SoftmaxCrossEntropy>(loss_opts);

    auto model = std::make_shared<Model>(encoder, std::dynamic_pointer_cast<Decoder>(decoder), loss);
    return model;
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_last_error.c_str(); }

int ref_num_threads() { return torch::get_num_threads(); }
void ref_set_num_threads(int n) { torch::set_num_threads(n); }

// InMemory::indexRead (storage.cpp:606-649), CPU table.
int ref_inmemory_index_read(const float *table, int64_t rows, int64_t d, const int64_t *idx, int64_t n, float *out) {
    try {
        InMemory storage(f32(table, {rows, d}));
        torch::Tensor r = storage.indexRead(i64(idx, {n}));
        copy_out(out, r);
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

// Same call but with a 2-D index tensor, to pin the reference's error convention
// (std::runtime_error on rank != 1): returns 1 when the reference throws.
int ref_inmemory_index_read_bad_rank(const float *table, int64_t rows, int64_t d) {
    try {
        InMemory storage(f32(table, {rows, d}));
        storage.indexRead(torch::zeros({2, 2}, torch::kInt64));
        return 0;
    } catch (const std::runtime_error &e) {
        return 1;
    }
}

// InMemory::indexAdd (storage.cpp:651-673), CPU table, in place.
int ref_inmemory_index_add(float *table, int64_t rows, int64_t d, const int64_t *idx, int64_t n, const float *vals, int64_t vals_rows, int64_t vals_cols) {
    try {
        InMemory storage(f32(table, {rows, d}));
        storage.indexAdd(i64(idx, {n}), f32(vals, {vals_rows, vals_cols}));
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

// PartitionBuffer over a flat fp32 file (buffer.cpp): load the given buffer state,
// indexRead / indexAdd with buffer-local ids, return both global->local maps, then
// (optionally) perform swaps and report admit/evict ids.  `states` is [num_states, capacity].
int ref_partition_buffer_exercise(const char *filename, int capacity, int num_partitions, int64_t partition_size, int d, int64_t total, const int64_t *states,
                                  int num_states, const int64_t *idx, int64_t n, float *read_out, const float *add_vals, float *read_after_add,
                                  int64_t *map_current, int64_t *map_next, int64_t *admits /*[num_states-1]*/, int64_t *evicts /*[num_states-1]*/) {
    try {
        PartitionBuffer pb(capacity, num_partitions, 1, partition_size, d, total, torch::kFloat32, std::string(filename), false);
        std::vector<torch::Tensor> st;
        for (int i = 0; i < num_states; i++) st.emplace_back(i64(states + (int64_t)i * capacity, {capacity}).clone());
        pb.setBufferOrdering(st);
        pb.load();
        torch::Tensor ids = i64(idx, {n});
        copy_out(read_out, pb.indexRead(ids));
        if (map_current) std::memcpy(map_current, pb.getGlobalToLocalMap(true).data_ptr<int64_t>(), sizeof(int64_t) * total);
        if (map_next && num_states > 1) std::memcpy(map_next, pb.getGlobalToLocalMap(false).data_ptr<int64_t>(), sizeof(int64_t) * total);
        if (add_vals) {
            pb.indexAdd(ids, f32(add_vals, {n, (int64_t)d}));
            copy_out(read_after_add, pb.indexRead(ids));
        }
        for (int s = 0; s + 1 < num_states; s++) {
            auto a = pb.getNextAdmit();
            auto e = pb.getNextEvict();
            if (admits) admits[s] = a.empty() ? -1 : a[0];
            if (evicts) evicts[s] = e.empty() ? -1 : e[0];
            pb.performNextSwap();
        }
        pb.unload(true);
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

// map_tensors (util.cpp:180-205) over cat(src, dst, src_negs, dst_negs) the way
// DataLoader::edgeSample does (dataloader.cpp:399-468).  Returns U.
int64_t ref_map_tensors(const int64_t *all_ids, int64_t n, int64_t *unique_out, int64_t *mapped_out) {
    auto tup = map_tensors({i64(all_ids, {n})});
    torch::Tensor u = std::get<0>(tup);
    torch::Tensor m = std::get<1>(tup)[0].contiguous();
    std::memcpy(unique_out, u.data_ptr<int64_t>(), sizeof(int64_t) * u.numel());
    std::memcpy(mapped_out, m.data_ptr<int64_t>(), sizeof(int64_t) * n);
    return u.numel();
}

// One reference training batch on batch-local tensors:
//   forward_lp (scores) ; train_batch (loss.backward + Batch::accumulateGradients).
// Shapes: emb,state [U,d]; edges [B,3] local ids; dst_negs/src_negs [C,N] local ids
// (src_negs == NULL or inv_rel == NULL -> no inverse side).  Outputs (any may be NULL):
//   pos/inv_pos [Bp], neg/inv_neg [Bp,N] with Bp = C*ceil(B/C) (decoder_methods.cpp:103-111);
//   loss [1]; grad [U,d] = dLoss/d emb; delta_e, delta_s [U,d] (batch.cpp:62-79);
//   rel_grad, inv_rel_grad [R,d].
int ref_train_batch(int decoder_type, int d, int num_rel, const float *rel, const float *inv_rel, const float *emb, const float *state, int64_t U,
                    const int64_t *edges, int64_t B, const int64_t *dst_negs, const int64_t *src_negs, int C, int N, float lr, int reduction, float *pos,
                    float *neg, float *inv_pos, float *inv_neg, float *loss_out, float *grad, float *delta_e, float *delta_s, float *rel_grad,
                    float *inv_rel_grad) {
    try {
        bool inverse = (inv_rel != nullptr) && (src_negs != nullptr);
        auto model = make_model(decoder_type, d, num_rel, inverse, reduction, rel, inv_rel);
        model->sparse_lr_ = lr;
        auto decoder = std::dynamic_pointer_cast<EdgeDecoder>(model->decoder_);

        auto batch = std::make_shared<Batch>(true);
        batch->node_embeddings_ = f32(emb, {U, (int64_t)d}).clone();
        batch->node_embeddings_state_ = f32(state, {U, (int64_t)d}).clone();
        batch->edges_ = i64(edges, {B, 3}).clone();
        batch->dst_neg_indices_mapping_ = i64(dst_negs, {C, N}).clone();
        if (inverse) batch->src_neg_indices_mapping_ = i64(src_negs, {C, N}).clone();

        {
            torch::NoGradGuard ng;
            auto scores = model->forward_lp(batch, true);
            copy_out(pos, std::get<0>(scores));
            copy_out(neg, std::get<1>(scores));
            copy_out(inv_pos, std::get<2>(scores));
            copy_out(inv_neg, std::get<3>(scores));
            if (loss_out) {
                torch::Tensor l = (*model->loss_function_)(std::get<0>(scores), std::get<1>(scores), true);
                if (std::get<3>(scores).defined()) l = (*model->loss_function_)(std::get<2>(scores), std::get<3>(scores), true) + l;
                loss_out[0] = l.item<float>();
            }
        }

        model->train_batch(batch, true);

        copy_out(grad, batch->node_embeddings_.grad());
        copy_out(delta_e, batch->node_gradients_);
        copy_out(delta_s, batch->node_state_update_);
        copy_out(rel_grad, decoder->relations_.grad());
        if (inverse) copy_out(inv_rel_grad, decoder->inverse_relations_.grad());
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

// Model::evaluate_batch (model.cpp:335-349) on batch-local tensors: forward_lp with the batch's score filters
// (model.cpp:279-285 -> negative.cpp:306-311) and LinkPredictionReporter::addResult -> computeRanks
// (reporting.cpp:56-72) once per corruption side.  dst_filter / src_filter: [F,2] int64 (row, column) pairs or NULL.
// Outputs: ranks / inv_ranks [Bp] int64 in the order the reporter collects them; neg / inv_neg [Bp,N] the filtered scores (may be NULL).
int ref_evaluate_batch(int decoder_type, int d, int num_rel, const float *rel, const float *inv_rel, const float *emb, int64_t U, const int64_t *edges,
                       int64_t B, const int64_t *dst_negs, const int64_t *src_negs, int C, int N, const int64_t *dst_filter, int64_t Fd,
                       const int64_t *src_filter, int64_t Fs, int64_t *ranks, int64_t *inv_ranks, float *pos, float *neg, float *inv_pos, float *inv_neg) {
    try {
        bool inverse = (inv_rel != nullptr) && (src_negs != nullptr);
        auto model = make_model(decoder_type, d, num_rel, inverse, 1, rel, inv_rel);
        auto reporter = std::make_shared<LinkPredictionReporter>();
        model->reporter_ = reporter;
        auto batch = std::make_shared<Batch>(false);
        batch->node_embeddings_ = f32(emb, {U, (int64_t)d}).clone();
        batch->edges_ = i64(edges, {B, 3}).clone();
        batch->dst_neg_indices_mapping_ = i64(dst_negs, {C, N}).clone();
        if (inverse) batch->src_neg_indices_mapping_ = i64(src_negs, {C, N}).clone();
        if (dst_filter && Fd > 0) batch->dst_neg_filter_ = i64(dst_filter, {Fd, 2}).clone();
        if (inverse && src_filter && Fs > 0) batch->src_neg_filter_ = i64(src_filter, {Fs, 2}).clone();
        torch::NoGradGuard ng;
        model->evaluate_batch(batch);
        if (reporter->per_batch_ranks_.size() != (inverse ? 2u : 1u)) throw std::runtime_error("unexpected number of rank tensors");
        torch::Tensor r0 = reporter->per_batch_ranks_[0].to(torch::kInt64).contiguous();
        std::memcpy(ranks, r0.data_ptr<int64_t>(), sizeof(int64_t) * r0.numel());
        if (inverse) {
            torch::Tensor r1 = reporter->per_batch_ranks_[1].to(torch::kInt64).contiguous();
            std::memcpy(inv_ranks, r1.data_ptr<int64_t>(), sizeof(int64_t) * r1.numel());
        }
        if (pos || neg || inv_pos || inv_neg) {
            auto scores = model->forward_lp(batch, false);
            copy_out(pos, std::get<0>(scores));
            copy_out(neg, std::get<1>(scores));
            copy_out(inv_pos, std::get<2>(scores));
            copy_out(inv_neg, std::get<3>(scores));
        }
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

// compute_filter_corruption_cpu (negative.cpp:62-195).  graph_edges != NULL: global ("filtered" evaluation) filter against a graph whose
// edge list is graph_edges [G,3] (sorted here by source / destination exactly as MariusGraph's callers do, with a stable argsort);
// graph_edges == NULL: local filter against the batch.  Writes up to cap (row, column) pairs into out and returns the number of pairs
// (or -1 on error; a count > cap means the buffer was too small).
int64_t ref_compute_filter(const int64_t *edges, int64_t B, int cols, const int64_t *corruption_nodes, int C, int N, int inverse,
                           const int64_t *graph_edges, int64_t G, int64_t num_nodes, int64_t *out, int64_t cap) {
    try {
        torch::Tensor e = i64(edges, {B, (int64_t)cols}).clone();
        torch::Tensor negs = i64(corruption_nodes, {C, N}).clone();
        shared_ptr<MariusGraph> graph = nullptr;
        bool global = graph_edges != nullptr;
        if (global) {
            torch::Tensor ge = i64(graph_edges, {G, (int64_t)cols}).clone();
            torch::Tensor by_src = ge.index_select(0, ge.select(1, 0).argsort(true));
            torch::Tensor by_dst = ge.index_select(0, ge.select(1, -1).argsort(true));
            graph = std::make_shared<MariusGraph>(by_src, by_dst, num_nodes);
            graph->all_src_sorted_edges_ = by_src;
            graph->all_dst_sorted_edges_ = by_dst;
        }
        torch::Tensor f = compute_filter_corruption_cpu(graph, e, negs, inverse != 0, global, LocalFilterMode::ALL, torch::Tensor()).contiguous();
        int64_t n = f.size(0);
        if (n <= cap && n > 0) std::memcpy(out, f.data_ptr<int64_t>(), sizeof(int64_t) * 2 * n);
        return n;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return -1;
    }
}

// RankingMetric::computeMetric (reporting.cpp:17-31) on a rank vector: out = {mean rank, MRR, Hits@1, Hits@3, Hits@10}
int ref_ranking_metrics(const int64_t *ranks, int64_t n, double *out) {
    try {
        torch::Tensor r = i64(ranks, {n}).clone();
        out[0] = MeanRankMetric().computeMetric(r).item<double>();
        out[1] = MeanReciprocalRankMetric().computeMetric(r).item<double>();
        out[2] = HitskMetric(1).computeMetric(r).item<double>();
        out[3] = HitskMetric(3).computeMetric(r).item<double>();
        out[4] = HitskMetric(10).computeMetric(r).item<double>();
        return 0;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return 1;
    }
}

// The reference's synchronous CPU hot loop (trainer.cpp:106-138 with the table in host
// memory): per batch  indexRead x2 -> train_batch -> indexAdd x2, over `num_batches`
// pre-built batches (sampling and unique-mapping are inputs).  The dense
// relation optimizer is the reference's own Adagrad (optim.cpp) when dense_lr > 0.
// Returns elapsed seconds for the timed loop, or -1 on error.
//   uniq   : concatenated unique id lists, uniq_off[num_batches+1]
//   edges  : [num_batches, B, 3] batch-local; dst_negs/src_negs : [num_batches, C, N] batch-local
double ref_train_loop(int decoder_type, int d, int num_rel, float *table, float *state_table, int64_t num_nodes, int num_batches, const int64_t *uniq,
                      const int64_t *uniq_off, const int64_t *edges, int64_t B, const int64_t *dst_negs, const int64_t *src_negs, int C, int N, float lr,
                      int reduction, int warmup_batches) {
    try {
        auto model = make_model(decoder_type, d, num_rel, src_negs != nullptr, reduction, nullptr, nullptr);
        model->sparse_lr_ = lr;
        InMemory emb_storage(f32(table, {num_nodes, (int64_t)d}));
        InMemory state_storage(f32(state_table, {num_nodes, (int64_t)d}));

        auto run = [&](int b) {
            auto batch = std::make_shared<Batch>(true);
            batch->unique_node_indices_ = i64(uniq + uniq_off[b], {uniq_off[b + 1] - uniq_off[b]});
            batch->edges_ = i64(edges + (int64_t)b * B * 3, {B, 3});
            batch->dst_neg_indices_mapping_ = i64(dst_negs + (int64_t)b * C * N, {C, N});
            if (src_negs) batch->src_neg_indices_mapping_ = i64(src_negs + (int64_t)b * C * N, {C, N});
            // DataLoader::loadCPUParameters (dataloader.cpp:505-527)
            batch->node_embeddings_ = emb_storage.indexRead(batch->unique_node_indices_);
            batch->node_embeddings_state_ = state_storage.indexRead(batch->unique_node_indices_);
            // ComputeWorkerCPU::run (pipeline_cpu.cpp:10-36)
            model->train_batch(batch, true);
            // DataLoader::updateEmbeddings (dataloader.cpp:550-564)
            emb_storage.indexAdd(batch->unique_node_indices_, batch->node_gradients_);
            state_storage.indexAdd(batch->unique_node_indices_, batch->node_state_update_);
        };
        for (int b = 0; b < warmup_batches && b < num_batches; b++) run(b);
        auto t0 = std::chrono::steady_clock::now();
        for (int b = 0; b < num_batches; b++) run(b);
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return -1.0;
    }
}

}  // extern "C"
