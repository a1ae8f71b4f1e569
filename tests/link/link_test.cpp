// tests/link/link_test.cpp -- TEST INFRASTRUCTURE: the drop-in boundary, linked against the reference itself.
//
// Compiled (oracle/Makefile, target `link`) against the UNMODIFIED reference headers and linked with oracle/_ref/libmarius_ref.so
// (every reference TU) and marius_b200/lib/libmarius_b200.so (the C ABI).  B200Table below derives from the REFERENCE's own abstract
// `Storage` (storage/storage.h:35-86) -- it is not a re-declaration -- and keeps its rows in HBM, moving them with mb_gather_rows /
// mb_scatter_add_rows.  The reference's GraphModelStorage (graph_storage.cpp:206-333) and DataLoader::loadGPUParameters /
// updateEmbeddings (dataloader.cpp:529-564) then run unchanged on top of it, in the call sequence of ComputeWorkerGPU::run
// (pipeline_gpu.cpp:49-91) / SynchronousTrainer::train (trainer.cpp:106-138):
//     loadGPUParameters(batch) -> train_batch -> updateEmbeddings(batch, gpu = true)
// with the compute in the middle done either by mb_train_batch (the C ABI's Model::train_batch) or by the reference's own
// Model::train_batch on the CPU for comparison.
#include <c10/cuda/CUDAStream.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "data/batch.h"
#include "data/dataloader.h"
#include "nn/decoders/edge/complex.h"
#include "nn/encoders/encoder.h"
#include "nn/layers/embedding/embedding.h"
#include "nn/loss.h"
#include "nn/model.h"
#include "storage/graph_storage.h"
#include "storage/storage.h"

#include "marius_b200.h"

namespace {

thread_local std::string g_err;

void check(int status) {
    if (status == MB_OK) return;
    std::string msg = mb_last_error();
    if (status == MB_ERR_INVALID) throw std::runtime_error(msg);  // the reference's convention (storage.cpp:607-610, 652-655)
    throw MariusRuntimeException(msg);
}

void* cur_stream(const torch::Device& dev) { return (void*)c10::cuda::getCurrentCUDAStream(dev.index()).stream(); }

// A device-resident table behind the reference's Storage interface (the DEVICE_MEMORY backend, storage.cpp:488-775)
class B200Table : public Storage {
   public:
    explicit B200Table(torch::Tensor data) {
        data_ = data;
        dim0_size_ = data.size(0);
        dim1_size_ = data.size(1);
        dtype_ = torch::kFloat32;
        device_ = torch::kCUDA;  // dataloader.cpp:507,531,552,558 route on this
        initialized_ = true;
        filename_ = "";
    }
    torch::Tensor indexRead(Indices indices) override {
        if (indices.sizes().size() != 1) throw std::runtime_error("");
        auto idx = indices.to(data_.device()).to(torch::kInt64).contiguous();
        auto out = torch::empty({idx.size(0), dim1_size_}, data_.options());
        check(mb_gather_rows(data_.data_ptr<float>(), dim0_size_, data_.stride(0), dim1_size_, idx.data_ptr<int64_t>(), idx.size(0), out.data_ptr<float>(),
                             dim1_size_, cur_stream(data_.device())));
        return out;
    }
    void indexAdd(Indices indices, torch::Tensor values) override {
        if (!values.defined() || indices.sizes().size() != 1 || indices.size(0) != values.size(0) || dim1_size_ != values.size(1)) throw std::runtime_error("");
        auto idx = indices.to(data_.device()).to(torch::kInt64).contiguous();
        auto v = values.to(data_.device()).contiguous();
        check(mb_scatter_add_rows(data_.data_ptr<float>(), dim0_size_, data_.stride(0), dim1_size_, idx.data_ptr<int64_t>(), idx.size(0), v.data_ptr<float>(),
                                  v.stride(0), cur_stream(data_.device())));
    }
    torch::Tensor range(int64_t offset, int64_t n) override { return data_.narrow(0, offset, n); }
    void indexPut(Indices indices, torch::Tensor values) override {
        auto idx = indices.to(data_.device()).to(torch::kInt64).contiguous();
        auto v = values.to(data_.device()).contiguous();
        check(mb_scatter_put_rows(data_.data_ptr<float>(), dim0_size_, data_.stride(0), dim1_size_, idx.data_ptr<int64_t>(), idx.size(0), v.data_ptr<float>(),
                                  v.stride(0), cur_stream(data_.device())));
    }
    void rangePut(int64_t offset, int64_t n, torch::Tensor values) override { data_.narrow(0, offset, n).copy_(values); }
    void load() override {}
    void write() override {}
    void unload(bool) override {}
    void shuffle() override { throw std::runtime_error(""); }
    void sort(bool) override { throw std::runtime_error(""); }
};

torch::Tensor f32(const float* p, std::vector<int64_t> sizes) { return torch::from_blob(const_cast<float*>(p), sizes, torch::kFloat32).clone(); }
torch::Tensor i64(const int64_t* p, std::vector<int64_t> sizes) { return torch::from_blob(const_cast<int64_t*>(p), sizes, torch::kInt64).clone(); }

shared_ptr<DataLoader> make_loader(shared_ptr<GraphModelStorage> gms, int C, int N) {
    auto tc = std::make_shared<TrainingConfig>();
    tc->batch_size = 1000;
    tc->negative_sampling = std::make_shared<NegativeSamplingConfig>();
    tc->negative_sampling->num_chunks = C;
    tc->negative_sampling->negatives_per_positive = N;
    tc->negative_sampling->degree_fraction = 0;
    tc->negative_sampling->filtered = false;
    tc->negative_sampling->local_filter_mode = LocalFilterMode::DEG;
    auto ec = std::make_shared<EvaluationConfig>();
    ec->batch_size = 1000;
    ec->negative_sampling = tc->negative_sampling;
    return std::make_shared<DataLoader>(gms, LearningTask::LINK_PREDICTION, tc, ec, nullptr);
}

}  // namespace

extern "C" {

const char* link_last_error() { return g_err.c_str(); }

// `steps` batches through  DataLoader::loadGPUParameters -> mb_train_batch -> DataLoader::updateEmbeddings  on a B200Table pair (cuda:0),
// every call into storage going through the reference's GraphModelStorage.  table / state [num_nodes, d] are updated in place (host
// memory in, host memory out); uniq_off [steps + 1] delimits the batches' unique-id lists.  Returns 0, or 1 with link_last_error().
int link_train_loop(float* table, float* state, int64_t num_nodes, int d, int num_rel, const float* rel, const float* inv_rel, int steps, const int64_t* uniq,
                    const int64_t* uniq_off, const int64_t* edges, int64_t B, const int64_t* dst_negs, const int64_t* src_negs, int C, int N, float lr,
                    float* losses) {
    try {
        auto dev = torch::Device(torch::kCUDA, 0);
        auto emb_t = f32(table, {num_nodes, d}).to(dev);
        auto st_t = f32(state, {num_nodes, d}).to(dev);
        GraphModelStoragePtrs ptrs;
        ptrs.node_embeddings = std::make_shared<B200Table>(emb_t);
        ptrs.node_optimizer_state = std::make_shared<B200Table>(st_t);
        ptrs.edges = std::make_shared<InMemory>(i64(edges, {B, 3}));  // the reference's own host InMemory (graph_storage.cpp:76 reads its size)
        auto gms = std::make_shared<GraphModelStorage>(ptrs, false);  // the reference's facade, unmodified
        auto loader = make_loader(gms, C, N);                         // the reference's DataLoader, unmodified
        auto rel_d = f32(rel, {num_rel, d}).to(dev), inv_d = f32(inv_rel, {num_rel, d}).to(dev);
        mb_context* ctx = nullptr;
        check(mb_create(0, &ctx));
        for (int s = 0; s < steps; s++) {
            auto batch = std::make_shared<Batch>(true);
            const int64_t U = uniq_off[s + 1] - uniq_off[s];
            batch->unique_node_indices_ = i64(uniq + uniq_off[s], {U}).to(dev);
            batch->edges_ = i64(edges + (int64_t)s * B * 3, {B, 3}).to(dev);
            batch->dst_neg_indices_mapping_ = i64(dst_negs + (int64_t)s * C * N, {C, N}).to(dev);
            batch->src_neg_indices_mapping_ = i64(src_negs + (int64_t)s * C * N, {C, N}).to(dev);
            loader->loadGPUParameters(batch);  // -> GraphModelStorage::getNodeEmbeddings / getNodeEmbeddingState -> B200Table::indexRead
            if (!batch->node_embeddings_.defined() || !batch->node_embeddings_state_.defined() || !batch->node_embeddings_.is_cuda())
                throw std::runtime_error("loadGPUParameters did not fill the batch from the device tables");
            // Model::train_batch through the C ABI: leaves node_gradients_ (delta_e) and node_state_update_ (delta_s) (batch.cpp:62-79)
            mb_batch mb;
            std::memset(&mb, 0, sizeof(mb));
            mb.decoder = MB_DECODER_COMPLEX;
            mb.U = U;
            mb.d = d;
            mb.B = B;
            mb.R = num_rel;
            mb.C = C;
            mb.N = N;
            mb.edges = batch->edges_.data_ptr<int64_t>();
            mb.edge_cols = 3;
            mb.dst_negs = batch->dst_neg_indices_mapping_.data_ptr<int64_t>();
            mb.src_negs = batch->src_neg_indices_mapping_.data_ptr<int64_t>();
            mb.rel = rel_d.data_ptr<float>();
            mb.inv_rel = inv_d.data_ptr<float>();
            auto loss = torch::zeros({1}, emb_t.options());
            batch->node_gradients_ = torch::empty({U, d}, emb_t.options());
            batch->node_state_update_ = torch::empty({U, d}, emb_t.options());
            check(mb_train_batch(ctx, &mb, batch->node_embeddings_.data_ptr<float>(), d, batch->node_embeddings_state_.data_ptr<float>(), d, lr, MB_REDUCTION_SUM,
                                 MB_PREC_BF16X3, loss.data_ptr<float>(), nullptr, batch->node_gradients_.data_ptr<float>(),
                                 batch->node_state_update_.data_ptr<float>(), nullptr, nullptr, cur_stream(dev)));
            loader->updateEmbeddings(batch, true);  // -> GraphModelStorage::updateAddNodeEmbeddings / ...State -> B200Table::indexAdd
            losses[s] = loss.item<float>();
        }
        std::memcpy(table, emb_t.cpu().contiguous().data_ptr<float>(), sizeof(float) * num_nodes * d);
        std::memcpy(state, st_t.cpu().contiguous().data_ptr<float>(), sizeof(float) * num_nodes * d);
        mb_destroy(ctx);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}

// error conventions through the reference's facade: bad index rank / value shape -> std::runtime_error (test_buffer.cpp:282,294-296)
int link_error_conventions() {
    try {
        auto dev = torch::Device(torch::kCUDA, 0);
        GraphModelStoragePtrs ptrs;
        ptrs.node_embeddings = std::make_shared<B200Table>(torch::zeros({16, 8}, torch::TensorOptions().device(dev)));
        ptrs.node_optimizer_state = std::make_shared<B200Table>(torch::zeros({16, 8}, torch::TensorOptions().device(dev)));
        ptrs.edges = std::make_shared<InMemory>(torch::zeros({4, 3}, torch::kInt64));
        GraphModelStorage gms(ptrs, false);
        int caught = 0;
        try {
            gms.getNodeEmbeddings(torch::zeros({2, 2}, torch::kInt64));
        } catch (const std::runtime_error&) {
            caught++;
        }
        try {
            gms.updateAddNodeEmbeddings(torch::arange(4), torch::zeros({5, 8}));
        } catch (const std::runtime_error&) {
            caught++;
        }
        try {
            gms.updateAddNodeEmbeddings(torch::arange(4), torch::zeros({4, 9}));
        } catch (const std::runtime_error&) {
            caught++;
        }
        return caught;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
