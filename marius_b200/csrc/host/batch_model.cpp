// batch_model.cpp -- Batch (data/batch.cpp) and Model (nn/model.cpp) for link prediction on the fused kernels.
#include "marius_host.h"

// ---- Batch -----------------------------------------------------------------------------------------------------
Batch::Batch(bool train) : train_(train) { clear(); }

void Batch::to(torch::Device device) {
    // batch.cpp:21-60: every defined tensor moves to the device (non_blocking from pinned memory)
    auto mv = [&](torch::Tensor& t) {
        if (t.defined()) t = t.to(device, /*non_blocking=*/true);
    };
    mv(edges_);
    mv(unique_node_indices_);
    mv(src_neg_indices_mapping_);
    mv(dst_neg_indices_mapping_);
    mv(src_neg_filter_);
    mv(dst_neg_filter_);
    mv(node_embeddings_);
    mv(node_embeddings_state_);
    device_id_ = device.is_cuda() ? device.index() : -1;
}

void Batch::accumulateGradients(float learning_rate) {
    // batch.cpp:62-79 -- one kernel instead of five elementwise passes; same fp32 operation order
    if (node_embeddings_.defined()) {
        auto g = node_embeddings_.grad();
        if (!g.defined()) throw UndefinedTensorException();
        if (!node_embeddings_state_.defined()) throw UndefinedTensorException();
        g = g.contiguous();
        auto s = node_embeddings_state_.contiguous();
        int64_t d = g.dim() > 1 ? g.size(-1) : g.numel();
        int64_t n = g.numel() / std::max<int64_t>(d, 1);
        node_gradients_ = torch::empty_like(g);
        node_state_update_ = torch::empty_like(g);
        mb_throw_on_error(mb_adagrad_deltas(g.data_ptr<float>(), s.data_ptr<float>(), n, d, d, learning_rate, node_gradients_.data_ptr<float>(),
                                            node_state_update_.data_ptr<float>(), mb_current_stream(g.device())));
    }
    node_embeddings_state_ = torch::Tensor();
}

void Batch::embeddingsToHost() {
    // batch.cpp:81-103
    auto pin = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCPU).pinned_memory(true);
    if (node_gradients_.defined() && node_gradients_.device().is_cuda()) {
        auto a = torch::empty(node_gradients_.sizes(), pin), b = torch::empty(node_state_update_.sizes(), pin);
        a.copy_(node_gradients_, true);
        b.copy_(node_state_update_, true);
        node_gradients_ = a;
        node_state_update_ = b;
    }
    if (unique_node_indices_.defined()) unique_node_indices_ = unique_node_indices_.to(torch::kCPU);
    if (torch::cuda::is_available()) torch::cuda::synchronize();
}

void Batch::clear() {
    unique_node_indices_ = node_embeddings_ = node_gradients_ = node_state_update_ = node_embeddings_state_ = torch::Tensor();
    src_neg_indices_mapping_ = dst_neg_indices_mapping_ = edges_ = src_neg_indices_ = dst_neg_indices_ = torch::Tensor();
    src_neg_filter_ = dst_neg_filter_ = torch::Tensor();
}

// ---- Model -----------------------------------------------------------------------------------------------------
Model::Model(shared_ptr<EdgeDecoder> decoder, shared_ptr<LossFunction> loss, torch::Device device) : decoder_(decoder), loss_function_(loss), device_(device) {
    if (decoder_ != nullptr) register_module("decoder", std::dynamic_pointer_cast<torch::nn::Module>(decoder_));
}

static torch::Tensor apply_score_filter(torch::Tensor scores, torch::Tensor filter) {
    if (filter.defined()) scores.index_put_({filter.select(1, 0), filter.select(1, 1)}, -1e9);  // negative.cpp:306-311
    return scores;
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> Model::forward_lp(shared_ptr<Batch> batch, bool train) {
    (void)train;
    torch::Tensor pos, neg, inv_pos, inv_neg;
    if (decoder_->decoder_method_ == EdgeDecoderMethod::ONLY_POS) {
        std::tie(pos, inv_pos) = only_pos_forward(decoder_, batch->edges_, batch->node_embeddings_);
    } else if (decoder_->decoder_method_ == EdgeDecoderMethod::CORRUPT_NODE) {
        std::tie(pos, neg, inv_pos, inv_neg) =
            node_corrupt_forward(decoder_, batch->edges_, batch->node_embeddings_, batch->dst_neg_indices_mapping_, batch->src_neg_indices_mapping_);
    } else {
        throw MariusRuntimeException("Decoder method currently unsupported.");  // model.cpp:266-275
    }
    if (neg.defined()) neg = apply_score_filter(neg, batch->dst_neg_filter_);
    if (inv_neg.defined()) inv_neg = apply_score_filter(inv_neg, batch->src_neg_filter_);
    return std::forward_as_tuple(pos, neg, inv_pos, inv_neg);
}

void Model::clear_grad() {
    for (auto& p : parameters()) {
        if (p.grad().defined()) p.mutable_grad() = torch::Tensor();
    }
}

void Model::step() {
    // dense Adagrad on the relation tables (AdagradOptimizer::step, optim.cpp:114-145)
    auto params = parameters();
    if (dense_state_.size() != params.size()) {
        dense_state_.clear();
        for (auto& p : params) dense_state_.push_back(torch::zeros_like(p));
    }
    torch::NoGradGuard ng;
    for (size_t i = 0; i < params.size(); i++) {
        auto& p = params[i];
        if (!p.grad().defined()) continue;
        auto g = p.grad().contiguous();
        mb_throw_on_error(mb_dense_adagrad_step(p.data_ptr<float>(), dense_state_[i].data_ptr<float>(), g.data_ptr<float>(), p.numel(), dense_lr_, 1e-10f,
                                                mb_current_stream(p.device())));
    }
}

static bool fusable(const Model& m, const Batch& b) {
    return std::dynamic_pointer_cast<SoftmaxCrossEntropy>(m.loss_function_) != nullptr && m.decoder_->decoder_method_ == EdgeDecoderMethod::CORRUPT_NODE &&
           !b.src_neg_filter_.defined() && !b.dst_neg_filter_.defined() && b.dst_neg_indices_mapping_.defined();
}

static mb_batch describe(const Model& m, const Batch& b, int64_t U, int64_t d, torch::Tensor& edges, torch::Tensor& dn, torch::Tensor& sn) {
    edges = b.edges_.to(torch::kInt64).contiguous();
    dn = b.dst_neg_indices_mapping_.to(torch::kInt64).contiguous();
    sn = b.src_neg_indices_mapping_.defined() ? b.src_neg_indices_mapping_.to(torch::kInt64).contiguous() : torch::Tensor();
    mb_batch mbb;
    bool has_rel = edges.size(1) == 3;
    bool inverse = has_rel && m.decoder_->use_inverse_relations_ && sn.defined();
    mbb.decoder = has_rel ? m.decoder_->decoder_kind_ : MB_DECODER_DOT;
    mbb.U = U;
    mbb.d = d;
    mbb.B = edges.size(0);
    mbb.R = has_rel ? m.decoder_->relations_.size(0) : 0;
    mbb.C = (int)dn.size(0);
    mbb.N = (int)dn.size(1);
    mbb.edges = edges.data_ptr<int64_t>();
    mbb.edge_cols = (int)edges.size(1);
    mbb.dst_negs = dn.data_ptr<int64_t>();
    mbb.src_negs = inverse ? sn.data_ptr<int64_t>() : nullptr;
    mbb.rel = has_rel ? m.decoder_->relations_.data_ptr<float>() : nullptr;
    mbb.inv_rel = inverse ? m.decoder_->inverse_relations_.data_ptr<float>() : nullptr;
    return mbb;
}

static void set_grad(torch::Tensor& param, torch::Tensor g) {
    if (param.grad().defined())
        param.mutable_grad() = param.grad() + g;  // call_step == false accumulates, like autograd
    else
        param.mutable_grad() = g;
}

void Model::train_batch(shared_ptr<Batch> batch, bool call_step) {
    // model.cpp:290-333
    if (call_step) clear_grad();
    if (!batch->node_embeddings_.defined()) throw UndefinedTensorException();
    if (fusable(*this, *batch) && batch->node_embeddings_state_.defined()) {
        // fused path: forward + SoftmaxCE + backward + Batch::accumulateGradients in one C-ABI call
        auto emb = batch->node_embeddings_.detach().contiguous();
        auto state = batch->node_embeddings_state_.contiguous();
        torch::Tensor edges, dn, sn;
        mb_batch mbb = describe(*this, *batch, emb.size(0), emb.size(1), edges, dn, sn);
        auto grad = torch::empty_like(emb), de = torch::empty_like(emb), ds = torch::empty_like(emb);
        auto loss = torch::empty({1}, emb.options());
        torch::Tensor rg, irg;
        if (mbb.rel) rg = torch::empty_like(decoder_->relations_);
        if (mbb.inv_rel) irg = torch::empty_like(decoder_->inverse_relations_);
        auto red = std::dynamic_pointer_cast<SoftmaxCrossEntropy>(loss_function_)->reduction_type_;
        mb_throw_on_error(mb_train_batch(mb_context_for(emb.device()), &mbb, emb.data_ptr<float>(), emb.stride(0), state.data_ptr<float>(), state.stride(0),
                                         sparse_lr_, red == LossReduction::SUM ? MB_REDUCTION_SUM : MB_REDUCTION_MEAN, mb_default_precision(),
                                         loss.data_ptr<float>(), grad.data_ptr<float>(), de.data_ptr<float>(), ds.data_ptr<float>(),
                                         mbb.rel ? rg.data_ptr<float>() : nullptr, mbb.inv_rel ? irg.data_ptr<float>() : nullptr,
                                         mb_current_stream(emb.device())));
        batch->node_embeddings_.requires_grad_();
        batch->node_embeddings_.mutable_grad() = grad;
        if (rg.defined()) set_grad(decoder_->relations_, rg);
        if (irg.defined()) set_grad(decoder_->inverse_relations_, irg);
        if (call_step) step();
        batch->node_gradients_ = de;        // what Batch::accumulateGradients leaves behind (batch.cpp:62-79)
        batch->node_state_update_ = ds;
        batch->node_embeddings_state_ = torch::Tensor();
        return;
    }
    // generic path: any libtorch loss over the fused decoder autograd::Function
    batch->node_embeddings_.requires_grad_();
    auto s = forward_lp(batch, true);
    torch::Tensor loss;
    if (std::get<3>(s).defined()) {
        loss = (*loss_function_)(std::get<2>(s), std::get<3>(s), true) + (*loss_function_)(std::get<0>(s), std::get<1>(s), true);  // model.cpp:309-312
    } else {
        loss = (*loss_function_)(std::get<0>(s), std::get<1>(s), true);
    }
    loss.backward();
    if (call_step) step();
    batch->accumulateGradients(sparse_lr_);
}

float Model::train_batch_fused(shared_ptr<Batch> batch, InMemory& embeddings, InMemory& state, bool call_step) {
    // ComputeWorkerGPU::run (pipeline_gpu.cpp:49-91): loadGPUParameters -> train_batch -> updateEmbeddings(batch, true), one C-ABI call
    if (!fusable(*this, *batch)) throw MariusRuntimeException("train_batch_fused needs SoftmaxCrossEntropy + CORRUPT_NODE without score filters");
    if (call_step) clear_grad();
    auto ids = batch->unique_node_indices_.to(embeddings.data_.device()).to(torch::kInt64).contiguous();
    auto& t = embeddings.data_;
    auto& st = state.data_;
    torch::Tensor edges, dn, sn;
    Batch dev_batch = *batch;
    dev_batch.edges_ = batch->edges_.to(t.device());
    dev_batch.dst_neg_indices_mapping_ = batch->dst_neg_indices_mapping_.to(t.device());
    if (batch->src_neg_indices_mapping_.defined()) dev_batch.src_neg_indices_mapping_ = batch->src_neg_indices_mapping_.to(t.device());
    mb_batch mbb = describe(*this, dev_batch, ids.size(0), t.size(1), edges, dn, sn);
    auto loss = torch::empty({1}, t.options());
    torch::Tensor rg, irg;
    if (mbb.rel) rg = torch::empty_like(decoder_->relations_);
    if (mbb.inv_rel) irg = torch::empty_like(decoder_->inverse_relations_);
    auto red = std::dynamic_pointer_cast<SoftmaxCrossEntropy>(loss_function_)->reduction_type_;
    mb_throw_on_error(mb_train_step(mb_context_for(t.device()), &mbb, t.data_ptr<float>(), st.data_ptr<float>(), t.size(0), t.stride(0),
                                    ids.data_ptr<int64_t>(), sparse_lr_, red == LossReduction::SUM ? MB_REDUCTION_SUM : MB_REDUCTION_MEAN,
                                    mb_default_precision(), loss.data_ptr<float>(), mbb.rel ? rg.data_ptr<float>() : nullptr,
                                    mbb.inv_rel ? irg.data_ptr<float>() : nullptr, mb_current_stream(t.device())));
    if (rg.defined()) set_grad(decoder_->relations_, rg);
    if (irg.defined()) set_grad(decoder_->inverse_relations_, irg);
    if (call_step) step();
    return loss.item<float>();
}

// ---- reporting (reporting/reporting.cpp) -----------------------------------------------------------------------------------------
HitskMetric::HitskMetric(int k) : k_(k) {
    name_ = "Hits@" + std::to_string(k);
    unit_ = "";
}
torch::Tensor HitskMetric::computeMetric(torch::Tensor ranks) {  // reporting.cpp:17
    return torch::tensor((double)ranks.le(k_).nonzero().size(0) / ranks.size(0), torch::kFloat64);
}
MeanRankMetric::MeanRankMetric() {
    name_ = "Mean Rank";
    unit_ = "";
}
torch::Tensor MeanRankMetric::computeMetric(torch::Tensor ranks) { return ranks.to(torch::kFloat64).mean(); }  // reporting.cpp:24
MeanReciprocalRankMetric::MeanReciprocalRankMetric() {
    name_ = "MRR";
    unit_ = "";
}
torch::Tensor MeanReciprocalRankMetric::computeMetric(torch::Tensor ranks) { return ranks.to(torch::kFloat32).reciprocal().mean(); }  // reporting.cpp:31

void LinkPredictionReporter::clear() {
    all_ranks_ = torch::Tensor();
    all_scores_ = torch::Tensor();
    per_batch_ranks_ = {};
    per_batch_scores_ = {};
    per_batch_edges_ = {};
}

torch::Tensor LinkPredictionReporter::computeRanks(torch::Tensor pos_scores, torch::Tensor neg_scores) {
    // reporting.cpp:56-58  (neg_scores >= pos_scores.unsqueeze(1)).sum(1) + 1
    if (!pos_scores.defined() || !neg_scores.defined()) throw UndefinedTensorException();
    if (pos_scores.dim() != 1 || neg_scores.dim() != 2 || pos_scores.size(0) != neg_scores.size(0)) throw TensorSizeMismatchException(neg_scores, "computeRanks");
    torch::Tensor pos = pos_scores.to(torch::kFloat32).contiguous();
    torch::Tensor neg = neg_scores.to(torch::kFloat32);
    if (neg.stride(1) != 1) neg = neg.contiguous();
    torch::Tensor ranks = torch::empty({pos.size(0)}, torch::TensorOptions().dtype(torch::kInt64).device(pos.device()));
    mb_throw_on_error(mb_compute_ranks(pos.data_ptr<float>(), neg.data_ptr<float>(), neg.size(0), neg.size(1), neg.size(0) > 1 ? neg.stride(0) : neg.size(1),
                                       ranks.data_ptr<int64_t>(), mb_current_stream(pos.device())));
    return ranks;
}

void LinkPredictionReporter::addResult(torch::Tensor pos_scores, torch::Tensor neg_scores, torch::Tensor edges) {
    std::lock_guard<std::mutex> guard(lock_);
    if (neg_scores.defined()) per_batch_ranks_.emplace_back(computeRanks(pos_scores, neg_scores));
    if (edges.defined()) {
        per_batch_scores_.emplace_back(pos_scores.to(torch::kCPU));
        per_batch_edges_.emplace_back(edges.to(torch::kCPU));
    }
}

void LinkPredictionReporter::addRanks(torch::Tensor ranks) {
    std::lock_guard<std::mutex> guard(lock_);
    per_batch_ranks_.emplace_back(ranks);
}

string LinkPredictionReporter::report() {  // reporting.cpp:74-95 (returns the text instead of logging it through spdlog)
    all_ranks_ = torch::cat(per_batch_ranks_).to(torch::kCPU);
    if (!per_batch_scores_.empty()) all_scores_ = torch::cat(per_batch_scores_);
    per_batch_ranks_ = {};
    per_batch_scores_ = {};
    string out = "\n=================================\nLink Prediction: " + std::to_string(all_ranks_.size(0)) + " edges evaluated\n";
    for (auto& m : metrics_) out += m->name_ + ": " + std::to_string(m->computeMetric(all_ranks_).item<double>()) + m->unit_ + "\n";
    return out + "=================================";
}

void Model::evaluate_batch(shared_ptr<Batch> batch) {  // model.cpp:335-349
    if (reporter_ == nullptr) throw UnexpectedNullPtrException("reporter_");
    if (decoder_->decoder_method_ != EdgeDecoderMethod::CORRUPT_NODE || !batch->dst_neg_indices_mapping_.defined()) {
        // the reference's generic route: forward_lp, then addResult per side that has negative scores
        auto scores = forward_lp(batch, true);
        if (std::get<1>(scores).defined()) reporter_->addResult(std::get<0>(scores), std::get<1>(scores));
        if (std::get<3>(scores).defined()) reporter_->addResult(std::get<2>(scores), std::get<3>(scores));
        return;
    }
    torch::NoGradGuard ng;
    if (!batch->node_embeddings_.defined()) throw UndefinedTensorException();
    torch::Tensor emb = batch->node_embeddings_.detach().contiguous();
    torch::Tensor edges, dn, sn;
    mb_batch mbb = describe(*this, *batch, emb.size(0), emb.size(1), edges, dn, sn);
    const int64_t Bc = (mbb.B + mbb.C - 1) / mbb.C, Bp = Bc * mbb.C;
    const bool inverse = mbb.src_negs != nullptr;
    auto iopt = torch::TensorOptions().dtype(torch::kInt64).device(emb.device());
    torch::Tensor ranks = torch::empty({Bp}, iopt), inv_ranks = inverse ? torch::empty({Bp}, iopt) : torch::Tensor();
    torch::Tensor df = batch->dst_neg_filter_.defined() ? batch->dst_neg_filter_.to(torch::kInt64).contiguous() : torch::Tensor();
    torch::Tensor sf = (inverse && batch->src_neg_filter_.defined()) ? batch->src_neg_filter_.to(torch::kInt64).contiguous() : torch::Tensor();
    mb_throw_on_error(mb_evaluate_batch(mb_context_for(emb.device()), &mbb, emb.data_ptr<float>(), emb.stride(0), mb_default_precision(),
                                        df.defined() ? df.data_ptr<int64_t>() : nullptr, df.defined() ? df.size(0) : 0,
                                        sf.defined() ? sf.data_ptr<int64_t>() : nullptr, sf.defined() ? sf.size(0) : 0, ranks.data_ptr<int64_t>(),
                                        inverse ? inv_ranks.data_ptr<int64_t>() : nullptr, nullptr, nullptr, mb_current_stream(emb.device())));
    reporter_->addRanks(ranks);
    if (inverse) reporter_->addRanks(inv_ranks);
}

// ---- compute stage ---------------------------------------------------------------------------------------------------------------
ComputeWorkerGPU::ComputeWorkerGPU(shared_ptr<Model> model, shared_ptr<InMemory> embeddings, shared_ptr<InMemory> state, size_t queue_size)
    : model_(model), embeddings_(embeddings), state_(state) {
    device_loaded_batches_ = std::make_shared<Queue<shared_ptr<Batch>>>(queue_size);
    device_update_batches_ = std::make_shared<Queue<shared_ptr<Batch>>>(1 << 20);
}

ComputeWorkerGPU::~ComputeWorkerGPU() { stop(); }

void ComputeWorkerGPU::run() {
    while (!done_) {
        auto tup = device_loaded_batches_->blocking_pop();
        if (!std::get<0>(tup)) break;  // queue closed and drained
        shared_ptr<Batch> batch = std::get<1>(tup);
        try {
            // loadGPUParameters + train_batch + updateEmbeddings(gpu) (pipeline_gpu.cpp:49-91) in one fused call on this device's tables
            batch->loss_ = model_->train_batch_fused(batch, *embeddings_, *state_, true);
            edges_processed_ += batch->edges_.size(0);
            batches_processed_ += 1;
        } catch (const std::exception& e) {
            error_ = e.what();
            done_ = true;
        }
        device_update_batches_->blocking_push(batch);
    }
}

void ComputeWorkerGPU::start() {
    if (thread_ == nullptr) thread_ = new std::thread(&ComputeWorkerGPU::run, this);
}

void ComputeWorkerGPU::stop() {
    if (thread_ != nullptr) {
        device_loaded_batches_->close();
        if (thread_->joinable()) thread_->join();
        delete thread_;
        thread_ = nullptr;
    }
}
