// decoder.cpp -- EdgeDecoder / DistMult / ComplEx and the decoder methods (nn/decoders/edge/*) on the fused kernels.
// node_corrupt_forward is a torch::autograd::Function: forward = mb_decoder_forward, backward = mb_decoder_backward, so
// `loss.backward()` (model.cpp:324) works with ANY of the reference's libtorch losses, while Model::train_batch takes the
// fully fused SoftmaxCE path when it can.
#include "marius_host.h"

namespace {

mb_batch make_batch_desc(const shared_ptr<EdgeDecoder>& dec, const torch::Tensor& edges, const torch::Tensor& emb, const torch::Tensor& dst_negs,
                         const torch::Tensor& src_negs) {
    mb_batch b;
    b.decoder = dec->decoder_kind_;
    b.U = emb.size(0);
    b.d = emb.size(1);
    b.B = edges.size(0);
    b.edge_cols = (int)edges.size(1);
    b.edges = edges.data_ptr<int64_t>();
    b.C = (int)dst_negs.size(0);
    b.N = (int)dst_negs.size(1);
    b.dst_negs = dst_negs.data_ptr<int64_t>();
    bool has_rel = edges.size(1) == 3 && dec->relations_.defined();
    bool inverse = has_rel && dec->use_inverse_relations_ && src_negs.defined();
    b.src_negs = inverse ? src_negs.data_ptr<int64_t>() : nullptr;
    b.R = has_rel ? dec->relations_.size(0) : 0;
    b.rel = has_rel ? dec->relations_.data_ptr<float>() : nullptr;
    b.inv_rel = inverse ? dec->inverse_relations_.data_ptr<float>() : nullptr;
    return b;
}

void check_edges(const torch::Tensor& edges) {
    if (!edges.defined()) throw UndefinedTensorException();
    if (edges.dim() != 2 || (edges.size(1) != 3 && edges.size(1) != 2))
        throw TensorSizeMismatchException(edges, "Edge list must be a 3 or 2 column tensor");  // decoder_methods.cpp:66-72
}

struct FusedNodeCorrupt : public torch::autograd::Function<FusedNodeCorrupt> {
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor node_embeddings, torch::Tensor relations,
                                                  torch::Tensor inverse_relations, torch::Tensor edges, torch::Tensor dst_negs, torch::Tensor src_negs,
                                                  int64_t decoder_kind, bool use_inverse, int64_t precision) {
        auto undef = [](const torch::Tensor& t) { return (t.defined() && t.numel() > 0) ? t : torch::Tensor(); };
        relations = undef(relations);
        inverse_relations = undef(inverse_relations);
        src_negs = undef(src_negs);
        auto dec = std::make_shared<EdgeDecoder>();
        dec->decoder_kind_ = (int)decoder_kind;
        dec->relations_ = relations;
        dec->inverse_relations_ = inverse_relations;
        dec->use_inverse_relations_ = use_inverse;
        auto emb = node_embeddings.contiguous();
        mb_batch b = make_batch_desc(dec, edges, emb, dst_negs, src_negs);
        const int64_t Bc = (b.B + b.C - 1) / b.C, Bp = Bc * b.C;
        const bool inverse = b.src_negs != nullptr;
        auto opts = emb.options();
        auto pos = torch::empty({Bp}, opts), neg = torch::empty({Bp, (int64_t)b.N}, opts);
        torch::Tensor inv_pos, inv_neg;
        if (inverse) {
            inv_pos = torch::empty({Bp}, opts);
            inv_neg = torch::empty({Bp, (int64_t)b.N}, opts);
        }
        mb_throw_on_error(mb_decoder_forward(mb_context_for(emb.device()), &b, emb.data_ptr<float>(), emb.stride(0), (int)precision, pos.data_ptr<float>(),
                                             neg.data_ptr<float>(), inverse ? inv_pos.data_ptr<float>() : nullptr,
                                             inverse ? inv_neg.data_ptr<float>() : nullptr, mb_current_stream(emb.device())));
        auto ph = [&](const torch::Tensor& t, torch::Dtype dt) { return t.defined() ? t : torch::empty({0}, emb.options().dtype(dt)); };
        ctx->save_for_backward({emb, ph(relations, torch::kFloat32), ph(inverse_relations, torch::kFloat32), edges, dst_negs, ph(src_negs, torch::kInt64)});
        ctx->saved_data["kind"] = decoder_kind;
        ctx->saved_data["inverse"] = use_inverse;
        ctx->saved_data["precision"] = precision;
        if (!inverse) {  // undefined outputs are not allowed in a variable_list: return empty placeholders
            inv_pos = torch::empty({0}, opts);
            inv_neg = torch::empty({0}, opts);
        }
        return {pos, neg, inv_pos, inv_neg};
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grads) {
        auto saved = ctx->get_saved_variables();
        auto undef = [](const torch::Tensor& t) { return (t.defined() && t.numel() > 0) ? t : torch::Tensor(); };
        auto emb = saved[0], relations = undef(saved[1]), inverse_relations = undef(saved[2]), edges = saved[3], dst_negs = saved[4],
             src_negs = undef(saved[5]);
        auto dec = std::make_shared<EdgeDecoder>();
        dec->decoder_kind_ = (int)ctx->saved_data["kind"].toInt();
        dec->relations_ = relations;
        dec->inverse_relations_ = inverse_relations;
        dec->use_inverse_relations_ = ctx->saved_data["inverse"].toBool();
        mb_batch b = make_batch_desc(dec, edges, emb, dst_negs, src_negs);
        const int64_t Bc = (b.B + b.C - 1) / b.C, Bp = Bc * b.C;
        const bool inverse = b.src_negs != nullptr;
        auto opts = emb.options();
        auto dense = [&](const torch::Tensor& g, std::vector<int64_t> shape) {
            return g.defined() ? g.to(torch::kFloat32).contiguous() : torch::zeros(shape, opts);
        };
        auto gpos = dense(grads[0], {Bp}), gneg = dense(grads[1], {Bp, (int64_t)b.N});
        torch::Tensor gipos, gineg;
        if (inverse) {
            gipos = dense(grads[2], {Bp});
            gineg = dense(grads[3], {Bp, (int64_t)b.N});
        }
        auto grad_emb = torch::empty_like(emb);
        torch::Tensor grad_rel, grad_inv;
        if (b.rel) grad_rel = torch::empty_like(relations);
        if (inverse) grad_inv = torch::empty_like(inverse_relations);
        else if (inverse_relations.defined()) grad_inv = torch::zeros_like(inverse_relations);
        mb_throw_on_error(mb_decoder_backward(mb_context_for(emb.device()), &b, emb.data_ptr<float>(), emb.stride(0), (int)ctx->saved_data["precision"].toInt(),
                                              gpos.data_ptr<float>(), gneg.data_ptr<float>(), inverse ? gipos.data_ptr<float>() : nullptr,
                                              inverse ? gineg.data_ptr<float>() : nullptr, grad_emb.data_ptr<float>(),
                                              b.rel ? grad_rel.data_ptr<float>() : nullptr, inverse ? grad_inv.data_ptr<float>() : nullptr,
                                              mb_current_stream(emb.device())));
        return {grad_emb, grad_rel, grad_inv, torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor()};
    }
};

}  // namespace

// ---- EdgeDecoder -----------------------------------------------------------------------------------------------
torch::Tensor EdgeDecoder::select_relations(torch::Tensor indices, bool inverse) {
    if (inverse) {
        if (!inverse_relations_.defined()) throw UndefinedTensorException();  // edge_decoder.cpp:12-15
        return inverse_relations_.index_select(0, indices);
    }
    return relations_.index_select(0, indices);
}

torch::Tensor EdgeDecoder::apply_relation(torch::Tensor nodes, torch::Tensor relations) {
    // Stand-alone operator (relation_operators.cpp) for callers that compose the decoder by hand: a [B,d] elementwise op,
    // kept in libtorch.  The training / scoring paths never call it: the operator is fused into the edge_prep kernel.
    if (!relations.defined() || decoder_kind_ == MB_DECODER_DOT) return nodes;
    if (decoder_kind_ == MB_DECODER_DISTMULT) return nodes * relations;
    int64_t h = nodes.size(1) / 2;
    auto er = nodes.narrow(1, 0, h), ei = nodes.narrow(1, h, nodes.size(1) - h);
    auto rr = relations.narrow(1, 0, h), ri = relations.narrow(1, h, nodes.size(1) - h);
    return torch::cat({er * rr - ei * ri, er * ri + ei * rr}, 1);
}

torch::Tensor EdgeDecoder::compute_scores(torch::Tensor src, torch::Tensor dst) {
    if (!src.defined() || !dst.defined()) throw UndefinedTensorException();  // comparators.cpp:63-65
    if (src.sizes() == dst.sizes()) return (src * dst).sum(-1);
    // chunked negatives [C,N,d]: route through the contraction kernels via the fused function with identity relations
    auto U0 = src.size(0);
    auto negs_flat = dst.reshape({-1, dst.size(2)});
    auto emb = torch::cat({src, negs_flat}, 0);
    auto edges = torch::stack({torch::arange(U0, src.options().dtype(torch::kInt64)), torch::arange(U0, src.options().dtype(torch::kInt64))}, 1);
    auto neg_ids = (torch::arange(negs_flat.size(0), src.options().dtype(torch::kInt64)) + U0).reshape({dst.size(0), dst.size(1)});
    auto none_f = torch::empty({0}, src.options());
    auto none_i = torch::empty({0}, src.options().dtype(torch::kInt64));
    auto out = FusedNodeCorrupt::apply(emb, none_f, none_f, edges.contiguous(), neg_ids.contiguous(), none_i, (int64_t)MB_DECODER_DOT, false,
                                       (int64_t)mb_default_precision());
    return out[1];
}

DistMult::DistMult(int num_relations, int embedding_dim, torch::TensorOptions tensor_options, bool use_inverse_relations, EdgeDecoderMethod decoder_method) {
    decoder_kind_ = MB_DECODER_DISTMULT;
    num_relations_ = num_relations;
    embedding_size_ = embedding_dim;
    use_inverse_relations_ = use_inverse_relations;
    tensor_options_ = tensor_options;
    decoder_method_ = decoder_method;
    reset();
}

void DistMult::reset() {
    // distmult.cpp:21-28: ones
    relations_ = register_parameter("relation_embeddings", torch::ones({num_relations_, embedding_size_}, tensor_options_));
    if (use_inverse_relations_) inverse_relations_ = register_parameter("inverse_relation_embeddings", torch::ones({num_relations_, embedding_size_}, tensor_options_));
}

ComplEx::ComplEx(int num_relations, int embedding_dim, torch::TensorOptions tensor_options, bool use_inverse_relations, EdgeDecoderMethod decoder_method) {
    decoder_kind_ = MB_DECODER_COMPLEX;
    num_relations_ = num_relations;
    embedding_size_ = embedding_dim;
    use_inverse_relations_ = use_inverse_relations;
    tensor_options_ = tensor_options;
    decoder_method_ = decoder_method;
    reset();
}

void ComplEx::reset() {
    // complex.cpp:21-30: real part 1, imaginary part 0
    auto init = torch::zeros({num_relations_, embedding_size_}, tensor_options_);
    init.narrow(1, 0, embedding_size_ / 2).fill_(1);
    relations_ = register_parameter("relation_embeddings", init.clone());
    if (use_inverse_relations_) inverse_relations_ = register_parameter("inverse_relation_embeddings", init.clone());
}

// ---- decoder methods -------------------------------------------------------------------------------------------
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> node_corrupt_forward(shared_ptr<EdgeDecoder> decoder, torch::Tensor positive_edges,
                                                                                            torch::Tensor node_embeddings, torch::Tensor dst_negs,
                                                                                            torch::Tensor src_negs) {
    check_edges(positive_edges);
    if (!node_embeddings.defined() || !dst_negs.defined()) throw UndefinedTensorException();
    auto edges = positive_edges.to(torch::kInt64).contiguous();
    auto dn = dst_negs.to(torch::kInt64).contiguous();
    auto sn = src_negs.defined() ? src_negs.to(torch::kInt64).contiguous() : torch::Tensor();
    bool has_rel = edges.size(1) == 3;
    auto none_f = torch::empty({0}, node_embeddings.options());
    auto none_i = torch::empty({0}, dn.options());
    auto out = FusedNodeCorrupt::apply(node_embeddings, has_rel ? decoder->relations_ : none_f,
                                       (has_rel && decoder->use_inverse_relations_) ? decoder->inverse_relations_ : none_f, edges, dn,
                                       sn.defined() ? sn : none_i, (int64_t)(has_rel ? decoder->decoder_kind_ : MB_DECODER_DOT),
                                       decoder->use_inverse_relations_, (int64_t)mb_default_precision());
    torch::Tensor inv_pos = out[2].numel() > 0 || (has_rel && decoder->use_inverse_relations_ && sn.defined()) ? out[2] : torch::Tensor();
    torch::Tensor inv_neg = out[3].numel() > 0 || (has_rel && decoder->use_inverse_relations_ && sn.defined()) ? out[3] : torch::Tensor();
    return std::forward_as_tuple(out[0], out[1], inv_pos, inv_neg);
}

std::tuple<torch::Tensor, torch::Tensor> only_pos_forward(shared_ptr<EdgeDecoder> decoder, torch::Tensor edges, torch::Tensor node_embeddings) {
    // decoder_methods.cpp:7-42: positives only == node_corrupt_forward against a single dummy negative, keeping (pos, inv_pos)
    check_edges(edges);
    auto dummy = torch::zeros({1, 1}, torch::TensorOptions().dtype(torch::kInt64).device(node_embeddings.device()));
    auto r = node_corrupt_forward(decoder, edges, node_embeddings, dummy, decoder->use_inverse_relations_ ? dummy : torch::Tensor());
    int64_t B = edges.size(0);
    auto pos = std::get<0>(r).narrow(0, 0, B);
    torch::Tensor inv_pos = std::get<2>(r).defined() ? std::get<2>(r).narrow(0, 0, B) : torch::Tensor();
    return std::forward_as_tuple(pos, inv_pos);
}

// ---- loss ------------------------------------------------------------------------------------------------------
torch::Tensor SoftmaxCrossEntropy::operator()(torch::Tensor pos, torch::Tensor neg, bool scores) {
    if (!scores) throw MariusRuntimeException("Input to SoftmaxCrossEntropy loss function must be scores.");  // loss.cpp:51-54
    if (!pos.defined() || !neg.defined()) throw UndefinedTensorException();
    if (pos.dim() != 1) throw TensorSizeMismatchException(pos, "Positive scores should be 1-dimensional");
    if (neg.dim() != 2) throw TensorSizeMismatchException(neg, "Negative scores should be 2-dimensional");
    if (pos.size(0) != neg.size(0)) throw TensorSizeMismatchException(pos, "First dimension of pos_scores and neg_scores should match.");
    auto y = torch::cat({pos.unsqueeze(1), neg.logsumexp(1, true)}, -1);
    auto li = y.logsumexp(1) - pos;
    return reduction_type_ == LossReduction::MEAN ? li.mean() : li.sum();
}
