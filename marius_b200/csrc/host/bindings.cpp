// bindings.cpp -- pybind11 module `marius_b200.lib._host`, mirroring the reference's python binding surface
// (src/cpp/python_bindings/{storage,nn,data}/*_wrap.cpp: property names without the trailing underscore).
#include "marius_host.h"

namespace py = pybind11;

PYBIND11_MODULE(_host, m) {
    m.doc() = "C++/libtorch adapters of the marius_b200 hot path (Storage / PartitionBuffer / EdgeDecoder / Batch / Model)";
    py::register_exception<MariusRuntimeException>(m, "MariusRuntimeException", PyExc_RuntimeError);

    m.def("set_default_precision", &mb_set_default_precision);
    m.def("default_precision", &mb_default_precision);

    py::enum_<LossReduction>(m, "LossReduction").value("MEAN", LossReduction::MEAN).value("SUM", LossReduction::SUM);
    py::enum_<EdgeDecoderMethod>(m, "EdgeDecoderMethod")
        .value("ONLY_POS", EdgeDecoderMethod::ONLY_POS)
        .value("POS_AND_NEG", EdgeDecoderMethod::POS_AND_NEG)
        .value("CORRUPT_NODE", EdgeDecoderMethod::CORRUPT_NODE)
        .value("CORRUPT_REL", EdgeDecoderMethod::CORRUPT_REL);

    // storage_wrap.cpp:33-52
    py::class_<Storage, shared_ptr<Storage>>(m, "Storage")
        .def_readwrite("dim0_size", &Storage::dim0_size_)
        .def_readwrite("dim1_size", &Storage::dim1_size_)
        .def_readwrite("data", &Storage::data_)
        .def_readwrite("filename", &Storage::filename_)
        .def("indexRead", &Storage::indexRead, py::arg("indices"))
        .def("indexAdd", &Storage::indexAdd, py::arg("indices"), py::arg("values"))
        .def("range", &Storage::range, py::arg("offset"), py::arg("n"))
        .def("indexPut", &Storage::indexPut, py::arg("indices"), py::arg("values"))
        .def("rangePut", &Storage::rangePut, py::arg("offset"), py::arg("n"), py::arg("values"))
        .def("load", &Storage::load)
        .def("write", &Storage::write)
        .def("unload", &Storage::unload, py::arg("write") = false)
        .def("shuffle", &Storage::shuffle)
        .def("sort", &Storage::sort, py::arg("src"))
        .def("getDim0", &Storage::getDim0);

    py::class_<InMemory, Storage, shared_ptr<InMemory>>(m, "InMemory")
        .def(py::init<torch::Tensor>(), py::arg("data"))
        .def(py::init([](string filename, int64_t dim0, int64_t dim1, torch::Device device) {
                 return std::make_shared<InMemory>(filename, dim0, dim1, torch::kFloat32, device);
             }),
             py::arg("filename"), py::arg("dim0_size"), py::arg("dim1_size"), py::arg("device"))
        .def(py::init([](string filename, torch::Tensor data, torch::Device device) { return std::make_shared<InMemory>(filename, data, device); }),
             py::arg("filename"), py::arg("data"), py::arg("device"))
        .def_static("adagradUpdate", &InMemory::adagradUpdate, py::arg("embeddings"), py::arg("state"), py::arg("indices"), py::arg("gradients"),
                    py::arg("learning_rate"));

    py::class_<PartitionBuffer, shared_ptr<PartitionBuffer>>(m, "PartitionBuffer")
        .def(py::init([](int capacity, int num_partitions, int fine_to_coarse_ratio, int64_t partition_size, int embedding_size, int64_t total_embeddings,
                         string filename, bool prefetching, torch::Device device) {
                 return std::make_shared<PartitionBuffer>(capacity, num_partitions, fine_to_coarse_ratio, partition_size, embedding_size, total_embeddings,
                                                          torch::kFloat32, filename, prefetching, device);
             }),
             py::arg("capacity"), py::arg("num_partitions"), py::arg("fine_to_coarse_ratio"), py::arg("partition_size"), py::arg("embedding_size"),
             py::arg("total_embeddings"), py::arg("filename"), py::arg("prefetching") = false, py::arg("device") = torch::Device(torch::kCUDA, 0))
        .def("load", &PartitionBuffer::load)
        .def("write", &PartitionBuffer::write)
        .def("unload", &PartitionBuffer::unload, py::arg("write"))
        .def("getNextAdmit", &PartitionBuffer::getNextAdmit)
        .def("getNextEvict", &PartitionBuffer::getNextEvict)
        .def("getRandomIds", &PartitionBuffer::getRandomIds, py::arg("size"))
        .def("indexRead", &PartitionBuffer::indexRead, py::arg("indices"))
        .def("indexAdd", &PartitionBuffer::indexAdd, py::arg("indices"), py::arg("values"))
        .def("adagradUpdate", &PartitionBuffer::adagradUpdate, py::arg("state"), py::arg("indices"), py::arg("gradients"), py::arg("learning_rate"))
        .def("getGlobalToLocalMap", &PartitionBuffer::getGlobalToLocalMap, py::arg("get_current"))
        .def("setBufferOrdering", &PartitionBuffer::setBufferOrdering, py::arg("buffer_states"))
        .def("hasSwap", &PartitionBuffer::hasSwap)
        .def("performNextSwap", &PartitionBuffer::performNextSwap)
        .def("sync", &PartitionBuffer::sync)
        .def("getNumInMemory", &PartitionBuffer::getNumInMemory)
        .def("getBufferState", &PartitionBuffer::getBufferState)
        .def("bufferTensor", &PartitionBuffer::bufferTensor);

    // storage_wrap.cpp: the Storage facade over the buffer (storage.h:90-145)
    py::class_<PartitionBufferStorage, Storage, shared_ptr<PartitionBufferStorage>>(m, "PartitionBufferStorage")
        .def(py::init([](string filename, int64_t dim0, int64_t dim1, int num_partitions, int buffer_capacity, bool prefetching, int fine_to_coarse_ratio,
                         torch::Device device) {
                 auto o = std::make_shared<PartitionBufferOptions>();
                 o->num_partitions = num_partitions;
                 o->buffer_capacity = buffer_capacity;
                 o->prefetching = prefetching;
                 o->fine_to_coarse_ratio = fine_to_coarse_ratio;
                 return std::make_shared<PartitionBufferStorage>(filename, dim0, dim1, o, device);
             }),
             py::arg("filename"), py::arg("dim0_size"), py::arg("dim1_size"), py::arg("num_partitions"), py::arg("buffer_capacity"),
             py::arg("prefetching") = true, py::arg("fine_to_coarse_ratio") = 1, py::arg("device") = torch::Device(torch::kCUDA, 0))
        .def(py::init([](string filename, torch::Tensor data, int num_partitions, int buffer_capacity, bool prefetching, int fine_to_coarse_ratio,
                         torch::Device device) {
                 auto o = std::make_shared<PartitionBufferOptions>();
                 o->num_partitions = num_partitions;
                 o->buffer_capacity = buffer_capacity;
                 o->prefetching = prefetching;
                 o->fine_to_coarse_ratio = fine_to_coarse_ratio;
                 return std::make_shared<PartitionBufferStorage>(filename, data, o, device);
             }),
             py::arg("filename"), py::arg("data"), py::arg("num_partitions"), py::arg("buffer_capacity"), py::arg("prefetching") = true,
             py::arg("fine_to_coarse_ratio") = 1, py::arg("device") = torch::Device(torch::kCUDA, 0))
        .def("append", &PartitionBufferStorage::append, py::arg("values"))
        .def("hasSwap", &PartitionBufferStorage::hasSwap)
        .def("performNextSwap", &PartitionBufferStorage::performNextSwap)
        .def("getGlobalToLocalMap", &PartitionBufferStorage::getGlobalToLocalMap, py::arg("get_current"))
        .def("sync", &PartitionBufferStorage::sync)
        .def("setBufferOrdering", &PartitionBufferStorage::setBufferOrdering, py::arg("buffer_states"))
        .def("getNextAdmit", &PartitionBufferStorage::getNextAdmit)
        .def("getNextEvict", &PartitionBufferStorage::getNextEvict)
        .def("getNumInMemory", &PartitionBufferStorage::getNumInMemory)
        .def("getRandomIds", &PartitionBufferStorage::getRandomIds, py::arg("size"))
        .def("adagradUpdate",
             [](PartitionBufferStorage& self, PartitionBufferStorage& state, torch::Tensor indices, torch::Tensor gradients, float lr) {
                 self.buffer_->adagradUpdate(*state.buffer_, indices, gradients, lr);
             },
             py::arg("state"), py::arg("indices"), py::arg("gradients"), py::arg("learning_rate"));

    // edge_decoder_wrap.cpp:8-22
    py::class_<EdgeDecoder, shared_ptr<EdgeDecoder>>(m, "EdgeDecoder")
        .def_readwrite("relations", &EdgeDecoder::relations_)
        .def_readwrite("inverse_relations", &EdgeDecoder::inverse_relations_)
        .def_readwrite("num_relations", &EdgeDecoder::num_relations_)
        .def_readwrite("embedding_size", &EdgeDecoder::embedding_size_)
        .def_readwrite("decoder_method", &EdgeDecoder::decoder_method_)
        .def_readwrite("use_inverse_relations", &EdgeDecoder::use_inverse_relations_)
        .def("apply_relation", &EdgeDecoder::apply_relation, py::arg("nodes"), py::arg("relations"))
        .def("compute_scores", &EdgeDecoder::compute_scores, py::arg("src"), py::arg("dst"))
        .def("select_relations", &EdgeDecoder::select_relations, py::arg("indices"), py::arg("inverse") = false);

    auto mode_to_method = [](const string& mode) {
        if (mode == "infer") return EdgeDecoderMethod::ONLY_POS;   // distmult_wrap.cpp: "infer" / "train"
        if (mode == "train") return EdgeDecoderMethod::CORRUPT_NODE;
        throw std::runtime_error("Unsupported decoder mode");
    };
    py::class_<DistMult, EdgeDecoder, shared_ptr<DistMult>>(m, "DistMult")
        .def(py::init([mode_to_method](int num_relations, int embedding_dim, bool use_inverse_relations, torch::Device device, string mode) {
                 return std::make_shared<DistMult>(num_relations, embedding_dim, torch::TensorOptions().dtype(torch::kFloat32).device(device),
                                                   use_inverse_relations, mode_to_method(mode));
             }),
             py::arg("num_relations"), py::arg("embedding_dim"), py::arg("use_inverse_relations") = true, py::arg("device") = torch::Device(torch::kCUDA, 0),
             py::arg("mode") = "train");
    py::class_<ComplEx, EdgeDecoder, shared_ptr<ComplEx>>(m, "ComplEx")
        .def(py::init([mode_to_method](int num_relations, int embedding_dim, bool use_inverse_relations, torch::Device device, string mode) {
                 return std::make_shared<ComplEx>(num_relations, embedding_dim, torch::TensorOptions().dtype(torch::kFloat32).device(device),
                                                  use_inverse_relations, mode_to_method(mode));
             }),
             py::arg("num_relations"), py::arg("embedding_dim"), py::arg("use_inverse_relations") = true, py::arg("device") = torch::Device(torch::kCUDA, 0),
             py::arg("mode") = "train");

    m.def("node_corrupt_forward", &node_corrupt_forward, py::arg("decoder"), py::arg("positive_edges"), py::arg("node_embeddings"), py::arg("dst_negs"),
          py::arg("src_negs"));
    m.def("only_pos_forward", &only_pos_forward, py::arg("decoder"), py::arg("edges"), py::arg("node_embeddings"));

    py::class_<LossFunction, shared_ptr<LossFunction>>(m, "LossFunction")
        .def("__call__", [](LossFunction& l, torch::Tensor a, torch::Tensor b, bool scores) { return l(a, b, scores); }, py::arg("y_pred"),
             py::arg("targets"), py::arg("scores") = true);
    py::class_<SoftmaxCrossEntropy, LossFunction, shared_ptr<SoftmaxCrossEntropy>>(m, "SoftmaxCrossEntropy")
        .def(py::init([](string reduction) {
                 return std::make_shared<SoftmaxCrossEntropy>(reduction == "mean" ? LossReduction::MEAN : LossReduction::SUM);
             }),
             py::arg("reduction") = "sum");

    // batch_wrap.cpp
    py::class_<Batch, shared_ptr<Batch>>(m, "Batch")
        .def(py::init<bool>(), py::arg("train"))
        .def_readwrite("batch_id", &Batch::batch_id_)
        .def_readwrite("train", &Batch::train_)
        .def_readwrite("device_id", &Batch::device_id_)
        .def_readwrite("unique_node_indices", &Batch::unique_node_indices_)
        .def_readwrite("node_embeddings", &Batch::node_embeddings_)
        .def_readwrite("node_gradients", &Batch::node_gradients_)
        .def_readwrite("node_embeddings_state", &Batch::node_embeddings_state_)
        .def_readwrite("node_state_update", &Batch::node_state_update_)
        .def_readwrite("src_neg_indices_mapping", &Batch::src_neg_indices_mapping_)
        .def_readwrite("dst_neg_indices_mapping", &Batch::dst_neg_indices_mapping_)
        .def_readwrite("edges", &Batch::edges_)
        .def_readwrite("src_neg_filter", &Batch::src_neg_filter_)
        .def_readwrite("dst_neg_filter", &Batch::dst_neg_filter_)
        .def_readwrite("loss", &Batch::loss_)
        .def("to", &Batch::to, py::arg("device"))
        .def("accumulateGradients", &Batch::accumulateGradients, py::arg("learning_rate"))
        .def("embeddingsToHost", &Batch::embeddingsToHost)
        .def("clear", &Batch::clear);

    // the compute stage of the GPU pipeline (pipeline/pipeline_gpu.cpp:33-104) for device-resident tables
    py::class_<ComputeWorkerGPU, shared_ptr<ComputeWorkerGPU>>(m, "ComputeWorkerGPU")
        .def(py::init<shared_ptr<Model>, shared_ptr<InMemory>, shared_ptr<InMemory>, size_t>(), py::arg("model"), py::arg("embeddings"), py::arg("state"),
             py::arg("queue_size") = 4)
        .def("start", &ComputeWorkerGPU::start)
        .def("stop", &ComputeWorkerGPU::stop, py::call_guard<py::gil_scoped_release>())
        .def("push", [](ComputeWorkerGPU& w, shared_ptr<Batch> b) { w.device_loaded_batches_->blocking_push(b); }, py::arg("batch"),
             py::call_guard<py::gil_scoped_release>())
        .def("pop_finished",
             [](ComputeWorkerGPU& w) {
                 auto t = w.device_update_batches_->blocking_pop();
                 return std::get<1>(t);
             },
             py::call_guard<py::gil_scoped_release>())
        .def_property_readonly("edges_processed", [](ComputeWorkerGPU& w) { return (int64_t)w.edges_processed_; })
        .def_property_readonly("batches_processed", [](ComputeWorkerGPU& w) { return (int64_t)w.batches_processed_; })
        .def_property_readonly("error", [](ComputeWorkerGPU& w) { return w.error(); });

    // reporting_wrap.cpp
    py::class_<RankingMetric, shared_ptr<RankingMetric>>(m, "RankingMetric")
        .def_readwrite("name", &RankingMetric::name_)
        .def("compute_metric", &RankingMetric::computeMetric, py::arg("ranks"));
    py::class_<HitskMetric, RankingMetric, shared_ptr<HitskMetric>>(m, "Hitsk").def(py::init<int>(), py::arg("k"));
    py::class_<MeanRankMetric, RankingMetric, shared_ptr<MeanRankMetric>>(m, "MeanRank").def(py::init<>());
    py::class_<MeanReciprocalRankMetric, RankingMetric, shared_ptr<MeanReciprocalRankMetric>>(m, "MeanReciprocalRank").def(py::init<>());
    py::class_<LinkPredictionReporter, shared_ptr<LinkPredictionReporter>>(m, "LinkPredictionReporter")
        .def(py::init<>())
        .def_readwrite("per_batch_ranks", &LinkPredictionReporter::per_batch_ranks_)
        .def_readwrite("all_ranks", &LinkPredictionReporter::all_ranks_)
        .def("add_metric", &LinkPredictionReporter::addMetric, py::arg("metric"))
        .def("clear", &LinkPredictionReporter::clear)
        .def("compute_ranks", &LinkPredictionReporter::computeRanks, py::arg("pos_scores"), py::arg("neg_scores"))
        .def("add_result", &LinkPredictionReporter::addResult, py::arg("pos_scores"), py::arg("neg_scores"), py::arg("edges") = torch::Tensor())
        .def("report", &LinkPredictionReporter::report);

    // model_wrap.cpp
    py::class_<Model, shared_ptr<Model>>(m, "Model")
        .def(py::init([](shared_ptr<EdgeDecoder> decoder, shared_ptr<LossFunction> loss, torch::Device device) {
                 return std::make_shared<Model>(decoder, loss, device);
             }),
             py::arg("decoder"), py::arg("loss"), py::arg("device") = torch::Device(torch::kCUDA, 0))
        .def_readwrite("decoder", &Model::decoder_)
        .def_readwrite("loss_function", &Model::loss_function_)
        .def_readwrite("sparse_lr", &Model::sparse_lr_)
        .def_readwrite("dense_lr", &Model::dense_lr_)
        .def("forward_lp", &Model::forward_lp, py::arg("batch"), py::arg("train"))
        .def("train_batch", &Model::train_batch, py::arg("batch"), py::arg("call_step") = true)
        .def("train_batch_fused", &Model::train_batch_fused, py::arg("batch"), py::arg("embeddings"), py::arg("state"), py::arg("call_step") = true)
        .def_readwrite("reporter", &Model::reporter_)
        .def("evaluate_batch", &Model::evaluate_batch, py::arg("batch"))
        .def("clear_grad", &Model::clear_grad)
        .def("step", &Model::step)
        .def("parameters", [](Model& mdl) { return mdl.parameters(); });
}
