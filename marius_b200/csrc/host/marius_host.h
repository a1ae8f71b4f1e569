// marius_host.h -- C++/libtorch adapters that keep Marius's operator surface for the embedding hot path
//  on top of the C ABI (include/marius_b200.h).  libtorch is plumbing here (tensor handles, autograd
// bookkeeping, streams); every row of data is moved / computed by the sm_100a kernels behind mb_*.
//
//   reference type (src/cpp/include/...)            this file
//   storage/storage.h   Storage, InMemory            Storage, InMemory  (device-resident table)
//   storage/buffer.h    Partition, PartitionedFile,  Partition, PartitionedFile, PartitionBuffer (HBM slab, host file backing),
//                       PartitionBuffer              PartitionBufferStorage
//   nn/decoders/edge/*  EdgeDecoder, DistMult,       EdgeDecoder, DistMult, ComplEx, DotDecoder + only_pos_forward /
//                       ComplEx, decoder methods     node_corrupt_forward (fused autograd::Function)
//   nn/loss.h           LossFunction, SoftmaxCE      LossFunction, SoftmaxCrossEntropy (+ generic libtorch losses via autograd)
//   data/batch.h        Batch                        Batch
//   nn/model.h          Model                        Model (forward_lp / train_batch / evaluate-side scores)
// Error conventions follow common/exception.h:12-42 (all derive std::runtime_error).
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include <torch/extension.h>
#include <ATen/cuda/CUDAEvent.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "marius_b200.h"

using std::shared_ptr;
using std::string;
using std::vector;
typedef torch::Tensor Indices;  // common/datatypes.h

// ---- exceptions (common/exception.h:12-42) -------------------------------------------------------------------
struct MariusRuntimeException : public std::runtime_error {
    explicit MariusRuntimeException(const string& msg) : std::runtime_error(msg) {}
};
struct UndefinedTensorException : public MariusRuntimeException {
    UndefinedTensorException() : MariusRuntimeException("Tensor undefined") {}
};
struct TensorSizeMismatchException : public MariusRuntimeException {
    TensorSizeMismatchException(const torch::Tensor& t, const string& msg) : MariusRuntimeException(msg) { (void)t; }
};
struct UnexpectedNullPtrException : public MariusRuntimeException {
    explicit UnexpectedNullPtrException(const string& msg = "") : MariusRuntimeException(msg) {}
};

// ---- options (configuration/options.h) ------------------------------------------------------------------------
enum class LossReduction { MEAN, SUM };                                          // options.h:24
enum class EdgeDecoderMethod { ONLY_POS, POS_AND_NEG, CORRUPT_NODE, CORRUPT_REL };  // options.h:64
enum class LearningTask { NODE_CLASSIFICATION, LINK_PREDICTION, ENCODE };         // options.h:12

// ---- glue ------------------------------------------------------------------------------------------------------
void mb_throw_on_error(int status);          // MB_ERR_INVALID -> std::runtime_error (storage.cpp:607-610); others -> MariusRuntimeException
void* mb_current_stream(const torch::Device& device);
mb_context* mb_context_for(const torch::Device& device);  // one context per (thread, device): re-entrant like the reference's workers
void mb_set_default_precision(int precision);             // MB_PREC_BF16X3 by default
int mb_default_precision();

// ---- storage ---------------------------------------------------------------------------------------------------
/** Abstract storage class (storage/storage.h:35-86) */
class Storage {
   public:
    int64_t dim0_size_ = 0;
    int64_t dim1_size_ = 0;
    torch::Dtype dtype_ = torch::kFloat32;
    bool initialized_ = false;
    vector<int64_t> edge_bucket_sizes_;
    torch::Tensor data_;
    torch::Device device_ = torch::kCPU;
    string filename_;

    virtual ~Storage() {}
    virtual torch::Tensor indexRead(Indices indices) = 0;
    virtual void indexAdd(Indices indices, torch::Tensor values) = 0;
    virtual torch::Tensor range(int64_t offset, int64_t n) = 0;
    virtual void indexPut(Indices indices, torch::Tensor values) = 0;
    virtual void rangePut(int64_t offset, int64_t n, torch::Tensor values) = 0;
    virtual void load() = 0;
    virtual void write() = 0;
    virtual void unload(bool write = false) = 0;
    virtual void shuffle() = 0;
    virtual void sort(bool src) = 0;
    int64_t getDim0() { return dim0_size_; }
    bool isInitialized() { return initialized_; }
    void setInitialized(bool init) { initialized_ = init; }
};

/** Device-resident table: the DEVICE_MEMORY backend of the reference (InMemory with device_ == cuda, storage.cpp:488-775) */
class InMemory : public Storage {
    bool loaded_ = false;

   public:
    InMemory(string filename, int64_t dim0_size, int64_t dim1_size, torch::Dtype dtype, torch::Device device);
    InMemory(string filename, torch::Tensor data, torch::Device device);
    explicit InMemory(torch::Tensor data);
    void load() override;
    void write() override;
    void unload(bool perform_write) override;
    torch::Tensor indexRead(Indices indices) override;
    void indexAdd(Indices indices, torch::Tensor values) override;
    torch::Tensor range(int64_t offset, int64_t n) override;
    void indexPut(Indices indices, torch::Tensor values) override;
    void rangePut(int64_t offset, int64_t n, torch::Tensor values) override;
    void shuffle() override;
    void sort(bool src) override;
    /** fused accumulateGradients + indexAdd(embeddings) + indexAdd(state) on two device tables (dataloader.cpp:550-557) */
    static void adagradUpdate(InMemory& embeddings, InMemory& state, Indices indices, torch::Tensor gradients, float learning_rate);
};

/** storage/buffer.h:16-41 */
class Partition {
   public:
    std::mutex* lock_;
    std::condition_variable* cv_;
    int partition_id_;
    bool present_ = false;
    bool evicting_ = false;  // an asynchronous write-back of this partition is in flight (buffer.cpp:276-281): readers of the file wait
    int64_t partition_size_;
    int embedding_size_;
    int64_t total_size_;
    int64_t idx_offset_;
    int64_t file_offset_;
    int buffer_idx_ = -1;
    Partition(int partition_id, int64_t partition_size, int embedding_size, int64_t idx_offset, int64_t file_offset)
        : lock_(new std::mutex()), cv_(new std::condition_variable()), partition_id_(partition_id), partition_size_(partition_size),
          embedding_size_(embedding_size), total_size_(partition_size * embedding_size * 4), idx_offset_(idx_offset), file_offset_(file_offset) {}
    ~Partition() {
        delete lock_;
        delete cv_;
    }
};

/** Flat fp32 row-major file of all partitions back to back (storage/buffer.cpp:65-116; the embeddings.bin format, constants.h:39-42) */
class PartitionedFile {
   public:
    string filename_;
    int fd_ = -1;
    explicit PartitionedFile(string filename);
    ~PartitionedFile();
    void readPartition(void* host_addr, Partition* partition);
    void writePartition(const void* host_addr, Partition* partition);
};

/** One set of staging buffers for `n` partitions: pinned host memory (the file side) + HBM slots (the slab side), with its own copy
 *  stream; the two asynchronous blocks below are built on it. */
struct SwapStaging {
    vector<torch::Tensor> host;  // pinned, [partition_size, d] fp32 each
    vector<torch::Tensor> dev;   // HBM,    [partition_size, d] fp32 each
    c10::cuda::CUDAStream stream = c10::cuda::getDefaultCUDAStream();
    void init(int n, int64_t rows, int64_t d, torch::Device device);
};

/** LookaheadBlock (buffer.h:43-76, buffer.cpp:118-220) for an HBM slab: a reader thread brings the partitions of the NEXT admit set from
 *  the file into pinned memory and on into spare HBM slots over its copy stream while training runs on the current buffer state;
 *  move_to_buffer is then a device-to-device copy ordered on the trainer's stream. */
class LookaheadBlock {
    PartitionedFile* partitioned_file_;
    vector<Partition*> partitions_;
    SwapStaging staging_;
    torch::Device device_;
    at::cuda::CUDAEvent moved_event_;  // the last move_to_buffer's copies out of the HBM slots (recorded on the trainer's stream)
    std::mutex lock_;
    std::condition_variable cv_;
    std::thread* thread_ = nullptr;
    std::atomic<bool> present_{false}, done_{false};
    string error_;
    void run();

   public:
    LookaheadBlock(int64_t partition_rows, int64_t d, PartitionedFile* partitioned_file, int num_per_lookahead, torch::Device device);
    ~LookaheadBlock();
    void start(vector<Partition*> first_partitions);
    void stop();
    /** copies the prefetched partitions into `slab_addrs` (device pointers of the evicted slots), then starts prefetching `next_partitions` */
    void move_to_buffer(vector<torch::Tensor> slab_slots, vector<int64_t> buffer_idxs, vector<Partition*> next_partitions);
};

/** AsyncWriteBlock (buffer.h:78-110, buffer.cpp:222-322) for an HBM slab: async_write snapshots the evicted slots into spare HBM on the
 *  trainer's stream (so the slots are free for the admitted partitions at once); a writer thread drains them to pinned memory and the file. */
class AsyncWriteBlock {
    PartitionedFile* partitioned_file_;
    vector<Partition*> partitions_;
    SwapStaging staging_;
    torch::Device device_;
    at::cuda::CUDAEvent staged_event_;  // the snapshot copies of the last async_write (recorded on the trainer's stream)
    std::mutex lock_;
    std::condition_variable cv_;
    std::thread* thread_ = nullptr;
    std::atomic<bool> present_{false}, done_{false};
    string error_;
    void run();

   public:
    AsyncWriteBlock(int64_t partition_rows, int64_t d, PartitionedFile* partitioned_file, int num_per_evict, torch::Device device);
    ~AsyncWriteBlock();
    void start();
    void stop();
    void async_write(vector<Partition*> partitions, vector<torch::Tensor> slab_slots);
    void wait_idle();  // every queued write-back has reached the file
};

/** PartitionBuffer (storage/buffer.h:116-190) with the slab in HBM: `capacity` partitions resident on the GPU, the rest in the
 *  backing file; swaps follow the caller-supplied buffer-state sequence (BETA/COMET orderings are reused unchanged).  With
 *  prefetching the swap engine is asynchronous (LookaheadBlock / AsyncWriteBlock above): HBM holds capacity + 2 * fine_to_coarse_ratio
 *  partition slots, and performNextSwap costs two device-to-device partition copies instead of a disk + PCIe round trip. */
class PartitionBuffer {
    int capacity_, num_partitions_, fine_to_coarse_ratio_, embedding_size_;
    int64_t partition_size_, total_embeddings_;
    bool loaded_ = false, prefetching_;
    torch::Device device_;
    torch::Tensor buffer_tensor_view_;  // [capacity * partition_size, embedding_size] fp32 in HBM
    torch::Tensor staging_;             // pinned host bounce buffer, one partition (synchronous paths: load / sync / no prefetching)
    vector<Partition*> partition_table_;
    string filename_;
    PartitionedFile* partitioned_file_;
    LookaheadBlock* lookahead_block_ = nullptr;
    AsyncWriteBlock* async_write_block_ = nullptr;
    vector<int> buffer_state_;                // partition ids resident now, by position
    vector<vector<int>> buffer_states_;       // the whole ordering (cached on the host: no per-element tensor reads)
    size_t state_pos_ = 0;  // index of the NEXT state in buffer_states_
    void admit(vector<Partition*> admit_partitions, vector<int64_t> buffer_idxs);
    void evict(vector<Partition*> evict_partitions);
    torch::Tensor slot(int64_t buffer_idx, int64_t rows);
    void startThreads();
    void stopThreads();
    vector<int> admitOf(size_t next_pos, const vector<int>& current);

   public:
    PartitionBuffer(int capacity, int num_partitions, int fine_to_coarse_ratio, int64_t partition_size, int embedding_size, int64_t total_embeddings,
                    torch::Dtype dtype, string filename, bool prefetching, torch::Device device = torch::Device(torch::kCUDA, 0));
    ~PartitionBuffer();
    void load();
    void write();
    void unload(bool write);
    vector<int> getNextAdmit();
    vector<int> getNextEvict();
    Indices getRandomIds(int64_t size);
    torch::Tensor indexRead(torch::Tensor indices);
    torch::Tensor getGlobalToLocalMap(bool get_current);
    torch::Tensor getBufferState();
    void indexAdd(torch::Tensor indices, torch::Tensor values);
    void adagradUpdate(PartitionBuffer& state, torch::Tensor indices, torch::Tensor gradients, float learning_rate);
    void setBufferOrdering(vector<torch::Tensor> buffer_states);
    bool hasSwap();
    void performNextSwap();
    void sync();
    int64_t getNumInMemory() { return buffer_tensor_view_.defined() ? buffer_tensor_view_.size(0) : 0; }
    torch::Tensor bufferTensor() { return buffer_tensor_view_; }
};

/** storage/storage.h:90-145: the Storage facade GraphModelStorage holds for buffered node embeddings / optimizer state */
struct PartitionBufferOptions {  // configuration/options.h (the fields the buffer reads)
    int num_partitions = 1;
    int buffer_capacity = 1;
    bool prefetching = true;
    int fine_to_coarse_ratio = 1;
    torch::Dtype dtype = torch::kFloat32;
};

class PartitionBufferStorage : public Storage {
    bool loaded_ = false;

   public:
    PartitionBuffer* buffer_ = nullptr;
    shared_ptr<PartitionBufferOptions> options_;
    PartitionBufferStorage(string filename, int64_t dim0_size, int64_t dim1_size, shared_ptr<PartitionBufferOptions> options,
                           torch::Device device = torch::Device(torch::kCUDA, 0));
    PartitionBufferStorage(string filename, torch::Tensor data, shared_ptr<PartitionBufferOptions> options,
                           torch::Device device = torch::Device(torch::kCUDA, 0));
    ~PartitionBufferStorage();
    void rangePut(int64_t offset, torch::Tensor values);
    void append(torch::Tensor values);
    void load() override;
    void unload(bool perform_write) override;
    void write() override;
    torch::Tensor indexRead(Indices indices) override;
    void indexAdd(Indices indices, torch::Tensor values) override;
    torch::Tensor range(int64_t offset, int64_t n) override;
    void indexPut(Indices indices, torch::Tensor values) override;
    void rangePut(int64_t offset, int64_t n, torch::Tensor values) override;
    void shuffle() override;
    void sort(bool src) override;
    Indices getRandomIds(int64_t size) { return buffer_->getRandomIds(size); }
    bool hasSwap() { return buffer_->hasSwap(); }
    void performNextSwap() { buffer_->performNextSwap(); }
    torch::Tensor getGlobalToLocalMap(bool get_current) { return buffer_->getGlobalToLocalMap(get_current); }
    void sync() { buffer_->sync(); }
    void setBufferOrdering(vector<torch::Tensor> buffer_states) { buffer_->setBufferOrdering(buffer_states); }
    std::vector<int> getNextAdmit() { return buffer_->getNextAdmit(); }
    std::vector<int> getNextEvict() { return buffer_->getNextEvict(); }
    int64_t getNumInMemory() { return buffer_->getNumInMemory(); }
};

// shared helpers of storage.cpp / buffer.cpp
void mbh_check_indices(const Indices& indices);
void mbh_check_values(const torch::Tensor& table, const Indices& indices, const torch::Tensor& values);
torch::Tensor mbh_device_rows_read(const torch::Tensor& table, Indices indices);
void mbh_device_rows_scatter(torch::Tensor& table, Indices indices, torch::Tensor values, bool add);

// ---- decoders --------------------------------------------------------------------------------------------------
/** nn/decoders/edge/edge_decoder.h:13-31 restricted to the DotCompare family the kernels implement */
class EdgeDecoder : public torch::nn::Module {
   public:
    int decoder_kind_;  // mb_decoder
    torch::Tensor relations_;
    torch::Tensor inverse_relations_;
    int num_relations_;
    int embedding_size_;
    torch::TensorOptions tensor_options_;
    EdgeDecoderMethod decoder_method_;
    bool use_inverse_relations_;
    LearningTask learning_task_ = LearningTask::LINK_PREDICTION;

    torch::Tensor apply_relation(torch::Tensor nodes, torch::Tensor relations);
    torch::Tensor compute_scores(torch::Tensor src, torch::Tensor dst);
    torch::Tensor select_relations(torch::Tensor indices, bool inverse = false);
};

class DistMult : public EdgeDecoder {
   public:
    DistMult(int num_relations, int embedding_dim, torch::TensorOptions tensor_options = torch::TensorOptions(), bool use_inverse_relations = true,
             EdgeDecoderMethod decoder_method = EdgeDecoderMethod::CORRUPT_NODE);
    void reset();
};

class ComplEx : public EdgeDecoder {
   public:
    ComplEx(int num_relations, int embedding_dim, torch::TensorOptions tensor_options = torch::TensorOptions(), bool use_inverse_relations = true,
            EdgeDecoderMethod decoder_method = EdgeDecoderMethod::CORRUPT_NODE);
    void reset();
};

/** decoder_methods.h:11-21 */
std::tuple<torch::Tensor, torch::Tensor> only_pos_forward(shared_ptr<EdgeDecoder> decoder, torch::Tensor edges, torch::Tensor node_embeddings);
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> node_corrupt_forward(shared_ptr<EdgeDecoder> decoder, torch::Tensor positive_edges,
                                                                                            torch::Tensor node_embeddings, torch::Tensor dst_negs,
                                                                                            torch::Tensor src_negs);

// ---- loss ------------------------------------------------------------------------------------------------------
class LossFunction {
   public:
    virtual ~LossFunction() {}
    virtual torch::Tensor operator()(torch::Tensor y_pred, torch::Tensor targets, bool scores) = 0;
};

/** nn/loss.cpp:50-67 (libtorch ops: the generic autograd path; Model::train_batch uses the fused kernels instead) */
class SoftmaxCrossEntropy : public LossFunction {
   public:
    LossReduction reduction_type_;
    explicit SoftmaxCrossEntropy(LossReduction reduction) : reduction_type_(reduction) {}
    torch::Tensor operator()(torch::Tensor y_pred, torch::Tensor targets, bool scores) override;
};

// ---- batch -----------------------------------------------------------------------------------------------------
/** data/batch.h:32-89 (link-prediction fields) */
class Batch {
   public:
    int batch_id_ = 0;
    int64_t start_idx_ = 0;
    int batch_size_ = 0;
    bool train_;
    int device_id_ = -1;
    Indices unique_node_indices_;
    torch::Tensor node_embeddings_;
    torch::Tensor node_gradients_;
    torch::Tensor node_embeddings_state_;
    torch::Tensor node_state_update_;
    Indices src_neg_indices_mapping_;
    Indices dst_neg_indices_mapping_;
    torch::Tensor edges_;
    Indices src_neg_indices_;
    Indices dst_neg_indices_;
    torch::Tensor src_neg_filter_;
    torch::Tensor dst_neg_filter_;
    float loss_ = 0.f;  // (adapter: the batch's loss, recorded by the compute stage)

    explicit Batch(bool train);
    void to(torch::Device device);
    void accumulateGradients(float learning_rate);
    void embeddingsToHost();
    void clear();
};

// ---- reporting -------------------------------------------------------------------------------------------------
/** reporting/reporting.h:20-58: ranking metrics over the collected rank vector */
class RankingMetric {
   public:
    string name_;
    string unit_;
    virtual ~RankingMetric() {}
    virtual torch::Tensor computeMetric(torch::Tensor ranks) = 0;
};
class HitskMetric : public RankingMetric {
   public:
    int k_;
    explicit HitskMetric(int k);
    torch::Tensor computeMetric(torch::Tensor ranks) override;
};
class MeanRankMetric : public RankingMetric {
   public:
    MeanRankMetric();
    torch::Tensor computeMetric(torch::Tensor ranks) override;
};
class MeanReciprocalRankMetric : public RankingMetric {
   public:
    MeanReciprocalRankMetric();
    torch::Tensor computeMetric(torch::Tensor ranks) override;
};

/** reporting/reporting.h:76-100 / reporting.cpp:44-95: ranks are computed on the device by mb_compute_ranks and stay there until report() */
class LinkPredictionReporter {
   public:
    std::vector<shared_ptr<RankingMetric>> metrics_;
    std::vector<torch::Tensor> per_batch_ranks_;
    std::vector<torch::Tensor> per_batch_scores_;
    std::vector<torch::Tensor> per_batch_edges_;
    torch::Tensor all_ranks_;
    torch::Tensor all_scores_;
    std::mutex lock_;

    void addMetric(shared_ptr<RankingMetric> metric) { metrics_.emplace_back(metric); }
    void clear();
    torch::Tensor computeRanks(torch::Tensor pos_scores, torch::Tensor neg_scores);
    void addResult(torch::Tensor pos_scores, torch::Tensor neg_scores, torch::Tensor edges = torch::Tensor());
    /** adds rank vectors computed elsewhere (Model::evaluate_batch's fused scores -> filter -> ranks call) */
    void addRanks(torch::Tensor ranks);
    /** concatenates the collected ranks into all_ranks_ (CPU) and returns "name: value" lines in the reference's format */
    string report();
};

// ---- model -----------------------------------------------------------------------------------------------------
/** nn/model.h:16-63, link-prediction path with a pure-embedding encoder (EmbeddingLayer::forward is a view, embedding.cpp:17) */
class Model : public torch::nn::Module {
   public:
    shared_ptr<EdgeDecoder> decoder_;
    shared_ptr<LossFunction> loss_function_;
    torch::Device device_;
    LearningTask learning_task_ = LearningTask::LINK_PREDICTION;
    float sparse_lr_ = 0.1f;
    float dense_lr_ = 0.1f;                       // learning rate of the dense (relation) Adagrad optimizer
    std::vector<torch::Tensor> dense_state_;      // Adagrad "sum" per named parameter (optim.cpp:100-112)

    Model(shared_ptr<EdgeDecoder> decoder, shared_ptr<LossFunction> loss, torch::Device device);
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> forward_lp(shared_ptr<Batch> batch, bool train);
    void train_batch(shared_ptr<Batch> batch, bool call_step = true);
    /** train_batch + DataLoader::updateEmbeddings(batch, gpu=true) fused on device-resident tables (pipeline_gpu.cpp:49-91) */
    float train_batch_fused(shared_ptr<Batch> batch, InMemory& embeddings, InMemory& state, bool call_step = true);
    /** nn/model.cpp:335-349: scores (+ the batch's score filters) -> ranks of both corruption sides, handed to reporter_ */
    shared_ptr<LinkPredictionReporter> reporter_;
    void evaluate_batch(shared_ptr<Batch> batch);
    void clear_grad();
    void step();
};

// ---- compute stage (pipeline/pipeline_gpu.cpp:33-104, pipeline/queue.h) ----------------------------------------------------------
/** pipeline/queue.h: bounded blocking queue between pipeline stages */
template <class T>
class Queue {
    std::deque<T> q_;
    size_t max_size_;
    std::mutex m_;
    std::condition_variable cv_;
    bool open_ = true;

   public:
    explicit Queue(size_t max_size) : max_size_(max_size) {}
    void blocking_push(T item) {
        std::unique_lock<std::mutex> lock(m_);
        cv_.wait(lock, [this] { return q_.size() < max_size_ || !open_; });
        q_.push_back(std::move(item));
        lock.unlock();
        cv_.notify_all();
    }
    std::tuple<bool, T> blocking_pop() {
        std::unique_lock<std::mutex> lock(m_);
        cv_.wait(lock, [this] { return !q_.empty() || !open_; });
        if (q_.empty()) return std::forward_as_tuple(false, T());
        T item = std::move(q_.front());
        q_.pop_front();
        lock.unlock();
        cv_.notify_all();
        return std::forward_as_tuple(true, std::move(item));
    }
    void close() {  // wake every waiter; pops drain what is left, then report "not popped"
        {
            std::lock_guard<std::mutex> lock(m_);
            open_ = false;
        }
        cv_.notify_all();
    }
    size_t size() {
        std::lock_guard<std::mutex> lock(m_);
        return q_.size();
    }
};

/** ComputeWorkerGPU::run with device-resident embeddings (pipeline_gpu.cpp:49-91): a worker thread pops loaded batches, and for each
 *  does loadGPUParameters -> Model::train_batch -> updateEmbeddings(batch, gpu = true) as ONE fused C-ABI call
 *  (Model::train_batch_fused) on its device's tables, then hands the batch (its loss recorded) to the update-batches queue.  The
 *  reference's transfer / host-update stages have nothing left to do when the table lives in HBM. */
class ComputeWorkerGPU {
    shared_ptr<Model> model_;
    shared_ptr<InMemory> embeddings_, state_;
    std::thread* thread_ = nullptr;
    std::atomic<bool> done_{false};
    string error_;
    void run();

   public:
    shared_ptr<Queue<shared_ptr<Batch>>> device_loaded_batches_;
    shared_ptr<Queue<shared_ptr<Batch>>> device_update_batches_;
    std::atomic<int64_t> edges_processed_{0};
    std::atomic<int64_t> batches_processed_{0};
    ComputeWorkerGPU(shared_ptr<Model> model, shared_ptr<InMemory> embeddings, shared_ptr<InMemory> state, size_t queue_size = 4);
    ~ComputeWorkerGPU();
    void start();
    void stop();  // closes the input queue, lets the worker drain it, joins
    const string& error() const { return error_; }
};
