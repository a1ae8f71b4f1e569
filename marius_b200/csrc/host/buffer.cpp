// buffer.cpp -- PartitionBuffer (HBM slab over a partitioned backing file) with the asynchronous swap engine, and the
// PartitionBufferStorage facade.  Public surface, error conventions, slot assignment (an admitted partition takes the slot of the
// partition evicted in the same swap) and file format follow the reference's storage/buffer.{h,cpp} and storage/storage.cpp:67-201;
// what differs is where the bytes live and how they move:
//
//   reference (host slab)                                   here (HBM slab, 180 GB per B200)
//   LookaheadBlock: pread into spare host memory while       reader thread: pread -> pinned host -> spare HBM slots over a copy stream while
//   training runs, memcpy into the slab at the swap          training runs; the swap is a device-to-device copy on the trainer's stream
//   AsyncWriteBlock: memcpy the evicted slots out, a         the evicted slots are snapshotted into spare HBM on the trainer's stream (the
//   thread pwrite()s them                                    slots are free at once), a writer thread drains them: D2H on its stream, pwrite
//
// A swap therefore costs two partition-sized HBM copies (a few ms at 8.6 GB / partition) instead of a disk + PCIe round trip, and
// never blocks on the file unless the prefetch / write-back of the previous swap is still running.
#include <fcntl.h>
#include <unistd.h>

#include <cmath>
#include <fstream>

#include "marius_host.h"

// ---- PartitionedFile -------------------------------------------------------------------------------------------
PartitionedFile::PartitionedFile(string filename) : filename_(filename) {
    fd_ = open(filename_.c_str(), O_RDWR);
    if (fd_ == -1) throw MariusRuntimeException("Unable to open " + filename_);
}
PartitionedFile::~PartitionedFile() {
    if (fd_ != -1) close(fd_);
}
void PartitionedFile::readPartition(void* host_addr, Partition* p) {
    if (host_addr == nullptr || p == nullptr) throw std::runtime_error("");  // buffer.cpp:75-79
    int64_t done = 0;
    while (done < p->total_size_) {
        ssize_t r = pread(fd_, (char*)host_addr + done, std::min<int64_t>(p->total_size_ - done, (int64_t)1 << 30), p->file_offset_ + done);
        if (r <= 0) throw MariusRuntimeException("short partition read");
        done += r;
    }
}
void PartitionedFile::writePartition(const void* host_addr, Partition* p) {
    if (host_addr == nullptr || p == nullptr) throw std::runtime_error("");
    int64_t done = 0;
    while (done < p->total_size_) {
        ssize_t r = pwrite(fd_, (const char*)host_addr + done, std::min<int64_t>(p->total_size_ - done, (int64_t)1 << 30), p->file_offset_ + done);
        if (r <= 0) throw MariusRuntimeException("short partition write");
        done += r;
    }
}

// ---- staging ---------------------------------------------------------------------------------------------------
void SwapStaging::init(int n, int64_t rows, int64_t d, torch::Device device) {
    host.clear();
    dev.clear();
    for (int i = 0; i < n; i++) {
        host.push_back(torch::empty({rows, d}, torch::TensorOptions().dtype(torch::kFloat32).pinned_memory(true)));
        dev.push_back(torch::empty({rows, d}, torch::TensorOptions().dtype(torch::kFloat32).device(device)));
    }
    stream = c10::cuda::getStreamFromPool(false, device.index());
}

// ---- LookaheadBlock --------------------------------------------------------------------------------------------
LookaheadBlock::LookaheadBlock(int64_t partition_rows, int64_t d, PartitionedFile* partitioned_file, int num_per_lookahead, torch::Device device)
    : partitioned_file_(partitioned_file), device_(device) {
    staging_.init(num_per_lookahead, partition_rows, d, device);
    c10::cuda::CUDAGuard guard(device_);
    moved_event_.record(c10::cuda::getCurrentCUDAStream(device_.index()));  // "nothing to wait for"
}

LookaheadBlock::~LookaheadBlock() { stop(); }

void LookaheadBlock::run() {
    c10::cuda::CUDAGuard guard(device_);
    while (!done_) {
        std::unique_lock<std::mutex> lock(lock_);
        cv_.wait(lock, [this] { return done_.load() || (!present_.load() && !partitions_.empty()); });  // wait until the block is empty and has work
        if (done_) break;
        try {
            // the HBM slots may still be read by the previous move_to_buffer (device-to-device copies on the trainer's stream)
            moved_event_.synchronize();
            for (size_t i = 0; i < partitions_.size(); i++) {
                Partition* partition = partitions_[i];
                {
                    // a write-back of this partition may still be in flight: the file holds stale rows until it lands (buffer.cpp:160-163)
                    std::unique_lock<std::mutex> plock(*partition->lock_);
                    partition->cv_->wait(plock, [partition] { return partition->evicting_ == false; });
                    partitioned_file_->readPartition(staging_.host[i].data_ptr<float>(), partition);
                }
                partition->cv_->notify_all();
                c10::cuda::CUDAStreamGuard sg(staging_.stream);
                staging_.dev[i].narrow(0, 0, partition->partition_size_).copy_(staging_.host[i].narrow(0, 0, partition->partition_size_), /*non_blocking=*/true);
            }
            staging_.stream.synchronize();  // the partitions are in HBM; the pinned buffers are free again
        } catch (const std::exception& e) {
            error_ = e.what();
        }
        present_ = true;
        lock.unlock();
        cv_.notify_all();
    }
}

void LookaheadBlock::start(vector<Partition*> first_partitions) {
    {
        std::lock_guard<std::mutex> lock(lock_);
        partitions_ = first_partitions;
    }
    if (thread_ == nullptr) thread_ = new std::thread(&LookaheadBlock::run, this);
    cv_.notify_all();
}

void LookaheadBlock::stop() {
    if (thread_ != nullptr) {
        {
            std::lock_guard<std::mutex> lock(lock_);
            done_ = true;
        }
        cv_.notify_all();
        if (thread_->joinable()) thread_->join();
        delete thread_;
        thread_ = nullptr;
    }
}

void LookaheadBlock::move_to_buffer(vector<torch::Tensor> slab_slots, vector<int64_t> buffer_idxs, vector<Partition*> next_partitions) {
    if (partitions_.size() > slab_slots.size() || partitions_.size() > buffer_idxs.size()) throw std::runtime_error("");  // buffer.cpp:187-190
    std::unique_lock<std::mutex> lock(lock_);
    cv_.wait(lock, [this] { return present_.load() || partitions_.empty(); });  // wait until the block is populated
    if (!error_.empty()) throw MariusRuntimeException("lookahead read failed: " + error_);
    auto stream = c10::cuda::getCurrentCUDAStream(device_.index());
    for (size_t i = 0; i < partitions_.size(); i++) {
        Partition* partition = partitions_[i];
        slab_slots[i].narrow(0, 0, partition->partition_size_).copy_(staging_.dev[i].narrow(0, 0, partition->partition_size_), /*non_blocking=*/true);
        partition->buffer_idx_ = (int)buffer_idxs[i];
        partition->present_ = true;
    }
    moved_event_.record(stream);
    // the next partitions are prefetched automatically (buffer.cpp:213-216)
    partitions_ = next_partitions;
    present_ = false;
    lock.unlock();
    cv_.notify_all();
}

// ---- AsyncWriteBlock -------------------------------------------------------------------------------------------
AsyncWriteBlock::AsyncWriteBlock(int64_t partition_rows, int64_t d, PartitionedFile* partitioned_file, int num_per_evict, torch::Device device)
    : partitioned_file_(partitioned_file), device_(device) {
    staging_.init(num_per_evict, partition_rows, d, device);
}

AsyncWriteBlock::~AsyncWriteBlock() { stop(); }

void AsyncWriteBlock::run() {
    c10::cuda::CUDAGuard guard(device_);
    while (true) {
        std::unique_lock<std::mutex> lock(lock_);
        cv_.wait(lock, [this] { return present_.load() || done_.load(); });
        if (!present_ && done_) return;
        try {
            staged_event_.synchronize();  // the snapshots of the evicted slots are complete
            for (size_t i = 0; i < partitions_.size(); i++) {
                Partition* partition = partitions_[i];
                {
                    c10::cuda::CUDAStreamGuard sg(staging_.stream);
                    staging_.host[i].narrow(0, 0, partition->partition_size_).copy_(staging_.dev[i].narrow(0, 0, partition->partition_size_), /*non_blocking=*/true);
                }
                staging_.stream.synchronize();
                partitioned_file_->writePartition(staging_.host[i].data_ptr<float>(), partition);
                {
                    std::lock_guard<std::mutex> plock(*partition->lock_);
                    partition->evicting_ = false;
                }
                partition->cv_->notify_all();
            }
        } catch (const std::exception& e) {
            error_ = e.what();
            for (auto partition : partitions_) {
                std::lock_guard<std::mutex> plock(*partition->lock_);
                partition->evicting_ = false;
                partition->cv_->notify_all();
            }
        }
        present_ = false;
        lock.unlock();
        cv_.notify_all();
    }
}

void AsyncWriteBlock::start() {
    if (thread_ == nullptr) thread_ = new std::thread(&AsyncWriteBlock::run, this);
}

void AsyncWriteBlock::stop() {
    if (thread_ != nullptr) {
        {
            std::lock_guard<std::mutex> lock(lock_);
            done_ = true;
        }
        cv_.notify_all();
        if (thread_->joinable()) thread_->join();  // (drains a pending write first)
        delete thread_;
        thread_ = nullptr;
    }
}

void AsyncWriteBlock::async_write(vector<Partition*> partitions, vector<torch::Tensor> slab_slots) {
    if (partitions.size() > staging_.dev.size() || partitions.size() > slab_slots.size()) throw std::runtime_error("");  // buffer.cpp:296-299
    std::unique_lock<std::mutex> lock(lock_);
    cv_.wait(lock, [this] { return present_ == false; });  // wait until the block is empty
    if (!error_.empty()) throw MariusRuntimeException("asynchronous partition write failed: " + error_);
    partitions_ = partitions;
    auto stream = c10::cuda::getCurrentCUDAStream(device_.index());
    for (size_t i = 0; i < partitions_.size(); i++) {
        Partition* partition = partitions_[i];
        staging_.dev[i].narrow(0, 0, partition->partition_size_).copy_(slab_slots[i].narrow(0, 0, partition->partition_size_), /*non_blocking=*/true);
        std::lock_guard<std::mutex> plock(*partition->lock_);
        partition->evicting_ = true;
    }
    staged_event_.record(stream);
    present_ = true;
    lock.unlock();
    cv_.notify_all();
}

void AsyncWriteBlock::wait_idle() {
    std::unique_lock<std::mutex> lock(lock_);
    cv_.wait(lock, [this] { return present_ == false; });
    if (!error_.empty()) throw MariusRuntimeException("asynchronous partition write failed: " + error_);
}

// ---- PartitionBuffer -------------------------------------------------------------------------------------------
PartitionBuffer::PartitionBuffer(int capacity, int num_partitions, int fine_to_coarse_ratio, int64_t partition_size, int embedding_size,
                                 int64_t total_embeddings, torch::Dtype dtype, string filename, bool prefetching, torch::Device device)
    : capacity_(capacity), num_partitions_(num_partitions), fine_to_coarse_ratio_(fine_to_coarse_ratio), embedding_size_(embedding_size),
      partition_size_(partition_size), total_embeddings_(total_embeddings), prefetching_(prefetching), device_(device), filename_(filename) {
    if (dtype != torch::kFloat32) throw MariusRuntimeException("PartitionBuffer holds fp32 embeddings (buffer.cpp:447,467-470 assume float)");
    if (!device_.is_cuda()) throw MariusRuntimeException("marius_b200 PartitionBuffer keeps its slab in HBM: a CUDA device is required");
    if (!device_.has_index()) device_ = torch::Device(torch::kCUDA, 0);
    int64_t idx = 0, off = 0;
    for (int i = 0; i < num_partitions_; i++) {
        int64_t sz = (i == num_partitions_ - 1) ? total_embeddings_ - idx : partition_size_;  // buffer.cpp:347-350
        partition_table_.push_back(new Partition(i, sz, embedding_size_, idx, off));
        idx += sz;
        off += sz * embedding_size_ * 4;
    }
    partitioned_file_ = new PartitionedFile(filename_);
}

PartitionBuffer::~PartitionBuffer() {
    try {
        unload(true);
    } catch (...) {
    }
    delete partitioned_file_;
    for (auto p : partition_table_) delete p;
}

torch::Tensor PartitionBuffer::slot(int64_t buffer_idx, int64_t rows) { return buffer_tensor_view_.narrow(0, buffer_idx * partition_size_, rows); }

vector<int> PartitionBuffer::admitOf(size_t next_pos, const vector<int>& current) {
    vector<int> out;  // partitions of state `next_pos` that are not in `current` (buffer.cpp:543-561)
    if (next_pos >= buffer_states_.size()) return out;
    for (int id : buffer_states_[next_pos]) {
        bool found = false;
        for (int c : current) found |= (c == id);
        if (!found) out.push_back(id);
    }
    return out;
}

void PartitionBuffer::load() {
    if (loaded_) return;
    if (buffer_states_.empty()) throw MariusRuntimeException("setBufferOrdering must be called before load");
    c10::cuda::CUDAGuard guard(device_);
    buffer_tensor_view_ = torch::zeros({capacity_ * partition_size_, (int64_t)embedding_size_}, torch::TensorOptions().dtype(torch::kFloat32).device(device_));
    staging_ = torch::empty({partition_size_, (int64_t)embedding_size_}, torch::TensorOptions().dtype(torch::kFloat32).pinned_memory(true));
    for (size_t i = 0; i < buffer_state_.size(); i++) {  // buffer.cpp:383-392
        Partition* p = partition_table_[buffer_state_[i]];
        partitioned_file_->readPartition(staging_.data_ptr<float>(), p);
        slot((int64_t)i, p->partition_size_).copy_(staging_.narrow(0, 0, p->partition_size_));
        p->present_ = true;
        p->buffer_idx_ = (int)i;
    }
    if (prefetching_) {  // buffer.cpp:406-410
        lookahead_block_ = new LookaheadBlock(partition_size_, embedding_size_, partitioned_file_, fine_to_coarse_ratio_, device_);
        async_write_block_ = new AsyncWriteBlock(partition_size_, embedding_size_, partitioned_file_, fine_to_coarse_ratio_, device_);
        startThreads();
    }
    loaded_ = true;
}

void PartitionBuffer::startThreads() {  // buffer.cpp:698-713
    vector<Partition*> first;
    for (int id : getNextAdmit()) first.push_back(partition_table_[id]);
    lookahead_block_->start(first);
    async_write_block_->start();
}

void PartitionBuffer::stopThreads() {
    if (lookahead_block_) lookahead_block_->stop();
    if (async_write_block_) async_write_block_->stop();
}

void PartitionBuffer::write() { sync(); }

void PartitionBuffer::unload(bool write) {
    if (!loaded_) return;
    if (prefetching_) {
        stopThreads();  // (the writer drains its queue before it exits)
        delete lookahead_block_;
        delete async_write_block_;
        lookahead_block_ = nullptr;
        async_write_block_ = nullptr;
    }
    if (write) sync();
    for (auto p : partition_table_) {
        p->present_ = false;
        p->buffer_idx_ = -1;
    }
    buffer_tensor_view_ = torch::Tensor();
    staging_ = torch::Tensor();
    loaded_ = false;
}

vector<int> PartitionBuffer::getNextAdmit() { return admitOf(state_pos_, buffer_state_); }

vector<int> PartitionBuffer::getNextEvict() {
    vector<int> out;  // buffer.cpp:563-579
    if (state_pos_ >= buffer_states_.size()) return out;
    const auto& next = buffer_states_[state_pos_];
    for (int id : buffer_state_) {
        bool found = false;
        for (int n : next) found |= (n == id);
        if (!found) out.push_back(id);
    }
    return out;
}

torch::Tensor PartitionBuffer::getBufferState() { return torch::tensor(buffer_state_, torch::kInt64); }

Indices PartitionBuffer::getRandomIds(int64_t size) {
    int64_t n = 0;
    for (int id : buffer_state_) n += partition_table_[id]->partition_size_;
    return torch::randint(n, {size}, torch::kInt64);  // buffer.cpp:457
}

torch::Tensor PartitionBuffer::indexRead(torch::Tensor indices) {
    mbh_check_indices(indices);
    return mbh_device_rows_read(buffer_tensor_view_, indices);
}

void PartitionBuffer::indexAdd(torch::Tensor indices, torch::Tensor values) {
    mbh_check_values(buffer_tensor_view_, indices, values);
    mbh_device_rows_scatter(buffer_tensor_view_, indices, values, true);
}

void PartitionBuffer::adagradUpdate(PartitionBuffer& state, torch::Tensor indices, torch::Tensor gradients, float learning_rate) {
    mbh_check_values(buffer_tensor_view_, indices, gradients);
    auto& t = buffer_tensor_view_;
    auto& s = state.buffer_tensor_view_;
    if (!s.defined() || s.sizes() != t.sizes()) throw std::runtime_error("");
    indices = indices.to(t.device()).contiguous();
    gradients = gradients.to(t.device()).contiguous();
    mb_throw_on_error(mb_adagrad_update_rows(t.data_ptr<float>(), s.data_ptr<float>(), t.size(0), t.stride(0), t.size(1), indices.data_ptr<int64_t>(),
                                             indices.size(0), gradients.data_ptr<float>(), gradients.stride(0), learning_rate,
                                             mb_current_stream(t.device())));
}

torch::Tensor PartitionBuffer::getGlobalToLocalMap(bool get_current) {
    vector<int32_t> parts, slots;
    if (get_current) {
        for (int id : buffer_state_) {
            Partition* p = partition_table_[id];
            parts.push_back(p->partition_id_);
            slots.push_back(p->buffer_idx_);
        }
    } else {
        // mapping after the next swap: survivors keep their slot, admitted partitions take the evicted slots (buffer.cpp:603-630)
        if (state_pos_ >= buffer_states_.size()) throw MariusRuntimeException("no next buffer state");
        auto evict = getNextEvict();
        auto admit = getNextAdmit();
        for (int id : buffer_states_[state_pos_]) {
            Partition* p = partition_table_[id];
            if (p->buffer_idx_ != -1) {
                parts.push_back(p->partition_id_);
                slots.push_back(p->buffer_idx_);
            }
        }
        for (size_t i = 0; i < evict.size() && i < admit.size(); i++) {
            parts.push_back(admit[i]);
            slots.push_back(partition_table_[evict[i]]->buffer_idx_);
        }
    }
    auto map = torch::empty({total_embeddings_}, torch::TensorOptions().dtype(torch::kInt64).device(device_));
    mb_throw_on_error(mb_global_to_local_map(map.data_ptr<int64_t>(), total_embeddings_, partition_size_, parts.data(), slots.data(), (int)parts.size(),
                                             mb_current_stream(device_)));
    return map;
}

void PartitionBuffer::setBufferOrdering(vector<torch::Tensor> buffer_states) {
    if (buffer_states.empty()) throw MariusRuntimeException("empty buffer ordering");
    buffer_states_.clear();
    for (auto& t : buffer_states) {  // the ordering is cached on the host once: admit / evict sets are plain integer loops afterwards
        auto c = t.to(torch::kCPU).to(torch::kInt64).contiguous();
        const int64_t* ptr = c.data_ptr<int64_t>();
        buffer_states_.emplace_back(ptr, ptr + c.numel());
    }
    state_pos_ = 0;
    buffer_state_ = buffer_states_[state_pos_++];
    if (loaded_) {  // buffer.cpp:487-490
        unload(true);
        load();
    }
}

bool PartitionBuffer::hasSwap() { return state_pos_ < buffer_states_.size(); }

void PartitionBuffer::evict(vector<Partition*> evict_partitions) {
    if (prefetching_) {
        vector<torch::Tensor> slots;
        for (auto p : evict_partitions) slots.push_back(slot(p->buffer_idx_, partition_size_));
        async_write_block_->async_write(evict_partitions, slots);
    } else {
        for (auto p : evict_partitions) {  // HBM -> pinned host -> file
            staging_.narrow(0, 0, p->partition_size_).copy_(slot(p->buffer_idx_, p->partition_size_));
            partitioned_file_->writePartition(staging_.data_ptr<float>(), p);
        }
    }
    for (auto p : evict_partitions) p->present_ = false;
}

void PartitionBuffer::admit(vector<Partition*> admit_partitions, vector<int64_t> buffer_idxs) {
    if (admit_partitions.size() > buffer_idxs.size()) throw std::runtime_error("");  // buffer.cpp:653-656
    if (prefetching_) {
        vector<torch::Tensor> slots;
        for (auto idx : buffer_idxs) slots.push_back(slot(idx, partition_size_));
        vector<Partition*> next_partitions;
        for (int id : admitOf(state_pos_, buffer_state_)) next_partitions.push_back(partition_table_[id]);  // buffer.cpp:667-674
        lookahead_block_->move_to_buffer(slots, buffer_idxs, next_partitions);
    } else {
        for (size_t i = 0; i < admit_partitions.size(); i++) {
            Partition* p = admit_partitions[i];
            partitioned_file_->readPartition(staging_.data_ptr<float>(), p);
            slot(buffer_idxs[i], p->partition_size_).copy_(staging_.narrow(0, 0, p->partition_size_));
            p->present_ = true;
            p->buffer_idx_ = (int)buffer_idxs[i];
        }
    }
}

void PartitionBuffer::performNextSwap() {
    if (buffer_state_.empty() || state_pos_ >= buffer_states_.size()) return;
    c10::cuda::CUDAGuard guard(device_);
    auto evict_ids = getNextEvict();
    auto admit_ids = getNextAdmit();
    vector<Partition*> admit_partitions, evict_partitions;
    vector<int64_t> evict_buffer_idxs;
    for (int id : admit_ids) admit_partitions.push_back(partition_table_[id]);
    for (int id : evict_ids) {
        evict_partitions.push_back(partition_table_[id]);
        evict_buffer_idxs.push_back(partition_table_[id]->buffer_idx_);
    }
    buffer_state_ = buffer_states_[state_pos_++];
    evict(evict_partitions);
    for (auto p : evict_partitions) p->buffer_idx_ = -1;
    admit(admit_partitions, evict_buffer_idxs);
}

void PartitionBuffer::sync() {
    if (!loaded_) return;
    c10::cuda::CUDAGuard guard(device_);
    if (async_write_block_) async_write_block_->wait_idle();  // queued write-backs reach the file first (they hold older rows of other partitions)
    for (auto p : partition_table_) {  // buffer.cpp:685-696
        if (p->present_) {
            staging_.narrow(0, 0, p->partition_size_).copy_(slot(p->buffer_idx_, p->partition_size_));
            partitioned_file_->writePartition(staging_.data_ptr<float>(), p);
        }
    }
}

// ---- PartitionBufferStorage (storage.cpp:67-201) ---------------------------------------------------------------
PartitionBufferStorage::PartitionBufferStorage(string filename, int64_t dim0_size, int64_t dim1_size, shared_ptr<PartitionBufferOptions> options,
                                               torch::Device device) {
    filename_ = filename;
    dim0_size_ = dim0_size;
    dim1_size_ = dim1_size;
    options_ = options;
    dtype_ = options_->dtype;
    initialized_ = true;
    device_ = device;  // dataloader.cpp:507,531,552,558 route on device_: the buffered table is device-resident here
    int64_t partition_size = (int64_t)std::ceil((double)dim0_size_ / options_->num_partitions);
    buffer_ = new PartitionBuffer(options_->buffer_capacity, options_->num_partitions, options_->fine_to_coarse_ratio, partition_size, (int)dim1_size_,
                                  dim0_size_, dtype_, filename_, options_->prefetching, device_);
}

PartitionBufferStorage::PartitionBufferStorage(string filename, torch::Tensor data, shared_ptr<PartitionBufferOptions> options, torch::Device device) {
    filename_ = filename;
    dim0_size_ = 0;
    dim1_size_ = data.size(1);
    options_ = options;
    dtype_ = options_->dtype;
    append(data);
    initialized_ = true;
    device_ = device;
    int64_t partition_size = (int64_t)std::ceil((double)dim0_size_ / options_->num_partitions);
    buffer_ = new PartitionBuffer(options_->buffer_capacity, options_->num_partitions, options_->fine_to_coarse_ratio, partition_size, (int)dim1_size_,
                                  dim0_size_, dtype_, filename_, options_->prefetching, device_);
}

PartitionBufferStorage::~PartitionBufferStorage() { delete buffer_; }

void PartitionBufferStorage::rangePut(int64_t offset, torch::Tensor values) {  // storage.cpp:112-128: straight into the file
    int fd = open(filename_.c_str(), O_RDWR);
    if (fd == -1) throw std::runtime_error("");
    auto host = values.to(torch::kCPU).to(torch::kFloat32).contiguous();
    int64_t bytes = host.numel() * 4, done = 0, off = offset * dim1_size_ * 4;
    while (done < bytes) {
        ssize_t r = pwrite(fd, (const char*)host.data_ptr() + done, std::min<int64_t>(bytes - done, (int64_t)1 << 30), off + done);
        if (r <= 0) {
            close(fd);
            throw std::runtime_error("");
        }
        done += r;
    }
    close(fd);
}

void PartitionBufferStorage::append(torch::Tensor values) {  // storage.cpp:130-150
    auto host = values.to(torch::kCPU).to(torch::kFloat32).contiguous();
    std::ofstream out(filename_, dim0_size_ == 0 ? (std::ios::trunc | std::ios::binary) : (std::ios::binary | std::ios::app));
    if (!out) throw std::runtime_error("");
    dim0_size_ += host.size(0);
    dim1_size_ = host.size(1);
    dtype_ = torch::kFloat32;
    out.write((const char*)host.data_ptr(), host.numel() * 4);
    out.close();
}

void PartitionBufferStorage::load() {
    if (!loaded_ && initialized_) {
        buffer_->load();
        loaded_ = true;
    }
}
void PartitionBufferStorage::write() {
    if (loaded_) buffer_->sync();
}
void PartitionBufferStorage::unload(bool perform_write) {
    if (loaded_) {
        buffer_->unload(perform_write);
        loaded_ = false;
    }
}
torch::Tensor PartitionBufferStorage::indexRead(Indices indices) { return buffer_->indexRead(indices); }
void PartitionBufferStorage::indexAdd(Indices indices, torch::Tensor values) { buffer_->indexAdd(indices, values); }
torch::Tensor PartitionBufferStorage::range(int64_t, int64_t) { throw std::runtime_error(""); }             // storage.cpp:178-181
void PartitionBufferStorage::indexPut(Indices, torch::Tensor) { throw std::runtime_error(""); }              // storage.cpp:183-186
void PartitionBufferStorage::rangePut(int64_t, int64_t, torch::Tensor) { throw std::runtime_error(""); }     // storage.cpp:188-191
void PartitionBufferStorage::shuffle() { throw std::runtime_error(""); }                                     // storage.cpp:193-196
void PartitionBufferStorage::sort(bool) { throw std::runtime_error(""); }                                    // storage.cpp:198-201
