// storage.cpp -- Storage / InMemory / PartitionBuffer adapters over the C-ABI gather / scatter kernels.
// Same public surface and error conventions as the reference's storage/storage.{h,cpp} and storage/buffer.{h,cpp};
// the table lives in HBM, files keep the reference's flat fp32 row-major format.
#include <c10/cuda/CUDAStream.h>
#include <fcntl.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <thread>

#include "marius_host.h"

// ---- glue ------------------------------------------------------------------------------------------------------
static int g_default_precision = MB_PREC_BF16X3;
void mb_set_default_precision(int precision) { g_default_precision = precision; }
int mb_default_precision() { return g_default_precision; }

void mb_throw_on_error(int status) {
    if (status == MB_OK) return;
    string msg = mb_last_error();
    if (status == MB_ERR_INVALID) throw std::runtime_error(msg);  // the reference throws std::runtime_error("") (storage.cpp:607-610)
    throw MariusRuntimeException(msg);
}

void* mb_current_stream(const torch::Device& device) {
    if (!device.is_cuda()) throw MariusRuntimeException("marius_b200: the embedding hot path runs on CUDA devices only (no CPU fallback)");
    return (void*)c10::cuda::getCurrentCUDAStream(device.index()).stream();
}

mb_context* mb_context_for(const torch::Device& device) {
    if (!device.is_cuda()) throw MariusRuntimeException("marius_b200: the embedding hot path runs on CUDA devices only (no CPU fallback)");
    // A context (workspace, side streams, captured step graph) serves one host thread at a time.  Threads check contexts out of a
    // process-wide pool and hand them back when they exit; the thread-exit hook makes NO CUDA call (the autograd engine's worker
    // threads are torn down after the CUDA runtime's own thread state, so destroying device resources there is a use-after-free),
    // and the pool itself is never destroyed: the driver reclaims everything at process exit.
    struct Pool {
        std::mutex mu;
        std::multimap<int, mb_context*> idle;
    };
    static Pool* pool = new Pool();
    struct Holder {
        std::map<int, mb_context*> ctx;
        ~Holder() {
            std::lock_guard<std::mutex> lock(pool->mu);
            for (auto& kv : ctx) pool->idle.emplace(kv.first, kv.second);
        }
    };
    static thread_local Holder holder;
    int idx = device.has_index() ? device.index() : 0;
    auto it = holder.ctx.find(idx);
    if (it != holder.ctx.end()) return it->second;
    mb_context* c = nullptr;
    {
        std::lock_guard<std::mutex> lock(pool->mu);
        auto idle = pool->idle.find(idx);
        if (idle != pool->idle.end()) {
            c = idle->second;
            pool->idle.erase(idle);
        }
    }
    if (c == nullptr) mb_throw_on_error(mb_create(idx, &c));
    holder.ctx[idx] = c;
    return c;
}

static void check_indices(const Indices& indices) {
    if (!indices.defined() || indices.sizes().size() != 1) throw std::runtime_error("");  // storage.cpp:607-610, buffer.cpp:442-445
}

static void check_values(const torch::Tensor& table, const Indices& indices, const torch::Tensor& values) {
    // storage.cpp:652-655, buffer.cpp:461-464
    if (!values.defined() || !indices.defined() || indices.sizes().size() != 1 || values.sizes().size() != 2 || indices.size(0) != values.size(0) ||
        table.size(1) != values.size(1)) {
        throw std::runtime_error("");
    }
}

static torch::Tensor device_rows_read(const torch::Tensor& table, Indices indices) {
    indices = indices.to(table.device()).to(torch::kInt64).contiguous();
    auto out = torch::empty({indices.size(0), table.size(1)}, table.options());
    mb_throw_on_error(mb_gather_rows(table.data_ptr<float>(), table.size(0), table.stride(0), table.size(1), indices.data_ptr<int64_t>(), indices.size(0),
                                     out.data_ptr<float>(), out.size(1), mb_current_stream(table.device())));
    return out;
}

static void device_rows_scatter(torch::Tensor& table, Indices indices, torch::Tensor values, bool add) {
    indices = indices.to(table.device()).to(torch::kInt64).contiguous();
    values = values.to(table.device()).to(torch::kFloat32).contiguous();
    auto fn = add ? mb_scatter_add_rows : mb_scatter_put_rows;
    mb_throw_on_error(fn(table.data_ptr<float>(), table.size(0), table.stride(0), table.size(1), indices.data_ptr<int64_t>(), indices.size(0),
                         values.data_ptr<float>(), values.stride(0), mb_current_stream(table.device())));
}

static void read_file_into(const string& filename, void* dst, int64_t bytes, int64_t offset = 0) {
    int fd = open(filename.c_str(), O_RDONLY);
    if (fd == -1) throw MariusRuntimeException("Unable to open " + filename);
    int64_t done = 0;
    while (done < bytes) {
        ssize_t r = pread(fd, (char*)dst + done, std::min<int64_t>(bytes - done, (int64_t)1 << 30), offset + done);
        if (r <= 0) {
            close(fd);
            throw MariusRuntimeException("short read on " + filename);
        }
        done += r;
    }
    close(fd);
}

static void write_file_from(const string& filename, const void* src, int64_t bytes, int64_t offset = 0, bool create = true) {
    int fd = open(filename.c_str(), create ? (O_RDWR | O_CREAT) : O_RDWR, 0666);
    if (fd == -1) throw MariusRuntimeException("Unable to open " + filename);
    int64_t done = 0;
    while (done < bytes) {
        ssize_t r = pwrite(fd, (const char*)src + done, std::min<int64_t>(bytes - done, (int64_t)1 << 30), offset + done);
        if (r <= 0) {
            close(fd);
            throw MariusRuntimeException("short write on " + filename);
        }
        done += r;
    }
    close(fd);
}

// ---- InMemory --------------------------------------------------------------------------------------------------
InMemory::InMemory(string filename, int64_t dim0_size, int64_t dim1_size, torch::Dtype dtype, torch::Device device) {
    filename_ = filename;
    dim0_size_ = dim0_size;
    dim1_size_ = dim1_size;
    dtype_ = dtype;
    initialized_ = true;
    device_ = device;
    if (dtype != torch::kFloat32) throw MariusRuntimeException("marius_b200 InMemory holds fp32 node embeddings / optimizer state only");
}

InMemory::InMemory(string filename, torch::Tensor data, torch::Device device) : InMemory(filename, data.size(0), data.size(1), torch::kFloat32, device) {
    // storage.cpp:504-517: persist the initial values, load() brings them to the device
    auto host = data.to(torch::kCPU).to(torch::kFloat32).contiguous();
    write_file_from(filename_, host.data_ptr<float>(), host.numel() * 4);
}

InMemory::InMemory(torch::Tensor data) {
    if (data.sizes().size() == 2) {
        dim0_size_ = data.size(0);
        dim1_size_ = data.size(1);
    } else if (data.sizes().size() == 1) {
        dim0_size_ = data.size(0);
        dim1_size_ = 1;
    } else {
        throw MariusRuntimeException("Tensor must have 1 or two dimensions");  // storage.cpp:527-535
    }
    if (data.scalar_type() != torch::kFloat32) throw MariusRuntimeException("marius_b200 InMemory holds fp32 data only");
    if (!data.device().is_cuda()) throw MariusRuntimeException("marius_b200 InMemory(tensor) needs a CUDA tensor (DEVICE_MEMORY backend)");
    filename_ = "";
    data_ = data.reshape({dim0_size_, dim1_size_});
    initialized_ = true;
    dtype_ = torch::kFloat32;
    device_ = data.device();
    loaded_ = true;
}

void InMemory::load() {
    if (loaded_ || filename_.empty()) return;
    // storage.cpp:547-573: read the flat file, then move it to the device
    auto host = torch::empty({dim0_size_, dim1_size_}, torch::TensorOptions().dtype(torch::kFloat32).pinned_memory(true));
    read_file_into(filename_, host.data_ptr<float>(), host.numel() * 4);
    data_ = host.to(device_);
    loaded_ = true;
}

void InMemory::write() {
    if (!loaded_ || filename_.empty()) return;  // storage.cpp:575-590
    auto host = data_.to(torch::kCPU).contiguous();
    write_file_from(filename_, host.data_ptr<float>(), host.numel() * 4);
}

void InMemory::unload(bool perform_write) {
    if (!loaded_) return;
    if (perform_write) write();
    if (!filename_.empty()) {
        data_ = torch::Tensor();
        loaded_ = false;
    }
}

torch::Tensor InMemory::indexRead(Indices indices) {
    check_indices(indices);
    if (!data_.defined()) return torch::Tensor();  // storage.cpp:646-648
    return device_rows_read(data_, indices);
}

void InMemory::indexAdd(Indices indices, torch::Tensor values) {
    check_values(data_, indices, values);
    device_rows_scatter(data_, indices, values, true);
}

void InMemory::indexPut(Indices indices, torch::Tensor values) {
    check_values(data_, indices, values);
    device_rows_scatter(data_, indices, values, false);
}

torch::Tensor InMemory::range(int64_t offset, int64_t n) {
    if (n + offset > dim0_size_) throw MariusRuntimeException("Invalid range");  // storage.cpp:698-703
    return data_.narrow(0, offset, n);
}

void InMemory::rangePut(int64_t offset, int64_t n, torch::Tensor values) { data_.narrow(0, offset, n).copy_(values); }

void InMemory::shuffle() { throw MariusRuntimeException("shuffle is an edge-storage operation (FlatFile / host InMemory): outside the embedding hot path"); }
void InMemory::sort(bool src) {
    (void)src;
    throw MariusRuntimeException("sort is an edge-storage operation (FlatFile / host InMemory): outside the embedding hot path");
}

void InMemory::adagradUpdate(InMemory& embeddings, InMemory& state, Indices indices, torch::Tensor gradients, float learning_rate) {
    check_values(embeddings.data_, indices, gradients);
    auto& t = embeddings.data_;
    auto& s = state.data_;
    if (s.sizes() != t.sizes() || s.stride(0) != t.stride(0)) throw std::runtime_error("");
    indices = indices.to(t.device()).contiguous();
    gradients = gradients.to(t.device()).contiguous();
    mb_throw_on_error(mb_adagrad_update_rows(t.data_ptr<float>(), s.data_ptr<float>(), t.size(0), t.stride(0), t.size(1), indices.data_ptr<int64_t>(),
                                             indices.size(0), gradients.data_ptr<float>(), gradients.stride(0), learning_rate,
                                             mb_current_stream(t.device())));
}

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
        ssize_t r = pread(fd_, (char*)host_addr + done, p->total_size_ - done, p->file_offset_ + done);
        if (r <= 0) throw MariusRuntimeException("short partition read");
        done += r;
    }
}
void PartitionedFile::writePartition(const void* host_addr, Partition* p) {
    if (host_addr == nullptr || p == nullptr) throw std::runtime_error("");
    int64_t done = 0;
    while (done < p->total_size_) {
        ssize_t r = pwrite(fd_, (const char*)host_addr + done, p->total_size_ - done, p->file_offset_ + done);
        if (r <= 0) throw MariusRuntimeException("short partition write");
        done += r;
    }
}

// ---- PartitionBuffer -------------------------------------------------------------------------------------------
PartitionBuffer::PartitionBuffer(int capacity, int num_partitions, int fine_to_coarse_ratio, int64_t partition_size, int embedding_size,
                                 int64_t total_embeddings, torch::Dtype dtype, string filename, bool prefetching, torch::Device device)
    : capacity_(capacity), num_partitions_(num_partitions), fine_to_coarse_ratio_(fine_to_coarse_ratio), embedding_size_(embedding_size),
      partition_size_(partition_size), total_embeddings_(total_embeddings), prefetching_(prefetching), device_(device), filename_(filename) {
    if (dtype != torch::kFloat32) throw MariusRuntimeException("PartitionBuffer holds fp32 embeddings (buffer.cpp:447,467-470 assume float)");
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

void PartitionBuffer::load() {
    if (loaded_) return;
    if (!buffer_state_.defined()) throw MariusRuntimeException("setBufferOrdering must be called before load");
    buffer_tensor_view_ = torch::zeros({capacity_ * partition_size_, (int64_t)embedding_size_}, torch::TensorOptions().dtype(torch::kFloat32).device(device_));
    staging_ = torch::empty({partition_size_, (int64_t)embedding_size_}, torch::TensorOptions().dtype(torch::kFloat32).pinned_memory(true));
    for (int i = 0; i < buffer_state_.size(0); i++) {
        Partition* p = partition_table_[buffer_state_[i].item<int>()];
        partitioned_file_->readPartition(staging_.data_ptr<float>(), p);
        buffer_tensor_view_.narrow(0, i * partition_size_, p->partition_size_).copy_(staging_.narrow(0, 0, p->partition_size_));
        p->present_ = true;
        p->buffer_idx_ = i;
    }
    loaded_ = true;
}

void PartitionBuffer::write() { sync(); }

void PartitionBuffer::unload(bool write) {
    if (!loaded_) return;
    if (write) sync();
    for (auto p : partition_table_) {
        p->present_ = false;
        p->buffer_idx_ = -1;
    }
    buffer_tensor_view_ = torch::Tensor();
    staging_ = torch::Tensor();
    loaded_ = false;
}

vector<int> PartitionBuffer::getNextAdmit() {
    vector<int> out;  // partitions of the next state that are not in the current one (buffer.cpp:543-561)
    if (state_pos_ >= buffer_states_.size()) return out;
    auto next = buffer_states_[state_pos_];
    for (int i = 0; i < next.size(0); i++) {
        int id = next[i].item<int>();
        bool found = false;
        for (int j = 0; j < buffer_state_.size(0); j++) found |= (buffer_state_[j].item<int>() == id);
        if (!found) out.push_back(id);
    }
    return out;
}

vector<int> PartitionBuffer::getNextEvict() {
    vector<int> out;  // buffer.cpp:563-579
    if (state_pos_ >= buffer_states_.size()) return out;
    auto next = buffer_states_[state_pos_];
    for (int i = 0; i < buffer_state_.size(0); i++) {
        int id = buffer_state_[i].item<int>();
        bool found = false;
        for (int j = 0; j < next.size(0); j++) found |= (next[j].item<int>() == id);
        if (!found) out.push_back(id);
    }
    return out;
}

Indices PartitionBuffer::getRandomIds(int64_t size) {
    int64_t n = 0;
    for (int i = 0; i < buffer_state_.size(0); i++) n += partition_table_[buffer_state_[i].item<int>()]->partition_size_;
    return torch::randint(n, {size}, torch::kInt64);  // buffer.cpp:457
}

torch::Tensor PartitionBuffer::indexRead(torch::Tensor indices) {
    check_indices(indices);
    return device_rows_read(buffer_tensor_view_, indices);
}

void PartitionBuffer::indexAdd(torch::Tensor indices, torch::Tensor values) {
    check_values(buffer_tensor_view_, indices, values);
    device_rows_scatter(buffer_tensor_view_, indices, values, true);
}

void PartitionBuffer::adagradUpdate(PartitionBuffer& state, torch::Tensor indices, torch::Tensor gradients, float learning_rate) {
    check_values(buffer_tensor_view_, indices, gradients);
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
        for (int i = 0; i < buffer_state_.size(0); i++) {
            Partition* p = partition_table_[buffer_state_[i].item<int>()];
            parts.push_back(p->partition_id_);
            slots.push_back(p->buffer_idx_);
        }
    } else {
        // mapping after the next swap: survivors keep their slot, admitted partitions take the evicted slots (buffer.cpp:603-630)
        if (state_pos_ >= buffer_states_.size()) throw MariusRuntimeException("no next buffer state");
        auto next = buffer_states_[state_pos_];
        auto evict = getNextEvict();
        auto admit = getNextAdmit();
        for (int i = 0; i < next.size(0); i++) {
            Partition* p = partition_table_[next[i].item<int>()];
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
    buffer_states_ = buffer_states;
    state_pos_ = 0;
    buffer_state_ = buffer_states_[state_pos_++];
    if (loaded_) {  // buffer.cpp:487-490
        unload(true);
        load();
    }
}

bool PartitionBuffer::hasSwap() { return state_pos_ < buffer_states_.size(); }

void PartitionBuffer::evict(vector<Partition*> evict_partitions) {
    for (auto p : evict_partitions) {  // HBM -> pinned host -> file
        staging_.narrow(0, 0, p->partition_size_).copy_(buffer_tensor_view_.narrow(0, p->buffer_idx_ * partition_size_, p->partition_size_));
        partitioned_file_->writePartition(staging_.data_ptr<float>(), p);
        p->present_ = false;
    }
}

void PartitionBuffer::admit(vector<Partition*> admit_partitions, vector<int64_t> buffer_idxs) {
    if (admit_partitions.size() > buffer_idxs.size()) throw std::runtime_error("");  // buffer.cpp:653-656
    for (size_t i = 0; i < admit_partitions.size(); i++) {
        Partition* p = admit_partitions[i];
        partitioned_file_->readPartition(staging_.data_ptr<float>(), p);
        buffer_tensor_view_.narrow(0, buffer_idxs[i] * partition_size_, p->partition_size_).copy_(staging_.narrow(0, 0, p->partition_size_));
        p->present_ = true;
        p->buffer_idx_ = (int)buffer_idxs[i];
    }
}

void PartitionBuffer::performNextSwap() {
    if (!buffer_state_.defined() || state_pos_ >= buffer_states_.size()) return;
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
    for (auto p : partition_table_) {  // buffer.cpp:685-696
        if (p->present_) {
            staging_.narrow(0, 0, p->partition_size_).copy_(buffer_tensor_view_.narrow(0, p->buffer_idx_ * partition_size_, p->partition_size_));
            partitioned_file_->writePartition(staging_.data_ptr<float>(), p);
        }
    }
}
