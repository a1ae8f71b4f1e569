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

void mbh_check_indices(const Indices& indices) {
    if (!indices.defined() || indices.sizes().size() != 1) throw std::runtime_error("");  // storage.cpp:607-610, buffer.cpp:442-445
}

void mbh_check_values(const torch::Tensor& table, const Indices& indices, const torch::Tensor& values) {
    // storage.cpp:652-655, buffer.cpp:461-464
    if (!values.defined() || !indices.defined() || indices.sizes().size() != 1 || values.sizes().size() != 2 || indices.size(0) != values.size(0) ||
        table.size(1) != values.size(1)) {
        throw std::runtime_error("");
    }
}

torch::Tensor mbh_device_rows_read(const torch::Tensor& table, Indices indices) {
    indices = indices.to(table.device()).to(torch::kInt64).contiguous();
    auto out = torch::empty({indices.size(0), table.size(1)}, table.options());
    mb_throw_on_error(mb_gather_rows(table.data_ptr<float>(), table.size(0), table.stride(0), table.size(1), indices.data_ptr<int64_t>(), indices.size(0),
                                     out.data_ptr<float>(), out.size(1), mb_current_stream(table.device())));
    return out;
}

void mbh_device_rows_scatter(torch::Tensor& table, Indices indices, torch::Tensor values, bool add) {
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
    mbh_check_indices(indices);
    if (!data_.defined()) return torch::Tensor();  // storage.cpp:646-648
    return mbh_device_rows_read(data_, indices);
}

void InMemory::indexAdd(Indices indices, torch::Tensor values) {
    mbh_check_values(data_, indices, values);
    mbh_device_rows_scatter(data_, indices, values, true);
}

void InMemory::indexPut(Indices indices, torch::Tensor values) {
    mbh_check_values(data_, indices, values);
    mbh_device_rows_scatter(data_, indices, values, false);
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
    mbh_check_values(embeddings.data_, indices, gradients);
    auto& t = embeddings.data_;
    auto& s = state.data_;
    if (s.sizes() != t.sizes() || s.stride(0) != t.stride(0)) throw std::runtime_error("");
    indices = indices.to(t.device()).contiguous();
    gradients = gradients.to(t.device()).contiguous();
    mb_throw_on_error(mb_adagrad_update_rows(t.data_ptr<float>(), s.data_ptr<float>(), t.size(0), t.stride(0), t.size(1), indices.data_ptr<int64_t>(),
                                             indices.size(0), gradients.data_ptr<float>(), gradients.stride(0), learning_rate,
                                             mb_current_stream(t.device())));
}
