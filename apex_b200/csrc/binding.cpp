// Host runtime for the multi-tensor engine: builds and CACHES the device-resident tensor table.
// This is the only translation unit that sees torch headers; every kernel launcher is a torch-free C ABI symbol in
// libapex_b200_kernels.so, called from python via ctypes with raw pointers.
//
// Versus the reference launcher (csrc/multi_tensor_apply.cuh:32-103), which re-validates and re-packs every tensor list
// into <=4 KB by-value kernel arguments on every call and launches once per 24..110 tensors, this table is built once per
// parameter set; per step only the gradient pointer column is re-read (in C++, no python loop) and re-uploaded if it
// actually changed.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <cstring>
#include <vector>

namespace {

int dtype_code(at::ScalarType t) {
  switch (t) {
    case at::kFloat: return 0;
    case at::kHalf: return 1;
    case at::kBFloat16: return 2;
    case at::kDouble: return 3;
    case at::kByte: return 4;
    case at::kInt: return 5;
    case at::kLong: return 6;
    case at::kShort: return 7;
    case at::kFloat8_e4m3fn: return 8;
    case at::kFloat8_e5m2: return 9;
    default: return -1;
  }
}

class TensorTable {
 public:
  TensorTable() = default;

  // lists[d][t]; every list has the same length and, per t, the same numel.
  void build(const std::vector<std::vector<at::Tensor>>& lists, int64_t chunk) {
    TORCH_CHECK(!lists.empty(), "TensorTable: need at least one tensor list");
    TORCH_CHECK(chunk > 0 && chunk % 32 == 0, "TensorTable: chunk must be a positive multiple of 32");
    depth_ = (int)lists.size();
    n_ = (int)lists[0].size();
    chunk_ = (int)chunk;
    lists_ = lists;
    dtypes_.assign(depth_, -1);
    grad_of_.assign(depth_, -1);
    if (n_ == 0) { total_chunks_ = 0; dev_ = at::Tensor(); return; }
    device_ = lists[0][0].device();
    TORCH_CHECK(device_.is_cuda(), "TensorTable: tensors must live on a CUDA device");
    for (int d = 0; d < depth_; d++) {
      TORCH_CHECK((int)lists[d].size() == n_, "TensorTable: list ", d, " has ", lists[d].size(), " tensors, expected ", n_);
      dtypes_[d] = dtype_code(lists[d][0].scalar_type());
      for (int t = 0; t < n_; t++) check_tensor(lists[d][t], lists[0][t], d, t, lists[d][0].scalar_type());
    }
    host_.resize(arena_bytes());
    char* b = host_.data();
    void** ptrs = reinterpret_cast<void**>(b);
    int64_t* numel = reinterpret_cast<int64_t*>(b + sizeof(void*) * (size_t)depth_ * n_);
    int32_t* prefix = reinterpret_cast<int32_t*>(b + sizeof(void*) * (size_t)depth_ * n_ + sizeof(int64_t) * (size_t)n_);
    int64_t acc = 0;
    for (int t = 0; t < n_; t++) {
      const int64_t ne = lists[0][t].numel();
      numel[t] = ne;
      prefix[t] = (int32_t)acc;
      acc += (ne + chunk_ - 1) / chunk_;
      TORCH_CHECK(acc < (int64_t)INT32_MAX, "TensorTable: too many chunks");
    }
    prefix[n_] = (int32_t)acc;
    total_chunks_ = (int)acc;
    for (int d = 0; d < depth_; d++)
      for (int t = 0; t < n_; t++) ptrs[(size_t)d * n_ + t] = lists[d][t].data_ptr();
    dev_ = at::empty({(int64_t)host_.size()}, at::TensorOptions().dtype(at::kByte).device(device_));
    upload(0, host_.size());
    uploads_++;
  }

  // Declare that slot `grad_slot` holds the .grad of the tensors in `param_slot` (so refresh_grads() can follow them).
  void track_grads(int grad_slot, int param_slot) {
    TORCH_CHECK(grad_slot >= 0 && grad_slot < depth_ && param_slot >= 0 && param_slot < depth_);
    grad_of_[grad_slot] = param_slot;
  }

  // Re-read .grad pointers for every tracked slot. Returns false if the table can no longer describe the step
  // (a grad disappeared / changed dtype / changed size) and must be rebuilt by the caller.
  // Every parameter of the optimizer group + whether it is described by this table. refresh_grads() then also notices a
  // parameter that gained or lost its gradient since the table was built (frozen / unfrozen layers).
  void track_universe(const std::vector<at::Tensor>& all, const std::vector<bool>& in_table) {
    TORCH_CHECK(all.size() == in_table.size(), "TensorTable.track_universe: length mismatch");
    universe_ = all;
    in_table_ = in_table;
  }

  bool refresh_grads() {
    for (size_t i = 0; i < universe_.size(); i++)
      if (universe_[i].grad().defined() != in_table_[i]) return false;
    if (n_ == 0) return true;
    void** ptrs = reinterpret_cast<void**>(host_.data());
    // Non-gradient slots (parameters, moments, masters): ``p.data = ...``, ``model.half()`` / ``.to()`` or re-homing parameters into a
    // flat buffer swap the storage behind a cached tensor. A changed address is re-uploaded, a changed dtype / size forces a rebuild
    // (the reference rebuilds its lists every step, csrc/multi_tensor_apply.cuh:64-102, and cannot go stale).
    for (int d = 0; d < depth_; d++) {
      if (grad_of_[d] >= 0) continue;
      bool changed = false;
      for (int t = 0; t < n_; t++) {
        const at::Tensor& x = lists_[d][t];
        if (dtype_code(x.scalar_type()) != dtypes_[d] || x.numel() != lists_[0][t].numel() || !x.is_non_overlapping_and_dense() ||
            x.device() != device_)
          return false;
        void* p = x.data_ptr();
        if (ptrs[(size_t)d * n_ + t] != p) { ptrs[(size_t)d * n_ + t] = p; changed = true; }
      }
      if (changed) { upload(sizeof(void*) * (size_t)d * n_, sizeof(void*) * (size_t)n_); uploads_++; }
    }
    for (int d = 0; d < depth_; d++) {
      const int ps = grad_of_[d];
      if (ps < 0) continue;
      bool changed = false;
      for (int t = 0; t < n_; t++) {
        const at::Tensor& g = lists_[ps][t].grad();
        if (!g.defined() || dtype_code(g.scalar_type()) != dtypes_[d] || g.numel() != lists_[ps][t].numel() ||
            !g.is_non_overlapping_and_dense() || g.is_sparse())
          return false;
        void* p = g.data_ptr();
        if (ptrs[(size_t)d * n_ + t] != p) { ptrs[(size_t)d * n_ + t] = p; changed = true; }
        lists_[d][t] = g;  // keep the storage alive until the kernels that read it have been enqueued
      }
      if (changed) { upload(sizeof(void*) * (size_t)d * n_, sizeof(void*) * (size_t)n_); uploads_++; }
    }
    return true;
  }

  // Replace one slot's tensors (same count/numel) and re-upload that pointer column if it changed.
  void set_slot(int d, const std::vector<at::Tensor>& ts) {
    TORCH_CHECK(d >= 0 && d < depth_ && (int)ts.size() == n_, "TensorTable.set_slot: bad slot or length");
    if (n_ == 0) return;
    void** ptrs = reinterpret_cast<void**>(host_.data());
    bool changed = false;
    for (int t = 0; t < n_; t++) {
      check_tensor(ts[t], lists_[0][t], d, t, ts[0].scalar_type());
      void* p = ts[t].data_ptr();
      if (ptrs[(size_t)d * n_ + t] != p) { ptrs[(size_t)d * n_ + t] = p; changed = true; }
    }
    lists_[d] = ts;
    dtypes_[d] = dtype_code(ts[0].scalar_type());
    if (changed) { upload(sizeof(void*) * (size_t)d * n_, sizeof(void*) * (size_t)n_); uploads_++; }
  }

  int64_t arena() const { return dev_.defined() ? reinterpret_cast<int64_t>(dev_.data_ptr()) : 0; }
  int n() const { return n_; }
  int depth() const { return depth_; }
  int total_chunks() const { return total_chunks_; }
  int chunk() const { return chunk_; }
  int64_t uploads() const { return uploads_; }
  std::vector<int> dtypes() const { return dtypes_; }
  int64_t total_numel() const {
    int64_t s = 0;
    for (int t = 0; t < n_; t++) s += lists_[0][t].numel();
    return s;
  }
  std::vector<at::Tensor> slot(int d) const { return lists_.at(d); }

 private:
  size_t arena_bytes() const {
    return sizeof(void*) * (size_t)depth_ * n_ + sizeof(int64_t) * (size_t)n_ + sizeof(int32_t) * ((size_t)n_ + 1);
  }
  void check_tensor(const at::Tensor& x, const at::Tensor& ref, int d, int t, at::ScalarType st) const {
    TORCH_CHECK(x.defined(), "TensorTable: undefined tensor at list ", d, " index ", t);
    TORCH_CHECK(x.device() == device_, "TensorTable: tensor at list ", d, " index ", t, " is on ", x.device(),
                ", expected ", device_);
    TORCH_CHECK(x.scalar_type() == st, "TensorTable: mixed dtypes inside list ", d);
    TORCH_CHECK(x.numel() == ref.numel(), "TensorTable: size mismatch at list ", d, " index ", t);
    TORCH_CHECK(x.is_non_overlapping_and_dense(), "TensorTable: tensor at list ", d, " index ", t,
                " is not dense/contiguous");
  }
  void upload(size_t off, size_t bytes) {
    // pinned staging from torch's caching host allocator: it will not recycle the block until the copy has executed.
    at::Tensor stage = at::empty({(int64_t)bytes}, at::TensorOptions().dtype(at::kByte).pinned_memory(true));
    std::memcpy(stage.data_ptr(), host_.data() + off, bytes);
    dev_.narrow(0, (int64_t)off, (int64_t)bytes).copy_(stage, /*non_blocking=*/true);
  }

  int n_ = 0, depth_ = 0, chunk_ = 0, total_chunks_ = 0;
  int64_t uploads_ = 0;
  c10::Device device_{c10::kCPU};
  std::vector<std::vector<at::Tensor>> lists_;
  std::vector<int> dtypes_, grad_of_;
  std::vector<at::Tensor> universe_;
  std::vector<bool> in_table_;
  std::vector<char> host_;
  at::Tensor dev_;
};

// .grad of each tensor (undefined -> None) without a python loop
std::vector<c10::optional<at::Tensor>> grads_of(const std::vector<at::Tensor>& params) {
  std::vector<c10::optional<at::Tensor>> out;
  out.reserve(params.size());
  for (const auto& p : params) {
    const at::Tensor& g = p.grad();
    if (g.defined()) out.emplace_back(g); else out.emplace_back(c10::nullopt);
  }
  return out;
}

std::vector<int64_t> data_ptrs(const std::vector<at::Tensor>& ts) {
  std::vector<int64_t> out;
  out.reserve(ts.size());
  for (const auto& t : ts) out.push_back(reinterpret_cast<int64_t>(t.data_ptr()));
  return out;
}

int64_t current_stream_ptr(int64_t device_index) {
  return reinterpret_cast<int64_t>(c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device_index).stream());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "apex_b200 host runtime (tensor tables)";
  py::class_<TensorTable>(m, "TensorTable")
      .def(py::init<>())
      .def("build", &TensorTable::build, py::arg("lists"), py::arg("chunk"))
      .def("track_grads", &TensorTable::track_grads)
      .def("track_universe", &TensorTable::track_universe)
      .def("refresh_grads", &TensorTable::refresh_grads)
      .def("set_slot", &TensorTable::set_slot)
      .def("slot", &TensorTable::slot)
      .def_property_readonly("arena", &TensorTable::arena)
      .def_property_readonly("n", &TensorTable::n)
      .def_property_readonly("depth", &TensorTable::depth)
      .def_property_readonly("total_chunks", &TensorTable::total_chunks)
      .def_property_readonly("chunk", &TensorTable::chunk)
      .def_property_readonly("uploads", &TensorTable::uploads)
      .def_property_readonly("dtypes", &TensorTable::dtypes)
      .def_property_readonly("total_numel", &TensorTable::total_numel);
  m.def("grads_of", &grads_of);
  m.def("data_ptrs", &data_ptrs);
  m.def("current_stream_ptr", &current_stream_ptr);
}
