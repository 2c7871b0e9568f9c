// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal stand-in for the ten TensorFlow / Eigen headers that the reference
// translation units
//     /root/reference/cpp/PSROIPooling/ps_roi_align_op.cc       (:22-31)
//     /root/reference/cpp/PSROIPooling/ps_roi_align_grad_op.cc  (:22-32)
//     /root/reference/cpp/PSROIPooling/ps_roi_align_op.h        (:25-27)
//     /root/reference/cpp/PSROIPooling/work_sharder.h           (:21-22)
// include, so that those files compile UNMODIFIED, where they lie, with plain
// g++ and no TensorFlow install.  Only the surface those files touch exists.
// `Shard` is mapped to std::thread (the reference hands it to TF's intra-op
// pool, ps_roi_align_op.cc:195-199).
#ifndef XDET_ORACLE_TF_SHIM_H_
#define XDET_ORACLE_TF_SHIM_H_

#include <cstdint>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace Eigen {
struct ThreadPoolDevice {};
struct GpuDevice {};
}  // namespace Eigen

namespace tensorflow {

typedef long long int64;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  bool ok() const { return ok_; }
  static Status OK() { return Status(); }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline void AppendAll(std::ostringstream&) {}
template <typename A, typename... R>
inline void AppendAll(std::ostringstream& os, const A& a, const R&... r) {
  os << a;
  AppendAll(os, r...);
}
template <typename... Args>
inline Status InvalidArgument(const Args&... args) {
  std::ostringstream os;
  AppendAll(os, args...);
  return Status(os.str());
}
}  // namespace errors

class StringPiece {
 public:
  StringPiece(const char* s) : s_(s) {}
  StringPiece(const std::string& s) : s_(s) {}
  bool contains(const StringPiece& o) const { return s_.find(o.s_) != std::string::npos; }

 private:
  std::string s_;
};

template <typename T>
struct FlatView {
  T* p;
  int64 n;
  T* data() const { return p; }
  int64 size() const { return n; }
  FlatView& setZero() {
    std::memset((void*)p, 0, sizeof(T) * (size_t)n);
    return *this;
  }
};

template <typename T>
struct TTypes {
  typedef FlatView<T> Flat;
  typedef FlatView<const T> ConstFlat;
};

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : d_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : d_(d) {}
  int dims() const { return (int)d_.size(); }
  int64 dim_size(int i) const { return d_[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (auto v : d_) n *= v;
    return n;
  }
  bool operator==(const TensorShape& o) const { return d_ == o.d_; }

 private:
  std::vector<int64> d_;
};

// A tensor that either owns its storage or borrows a caller buffer.
class Tensor {
 public:
  Tensor() : ptr_(nullptr) {}
  Tensor(const TensorShape& s, void* borrowed) : shape_(s), ptr_(borrowed) {}
  const TensorShape& shape() const { return shape_; }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  template <typename T>
  typename TTypes<T>::ConstFlat flat() const {
    return typename TTypes<T>::ConstFlat{(const T*)ptr_, shape_.num_elements()};
  }
  template <typename T>
  typename TTypes<T>::Flat flat() {
    return typename TTypes<T>::Flat{(T*)ptr_, shape_.num_elements()};
  }

 private:
  TensorShape shape_;
  void* ptr_;
};

namespace thread {
class ThreadPool {};
}  // namespace thread

class DeviceBase {
 public:
  struct CpuWorkerThreads {
    int num_threads;
    thread::ThreadPool* workers;
  };
  explicit DeviceBase(int nthreads) : w_{nthreads, &pool_} {}
  const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &w_; }

 private:
  thread::ThreadPool pool_;
  CpuWorkerThreads w_;
};

class OpKernelConstruction {
 public:
  OpKernelConstruction(int gw, int gh, const std::string& method)
      : gw_(gw), gh_(gh), method_(method) {}
  Status GetAttr(const std::string& name, int32_t* v) const {
    if (name == "grid_dim_width") { *v = gw_; return Status::OK(); }
    if (name == "grid_dim_height") { *v = gh_; return Status::OK(); }
    return errors::InvalidArgument("no int attr ", name);
  }
  Status GetAttr(const std::string& name, std::string* v) const {
    if (name == "pool_method") { *v = method_; return Status::OK(); }
    return errors::InvalidArgument("no string attr ", name);
  }
  void SetStatus(const Status& s) { if (status_.ok()) status_ = s; }
  const Status& status() const { return status_; }

 private:
  int gw_, gh_;
  std::string method_;
  Status status_;
};

class OpKernelContext {
 public:
  OpKernelContext(int nthreads) : dev_(nthreads) {}
  void add_input(const TensorShape& s, const void* p) { in_.emplace_back(s, const_cast<void*>(p)); }
  void add_output_buffer(void* p) { out_buf_.push_back(p); }
  const Tensor& input(int i) const { return in_[i]; }
  DeviceBase* device() { return &dev_; }
  Status allocate_output(int i, const TensorShape& s, Tensor** t) {
    if ((size_t)i >= out_buf_.size()) return errors::InvalidArgument("no buffer for output ", i);
    if (out_.size() <= (size_t)i) out_.resize(i + 1);
    out_[i].reset(new Tensor(s, out_buf_[i]));
    *t = out_[i].get();
    return Status::OK();
  }
  template <typename D>
  const D& eigen_device() const {
    static D d;
    return d;
  }
  void SetStatus(const Status& s) { if (status_.ok()) status_ = s; }
  const Status& status() const { return status_; }

 private:
  DeviceBase dev_;
  std::vector<Tensor> in_;
  std::vector<void*> out_buf_;
  std::vector<std::unique_ptr<Tensor>> out_;
  Status status_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext*) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->SetStatus(STATUS);       \
      return;                         \
    }                                 \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)                      \
  do {                                                \
    ::tensorflow::Status _s(__VA_ARGS__);             \
    if (!_s.ok()) {                                   \
      (CTX)->SetStatus(_s);                           \
      return;                                         \
    }                                                 \
  } while (0)

#define TF_RETURN_IF_ERROR(...)                       \
  do {                                                \
    ::tensorflow::Status _s(__VA_ARGS__);             \
    if (!_s.ok()) return _s;                          \
  } while (0)

namespace shape_inference {
struct DimensionHandle {
  int64 v;
  DimensionHandle() : v(-1) {}
  DimensionHandle(int64 x) : v(x) {}
};
struct ShapeHandle {
  std::vector<int64> d;
};
class InferenceContext {
 public:
  ShapeHandle input(int) const { return ShapeHandle(); }
  DimensionHandle Dim(const ShapeHandle&, int) const { return DimensionHandle(); }
  Status GetAttr(const std::string&, int32_t* v) const { *v = 1; return Status::OK(); }
  Status Divide(DimensionHandle, int64, bool, DimensionHandle*) { return Status::OK(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle>) { return ShapeHandle(); }
  void set_output(int, const ShapeHandle&) {}
};
}  // namespace shape_inference

// REGISTER_OP("X").Attr(..)...SetShapeFn(lambda): a chainable no-op builder.
struct OpDefBuilderShim {
  OpDefBuilderShim& Attr(const char*) { return *this; }
  OpDefBuilderShim& Input(const char*) { return *this; }
  OpDefBuilderShim& Output(const char*) { return *this; }
  OpDefBuilderShim& Doc(const char*) { return *this; }
  template <typename F>
  OpDefBuilderShim& SetShapeFn(F) { return *this; }
};
#define XDET_SHIM_CAT_(a, b) a##b
#define XDET_SHIM_CAT(a, b) XDET_SHIM_CAT_(a, b)
#define REGISTER_OP(NAME) \
  static ::tensorflow::OpDefBuilderShim XDET_SHIM_CAT(xdet_shim_op_, __COUNTER__) = ::tensorflow::OpDefBuilderShim()
#define REGISTER_KERNEL_BUILDER(...)

// Threaded stand-in for TF's Shard (declared by the reference's work_sharder.h:47-48).
inline void ShardImpl(int max_parallelism, int64 total, const std::function<void(int64, int64)>& work) {
  int n = max_parallelism < 1 ? 1 : max_parallelism;
  if ((int64)n > total) n = total > 0 ? (int)total : 1;
  if (n == 1) { work(0, total); return; }
  std::vector<std::thread> th;
  int64 per = (total + n - 1) / n;
  for (int t = 0; t < n; ++t) {
    int64 s = per * t, e = s + per > total ? total : s + per;
    if (s >= e) break;
    th.emplace_back([&work, s, e] { work(s, e); });
  }
  for (auto& t : th) t.join();
}

}  // namespace tensorflow

#endif  // XDET_ORACLE_TF_SHIM_H_
