// autofunc_b200.hpp -- the host side of the drop-in, in C++17, header-only, on top of the C ABI (autofunc_b200.h).
//
// The reference is compiled code (Go) and its toolchain is absent from the build image, so the tested host mirror
// of its interfaces is this header; go/b200/ is the same logic as a cgo shim.  Names, argument meaning and panic
// texts follow the reference (file:line citations are into unixpickle/autofunc):
//
//   linalg::Vector                      github.com/unixpickle/num-analysis/linalg Vector ([]float64)
//   Result / RResult                    result.go:10-30, 41-77
//   Variable / RVariable / NewRVariable variable.go:19-123
//   Gradient / RGradient / RVector      gradient.go:11,62,93   (maps keyed by *Variable)
//   Func / RFunc / Composed(R)Func      func.go:7-45
//   Batcher / RBatcher / Composed(R)Batcher / FuncBatcher / RFuncBatcher   batcher.go:5-111
//   LinTran, LinAdd                     lin_tran.go, lin_add.go
//   Sigmoid, Exp, Log, LogSigmoid, Sin, Cos, Norm, Softmax, SquaredNorm     math_funcs.go
//   Add, Sub, Mul, Div, Scale, AddScaler, Square, Inverse, Pow, SumAll, ScaleFirst, AddFirst, AddLogDomain,
//   SumAllLogDomain (+R forms); SquaredError (the north star's cost head)   arithmetic.go
//   Concat, Slice, Repeat, Split, ConcatBatched (+R forms)                  slices.go (ConcatBatched: Concat under a batcher)
//   Pool, PoolAll, PoolSplit (+R forms)                                     pool.go
//   Fold, FoldR                                                             fold.go
//   MatMulVec(s), MatMulVec(s)R                                             lin_alg.go:14-133
//   b200::DenseSigmoid, b200::FuseDense, b200::MLP, b200::LSTM             the fused / whole-network fast paths
//
// A Go `panic(msg)` is a thrown autofunc::Panic carrying the same message.  Go's nil-able `Gradient` argument of
// PropagateRGradient (result.go:62-65) is a `Gradient*`.  Results are std::shared_ptr-owned (Go: garbage collected).
//
// Everything a node computes stays in HBM: nodes hand device buffers to each other (b200::Dev), parameter gradients
// accumulate in device accumulators (b200::Session) and are added (+=) into the caller's host maps before the
// outermost Propagate* call returns, which keeps the temp-variable protocol of pool.go:114-136 working.
// There is no CPU arithmetic in this file and no fallback: without the CUDA library af_ctx_create fails and
// b200::Context::Default() throws.
#ifndef AUTOFUNC_B200_HPP_
#define AUTOFUNC_B200_HPP_

#include <cstdint>
#include <cstdio>
#include <array>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <iterator>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "autofunc_b200.h"

namespace autofunc {

namespace linalg {
using Vector = std::vector<double>;
}

// Go panic(msg).  `code` is the AF_E* status when the message came from the library.
class Panic : public std::runtime_error {
 public:
  int code;
  explicit Panic(const std::string& msg, int c = AF_ESHAPE) : std::runtime_error(msg), code(c) {}
};

inline std::string Sprintf(const char* fmt, long long a, long long b) {
  char buf[160];
  std::snprintf(buf, sizeof buf, fmt, a, b);
  return buf;
}

// =================================================================================================
// b200: device plumbing (go/b200/b200.go: context(), devVec, session)
// =================================================================================================
namespace b200 {

inline void check(int rc) {
  if (rc != AF_OK) throw Panic(af_last_error(), rc);  // AF_ESHAPE carries the reference's own panic text
}

// One device + one stream (SURVEY 8b threading).  Process-wide default on cuda:$LOCAL_RANK.
class Context {
 public:
  af_ctx* h = nullptr;

  static Context& Default() {
    static Context* c = new Context();  // never destroyed: device buffers may outlive static destruction order
    return *c;
  }
  // Size-bucketed free lists: operators allocate their outputs per call as the reference does with make()
  // (lin_tran.go:81); cudaMalloc / cudaFree per node would synchronise the stream.
  double* Alloc(int64_t n, int64_t* cap) {
    int64_t c = 32;
    if (n <= (1 << 20)) {
      while (c < n) c <<= 1;
    } else {
      c = (n + (1 << 18) - 1) / (1 << 18) * (1 << 18);
    }
    *cap = c;
    auto& fl = pool_[c];
    if (!fl.empty()) {
      double* p = fl.back();
      fl.pop_back();
      return p;
    }
    double* p = nullptr;
    check(af_malloc(h, c, &p));
    return p;
  }
  void Release(double* p, int64_t cap) { pool_[cap].push_back(p); }
  void Sync() { check(af_sync(h)); }
  int64_t LaunchCount() { return af_launch_count(h); }
  void Trim() {
    Sync();
    for (auto& kv : pool_)
      for (double* p : kv.second) af_free(h, p);
    pool_.clear();
  }

 private:
  Context() {
    const char* lr = std::getenv("LOCAL_RANK");
    check(af_ctx_create(lr ? std::atoi(lr) : 0, &h));
  }
  std::unordered_map<int64_t, std::vector<double*>> pool_;
};

inline af_ctx* ctx() { return Context::Default().h; }

// Data parallelism, one process per device (LOCAL_RANK picks the device above): rank 0 calls DPUniqueID, the 128 bytes
// travel to the other ranks by the host's own means, every rank calls DPInit.  The gradient sum itself is
// af_allreduce_sum / MLP::StepFreshDP / LSTM::StepDP -- Gradient.Add (gradient.go:28-30,108-120) across ranks.
namespace b200 {
using DPId = std::array<unsigned char, AF_DP_ID_BYTES>;
inline DPId DPUniqueID() {
  DPId id{};
  check(af_dp_unique_id(id.data()));
  return id;
}
inline void DPInit(int nranks, int rank, const DPId& id) { check(af_dp_init(ctx(), nranks, rank, id.data())); }
inline void DPDestroy() { check(af_dp_destroy(ctx())); }
}  // namespace b200

class DevVec;
using Dev = std::shared_ptr<DevVec>;

// A float64 vector resident in HBM.
class DevVec {
 public:
  double* p = nullptr;
  int64_t n = 0;

  DevVec(const DevVec&) = delete;
  DevVec& operator=(const DevVec&) = delete;
  ~DevVec() {
    if (cap_ > 0) Context::Default().Release(p, cap_);
  }
  static Dev Empty(int64_t n) {
    Dev v(new DevVec());
    v->n = n;
    v->p = Context::Default().Alloc(n, &v->cap_);
    return v;
  }
  static Dev Zeros(int64_t n) {
    Dev v = Empty(n);
    check(af_zero(ctx(), v->p, n));
    return v;
  }
  static Dev Upload(const double* src, int64_t n) {
    Dev v = Empty(n);
    if (n > 0) check(af_upload(ctx(), v->p, src, n));
    return v;
  }
  static Dev Upload(const linalg::Vector& x) { return Upload(x.data(), (int64_t)x.size()); }
  // A buffer owned by someone else (a plan's slab, a view into `owner`).
  static Dev Wrap(double* p, int64_t n, Dev owner = nullptr) {
    Dev v(new DevVec());
    v->p = p;
    v->n = n;
    v->owner_ = std::move(owner);
    return v;
  }
  static Dev View(const Dev& d, int64_t start, int64_t end) { return Wrap(d->p + start, end - start, d); }
  linalg::Vector Download() const {
    linalg::Vector out((size_t)n);
    if (n > 0) check(af_download(ctx(), out.data(), p, n));
    return out;
  }
  Dev Copy() const {
    Dev v = Empty(n);
    check(af_copy(ctx(), v->p, p, n));
    return v;
  }

 private:
  DevVec() = default;
  int64_t cap_ = 0;
  Dev owner_;
};

// A recorded launch sequence (af_graph_*): ABI-level calls issued between construction and End() are captured into a
// CUDA graph; Launch() replays them -- same buffers, same sizes -- as one submission.  Host-layer nodes allocate their
// outputs per call, so record explicit-buffer (af_*) sequences, after one eager warm-up run of the same sequence.
class Graph {
 public:
  Graph() { check(af_graph_begin(ctx())); }
  Graph(const Graph&) = delete;
  Graph& operator=(const Graph&) = delete;
  ~Graph() {
    if (!ended_) af_graph_end(ctx(), nullptr);
    af_graph_destroy(g_);
  }
  void End() {
    ended_ = true;
    check(af_graph_end(ctx(), &g_));
  }
  void Launch() { check(af_graph_launch(g_)); }
  int64_t Launches() const { return af_graph_launches(g_); }

 private:
  af_graph* g_ = nullptr;
  bool ended_ = false;
};

inline const double* ptr(const Dev& v) { return v ? v->p : nullptr; }  // NULL R operand == identically zero
inline double* mptr(const Dev& v) { return v ? v->p : nullptr; }
inline Dev ZerosIfNull(const Dev& v, int64_t n) { return v ? v : DevVec::Zeros(n); }

}  // namespace b200

// =================================================================================================
// Gradient maps (gradient.go)
// =================================================================================================
class Variable;
using VariablePtr = std::shared_ptr<Variable>;
using Gradient = std::unordered_map<Variable*, linalg::Vector>;  // gradient.go:11
using RGradient = Gradient;                                       // gradient.go:62
using RVector = Gradient;                                         // gradient.go:93

namespace b200 {
// Device accumulators for parameter gradients during one backward pass; Flush adds them (+=) into the maps.
class Session {
 public:
  Dev Accumulator(Gradient& m, Variable* v, int64_t len) {
    auto key = std::make_pair((const void*)&m, (const void*)v);
    auto it = acc_.find(key);
    if (it != acc_.end()) return it->second.acc;
    Entry e{&m, v, DevVec::Zeros(len), false};
    acc_.emplace(key, e);
    return e.acc;
  }
  void Accumulate(Gradient& m, Variable* v, const Dev& u) {
    Dev a = Accumulator(m, v, u->n);
    check(af_axpy(ctx(), 1.0, u->p, a->p, a->n));
  }
  // Make `acc` the accumulator for (m, v) and keep it out of Flush: the map's vector for v lives in HBM
  // (DeviceGradient below), so backward passes accumulate straight into it and nothing is copied to the host.
  void Seed(Gradient& m, Variable* v, Dev acc) {
    acc_[std::make_pair((const void*)&m, (const void*)v)] = Entry{&m, v, std::move(acc), true};
  }
  // The temp-variable protocol of pool.go:114-136 / lin_alg.go:57-68 / fold.go:62-84 inside a device backward pass.
  // The caller inserted `m[v]` as an EMPTY vector (a zero vector that has not been materialised: every host-side
  // `+=` in this header treats an empty destination as zeros of the source's length), ran the inner backward pass,
  // and now removes v from m and takes everything accumulated for it -- the device accumulator plus whatever a
  // foreign host node added to the host vector -- without leaving HBM.  `len` sizes the all-zero answer.
  Dev Take(Gradient& m, Variable* v, int64_t len) {
    Dev d;
    auto it = acc_.find(std::make_pair((const void*)&m, (const void*)v));
    if (it != acc_.end()) {
      d = it->second.acc;
      acc_.erase(it);
    }
    auto mit = m.find(v);
    if (mit == m.end()) throw Panic("temporary variable is missing from the gradient map", AF_EINVAL);
    linalg::Vector host = std::move(mit->second);
    m.erase(mit);
    if (!host.empty()) {
      if ((int64_t)host.size() != len) throw Panic("vector sizes do not match");
      Dev up = DevVec::Upload(host);
      if (d) check(af_axpy(ctx(), 1.0, up->p, d->p, d->n));
      else d = up;
    }
    return d ? d : DevVec::Zeros(len);
  }
  // Drop whatever is still keyed by map `m` (a node-local Gradient{} that stood in for a nil one, pool.go:164-166).
  void Forget(Gradient& m) {
    for (auto it = acc_.begin(); it != acc_.end();) it = it->second.map == &m ? acc_.erase(it) : std::next(it);
  }
  void Flush() {
    for (auto it0 = acc_.begin(); it0 != acc_.end();) {
      Entry& e = it0->second;
      if (e.resident) {
        ++it0;
        continue;
      }
      auto it = e.map->find(e.v);
      if (it != e.map->end()) {
        linalg::Vector h = e.acc->Download();
        linalg::Vector& dst = it->second;
        if (dst.empty()) dst.assign(h.size(), 0.0);  // an unmaterialised zero vector (see Take)
        if (dst.size() != h.size()) throw Panic("vector sizes do not match");
        for (size_t i = 0; i < h.size(); i++) dst[i] += h[i];
      }
      it0 = acc_.erase(it0);
    }
  }

 private:
  struct Entry {
    Gradient* map;
    Variable* v;
    Dev acc;
    bool resident = false;
  };
  std::map<std::pair<const void*, const void*>, Entry> acc_;
};
}  // namespace b200

// =================================================================================================
// Result / RResult (result.go)
// =================================================================================================
// The lower-case members are the device protocol of go/b200's `deviceResult`; their defaults make any foreign
// host-only Result usable as an input of a device node (upload on the way in, download + host call on the way out).
class Result {
 public:
  virtual ~Result() = default;
  virtual const linalg::Vector& Output() = 0;
  virtual bool Constant(const Gradient& g) = 0;
  // The callee may clobber `upstream` (result.go:24-29).
  virtual void PropagateGradient(linalg::Vector upstream, Gradient& grad) = 0;

  virtual int64_t len() { return (int64_t)Output().size(); }
  virtual b200::Dev dev() { return b200::DevVec::Upload(Output()); }
  virtual void bwd(b200::Session& s, b200::Dev u, Gradient& grad) {
    s.Flush();
    PropagateGradient(u->Download(), grad);
  }
};
using ResultPtr = std::shared_ptr<Result>;

class RResult {
 public:
  virtual ~RResult() = default;
  virtual const linalg::Vector& Output() = 0;
  virtual const linalg::Vector& ROutput() = 0;
  virtual bool Constant(const RGradient& rg, const Gradient* g) = 0;
  // grad may be nullptr (result.go:62-65).  Upstreams may be clobbered (result.go:71-76).
  virtual void PropagateRGradient(linalg::Vector upstream, linalg::Vector upstreamR, RGradient& rgrad,
                                  Gradient* grad) = 0;

  virtual int64_t len() { return (int64_t)Output().size(); }
  virtual b200::Dev dev() { return b200::DevVec::Upload(Output()); }
  virtual b200::Dev devR() { return b200::DevVec::Upload(ROutput()); }  // nullptr == identically zero
  virtual void bwdR(b200::Session& s, b200::Dev u, b200::Dev uR, RGradient& rgrad, Gradient* grad) {
    s.Flush();
    PropagateRGradient(u->Download(), uR->Download(), rgrad, grad);
  }
};
using RResultPtr = std::shared_ptr<RResult>;

// =================================================================================================
// Variable / RVariable (variable.go)
// =================================================================================================
class Variable : public Result {
 public:
  linalg::Vector Vector;  // variable.go:19-21

  explicit Variable(linalg::Vector v) : Vector(std::move(v)) {}

  const linalg::Vector& Output() override {  // variable.go:39
    Sync();
    return Vector;
  }
  bool Constant(const Gradient& g) override {                               // variable.go:49-52
    return g.find(this) == g.end();
  }
  void PropagateGradient(linalg::Vector upstream, Gradient& grad) override {  // variable.go:43-47
    auto it = grad.find(this);
    if (it == grad.end()) return;
    if (it->second.empty()) it->second.assign(upstream.size(), 0.0);  // an unmaterialised zero vector (Session::Take)
    if (it->second.size() != upstream.size()) throw Panic("vector sizes do not match");
    for (size_t i = 0; i < upstream.size(); i++) it->second[i] += upstream[i];
  }
  int64_t len() override { return host_stale_ ? dev_->n : (int64_t)Vector.size(); }

  // The device copy is cached.  Vectors of up to 64Ki entries are compared with the snapshot taken at upload;
  // after mutating a larger one in place call MarkDirty().
  b200::Dev dev() override {
    const size_t kAuto = 1 << 16;
    if (host_stale_) return dev_;
    if (dev_ && Vector.size() <= kAuto && snapshot_ != Vector) dev_.reset();
    if (dev_ && (size_t)dev_->n != Vector.size()) dev_.reset();
    if (!dev_) {
      dev_ = b200::DevVec::Upload(Vector);
      if (Vector.size() <= kAuto) snapshot_ = Vector;
    }
    return dev_;
  }
  void MarkDirty() { dev_.reset(); }
  // `&Variable{Vector: in.Output()}` (pool.go:42, fold.go:24, lin_alg.go:37) sharing the producing node's device
  // buffer: pooling a device node costs neither an upload nor a download.  `Vector` is filled on the first Output()
  // (or Sync()); until then len() and dev() are the device copy's.
  static std::shared_ptr<Variable> Wrap(b200::Dev d) {
    auto v = std::make_shared<Variable>(linalg::Vector());
    v->dev_ = std::move(d);
    v->host_stale_ = true;
    return v;
  }
  // Vector += scale * g without leaving HBM (gradient.go:40-52's Axpy on the device copy).  The device copy is then
  // the authoritative one: call Sync() before reading or writing `Vector` on the host again.
  void AddScaledDevice(double scale, const b200::Dev& g) {
    b200::Dev d = dev();
    if (g->n != d->n) throw Panic("vector sizes do not match");
    b200::check(af_axpy(b200::ctx(), scale, g->p, d->p, d->n));
    host_stale_ = true;
  }
  void Sync() {
    if (!host_stale_) return;
    Vector = dev_->Download();
    if (Vector.size() <= (size_t)(1 << 16)) snapshot_ = Vector;
    host_stale_ = false;
  }
  void bwd(b200::Session& s, b200::Dev u, Gradient& grad) override {
    if (grad.find(this) != grad.end()) s.Accumulate(grad, this, u);
  }

  // Wire format (variable.go:24-37,61-68): little-endian u64 count, then little-endian f64s.
  std::string Serialize() {
    Sync();  // a variable updated or produced on the device is written with its current contents
    std::string out(8 + 8 * Vector.size(), '\0');
    uint64_t n = Vector.size();
    for (int i = 0; i < 8; i++) out[i] = (char)((n >> (8 * i)) & 0xff);
    for (size_t k = 0; k < Vector.size(); k++) {
      uint64_t bits;
      std::memcpy(&bits, &Vector[k], 8);
      for (int i = 0; i < 8; i++) out[8 + 8 * k + i] = (char)((bits >> (8 * i)) & 0xff);
    }
    return out;
  }
  // Reads one variable from `bytes` starting at *pos and advances *pos.
  static VariablePtr Deserialize(const std::string& bytes, size_t* pos) {
    auto rd = [&](size_t at) {
      uint64_t v = 0;
      for (int i = 0; i < 8; i++) v |= (uint64_t)(unsigned char)bytes.at(at + i) << (8 * i);
      return v;
    };
    uint64_t n = rd(*pos);
    if (n > (bytes.size() - *pos - 8) / 8) throw Panic("serialized variable is truncated", AF_EINVAL);
    linalg::Vector v((size_t)n);
    for (uint64_t k = 0; k < n; k++) {
      uint64_t bits = rd(*pos + 8 + 8 * k);
      std::memcpy(&v[k], &bits, 8);
    }
    *pos += 8 + 8 * n;
    return std::make_shared<Variable>(std::move(v));
  }

 private:
  b200::Dev dev_;
  linalg::Vector snapshot_;
  bool host_stale_ = false;
};

inline VariablePtr NewVariable(linalg::Vector v) { return std::make_shared<Variable>(std::move(v)); }

class RVariable : public RResult {
 public:
  VariablePtr Var;  // variable.go:74 `Variable *Variable`

  // variable.go:79-92: a variable absent from the RVector gets a zero R-vector.
  RVariable(VariablePtr v, const RVector& rv) : Var(std::move(v)) {
    auto it = rv.find(Var.get());
    has_r_ = it != rv.end();
    if (has_r_) r_ = &it->second;
  }
  // &RVariable{Variable: v, ROutputVec: rOutput} (pool.go:58-61, fold.go:110-113, seqfunc/map_batcher.go:85-88);
  // rdev, when given, is the device copy of rOutput.
  RVariable(VariablePtr v, linalg::Vector rOutput, b200::Dev rdev = nullptr)
      : Var(std::move(v)), has_r_(true), own_r_(std::move(rOutput)), rdev_(std::move(rdev)) {
    r_ = &own_r_;
  }
  RVariable(const RVariable&) = delete;
  RVariable& operator=(const RVariable&) = delete;
  // &RVariable{Variable: pool, ROutputVec: in.ROutput()} with the R-output still in HBM (nullptr == identically zero)
  static std::shared_ptr<RVariable> Wrap(VariablePtr v, b200::Dev rdev) {
    auto r = std::make_shared<RVariable>(std::move(v), linalg::Vector(), rdev);
    r->has_r_ = rdev != nullptr;
    r->own_stale_ = rdev != nullptr;
    return r;
  }
  const linalg::Vector& Output() override { return Var->Output(); }
  const linalg::Vector& ROutput() override {
    if (own_stale_) {
      own_r_ = rdev_->Download();
      own_stale_ = false;
    }
    if (has_r_) return *r_;
    if ((int64_t)zero_.size() != Var->len()) zero_.assign((size_t)Var->len(), 0.0);
    return zero_;
  }
  bool Constant(const RGradient& rg, const Gradient* g) override {  // variable.go:102-111
    if (rg.find(Var.get()) != rg.end()) return false;
    if (g == nullptr) return true;
    return g->find(Var.get()) == g->end();
  }
  void PropagateRGradient(linalg::Vector upstream, linalg::Vector upstreamR, RGradient& rgrad,
                          Gradient* grad) override {  // variable.go:113-123
    auto add = [](linalg::Vector& dst, const linalg::Vector& src) {
      if (dst.empty()) dst.assign(src.size(), 0.0);  // an unmaterialised zero vector (Session::Take)
      if (dst.size() != src.size()) throw Panic("vector sizes do not match");
      for (size_t i = 0; i < src.size(); i++) dst[i] += src[i];
    };
    if (grad != nullptr) {
      auto it = grad->find(Var.get());
      if (it != grad->end()) add(it->second, upstream);
    }
    auto it = rgrad.find(Var.get());
    if (it != rgrad.end()) add(it->second, upstreamR);
  }
  int64_t len() override { return Var->len(); }
  b200::Dev dev() override { return Var->dev(); }
  b200::Dev devR() override {
    if (!has_r_) return nullptr;
    if (!rdev_) rdev_ = b200::DevVec::Upload(*r_);
    return rdev_;
  }
  void bwdR(b200::Session& s, b200::Dev u, b200::Dev uR, RGradient& rgrad, Gradient* grad) override {
    if (grad != nullptr && grad->find(Var.get()) != grad->end()) s.Accumulate(*grad, Var.get(), u);
    if (rgrad.find(Var.get()) != rgrad.end()) s.Accumulate(rgrad, Var.get(), uR);
  }

 private:
  bool has_r_ = false, own_stale_ = false;
  const linalg::Vector* r_ = nullptr;  // aliases the RVector's entry, as the reference's slice does
  linalg::Vector own_r_, zero_;
  b200::Dev rdev_;
};
using RVariablePtr = std::shared_ptr<RVariable>;

inline RVariablePtr NewRVariable(VariablePtr v, const RVector& rv) {
  return std::make_shared<RVariable>(std::move(v), rv);
}

// gradient.go:13-19
inline Gradient NewGradient(const std::vector<VariablePtr>& vars) {
  Gradient g;
  for (auto& v : vars) g[v.get()] = linalg::Vector((size_t)v->len(), 0.0);
  return g;
}
inline RGradient NewRGradient(const std::vector<VariablePtr>& vars) { return NewGradient(vars); }
// gradient.go:22-24 Zero, :28-30 Add, :33-36 Scale, :40-52 AddToVars, :54-57 Copy
inline void GradientZero(Gradient& g) {
  for (auto& kv : g) std::fill(kv.second.begin(), kv.second.end(), 0.0);
}
inline void GradientAdd(Gradient& g, const Gradient& g1) {
  for (auto& kv : g) {
    const linalg::Vector& o = g1.at(kv.first);
    for (size_t i = 0; i < o.size(); i++) kv.second[i] += o[i];
  }
}
inline void GradientScale(Gradient& g, double f) {
  for (auto& kv : g)
    for (double& x : kv.second) x *= f;
}
inline void GradientAddToVars(const Gradient& g, double scale) {
  for (auto& kv : g) {
    linalg::Vector& v = kv.first->Vector;
    for (size_t i = 0; i < v.size(); i++) v[i] += scale * kv.second[i];
    kv.first->MarkDirty();
  }
}

// =================================================================================================
// Func / RFunc / Batcher / RBatcher and their compositions (func.go, batcher.go)
// =================================================================================================
class Func {
 public:
  virtual ~Func() = default;
  virtual ResultPtr Apply(ResultPtr in) = 0;  // func.go:7-9
};
class RFunc : public virtual Func {
 public:
  virtual RResultPtr ApplyR(const RVector& v, RResultPtr in) = 0;  // func.go:24-27
};
class Batcher {
 public:
  virtual ~Batcher() = default;
  virtual ResultPtr Batch(ResultPtr in, int n) = 0;  // batcher.go:5-20
};
class RBatcher : public virtual Batcher {
 public:
  virtual RResultPtr BatchR(const RVector& v, RResultPtr in, int n) = 0;  // batcher.go:24-29
};

// func.go:12-45.  A ComposedRFunc is also a ComposedFunc, as in Go where RFunc embeds Func.
class ComposedRFunc : public RFunc, public std::vector<std::shared_ptr<RFunc>> {
 public:
  using std::vector<std::shared_ptr<RFunc>>::vector;
  ResultPtr Apply(ResultPtr in) override {
    for (auto& f : *this) in = f->Apply(in);
    return in;
  }
  RResultPtr ApplyR(const RVector& v, RResultPtr in) override {
    for (auto& f : *this) in = f->ApplyR(v, in);
    return in;
  }
};
class ComposedFunc : public Func, public std::vector<std::shared_ptr<Func>> {
 public:
  using std::vector<std::shared_ptr<Func>>::vector;
  ResultPtr Apply(ResultPtr in) override {
    for (auto& f : *this) in = f->Apply(in);
    return in;
  }
};
// batcher.go:86-111
class ComposedRBatcher : public RBatcher, public std::vector<std::shared_ptr<RBatcher>> {
 public:
  using std::vector<std::shared_ptr<RBatcher>>::vector;
  ResultPtr Batch(ResultPtr in, int n) override {
    for (auto& b : *this) in = b->Batch(in, n);
    return in;
  }
  RResultPtr BatchR(const RVector& v, RResultPtr in, int n) override {
    for (auto& b : *this) in = b->BatchR(v, in, n);
    return in;
  }
};
class ComposedBatcher : public Batcher, public std::vector<std::shared_ptr<Batcher>> {
 public:
  using std::vector<std::shared_ptr<Batcher>>::vector;
  ResultPtr Batch(ResultPtr in, int n) override {
    for (auto& b : *this) in = b->Batch(in, n);
    return in;
  }
};

// =================================================================================================
// device-backed Result bases
// =================================================================================================
namespace b200 {

class DeviceResult : public Result {
 public:
  explicit DeviceResult(Dev d) : d_(std::move(d)) {}
  const linalg::Vector& Output() override {
    if (!have_) {
      host_ = d_->Download();
      have_ = true;
    }
    return host_;
  }
  int64_t len() override { return d_->n; }
  Dev dev() override { return d_; }
  void PropagateGradient(linalg::Vector upstream, Gradient& grad) override {
    if ((int64_t)upstream.size() != d_->n) throw Panic("invalid upstream size");
    Session s;
    bwd(s, DevVec::Upload(upstream), grad);
    s.Flush();
  }
  void bwd(Session& s, Dev u, Gradient& grad) override = 0;

 protected:
  Dev d_;

 private:
  linalg::Vector host_;
  bool have_ = false;
};

class DeviceRResult : public RResult {
 public:
  DeviceRResult(Dev d, Dev dr) : d_(std::move(d)), dr_(std::move(dr)) {}
  const linalg::Vector& Output() override {
    if (!have_) {
      host_ = d_->Download();
      have_ = true;
    }
    return host_;
  }
  const linalg::Vector& ROutput() override {
    if (!have_r_) {
      host_r_ = dr_ ? dr_->Download() : linalg::Vector((size_t)d_->n, 0.0);
      have_r_ = true;
    }
    return host_r_;
  }
  int64_t len() override { return d_->n; }
  Dev dev() override { return d_; }
  Dev devR() override { return dr_; }
  void PropagateRGradient(linalg::Vector upstream, linalg::Vector upstreamR, RGradient& rgrad,
                          Gradient* grad) override {
    if ((int64_t)upstream.size() != d_->n || (int64_t)upstreamR.size() != d_->n)
      throw Panic("invalid upstream size");
    Session s;
    bwdR(s, DevVec::Upload(upstream), DevVec::Upload(upstreamR), rgrad, grad);
    s.Flush();
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override = 0;

 protected:
  Dev d_, dr_;

 private:
  linalg::Vector host_, host_r_;
  bool have_ = false, have_r_ = false;
};

// a Gradient / RGradient whose vectors live in HBM (gradient.go:11-57 on the device), so a
// training loop never downloads a gradient.  `Keys()` is the map handed to Propagate* (its host vectors stay empty:
// only the keys matter, as everywhere in the reference); the sums land in the device vectors.  Device chains only:
// a foreign host-only Result below a DeviceGradient pass would find the empty host vectors.
class DeviceGradient {
 public:
  explicit DeviceGradient(const std::vector<VariablePtr>& vars) {  // gradient.go:13-19
    for (auto& v : vars) {
      keys_[v.get()] = linalg::Vector();
      vecs_[v.get()] = DevVec::Zeros(v->len());
    }
  }
  Gradient& Keys() { return keys_; }
  Dev operator[](const VariablePtr& v) { return vecs_.at(v.get()); }
  linalg::Vector Host(const VariablePtr& v) { return vecs_.at(v.get())->Download(); }
  void Zero() {  // gradient.go:22-24
    for (auto& kv : vecs_) check(af_zero(ctx(), kv.second->p, kv.second->n));
  }
  void Scale(double f) {  // gradient.go:33-36
    for (auto& kv : vecs_) check(af_scale(ctx(), kv.second->p, f, kv.second->n));
  }
  void Add(DeviceGradient& o) {  // gradient.go:28-30
    for (auto& kv : vecs_) check(af_axpy(ctx(), 1.0, o.vecs_.at(kv.first)->p, kv.second->p, kv.second->n));
  }
  void AddToVars(double scale) {  // gradient.go:40-52
    for (auto& kv : vecs_) kv.first->AddScaledDevice(scale, kv.second);
  }
  void SeedInto(Session& s) {
    for (auto& kv : vecs_) s.Seed(keys_, kv.first, kv.second);
  }

 private:
  Gradient keys_;
  std::unordered_map<Variable*, Dev> vecs_;
};

// out.PropagateGradient(upstream, grad) / out.PropagateRGradient(upstream, upstreamR, rgrad, grad) with the maps in HBM.
inline void PropagateGradient(Result& out, const linalg::Vector& upstream, DeviceGradient& grad) {
  if ((int64_t)upstream.size() != out.len()) throw Panic("invalid upstream size");
  Session s;
  grad.SeedInto(s);
  out.bwd(s, DevVec::Upload(upstream), grad.Keys());
  s.Flush();
}
inline void PropagateRGradient(RResult& out, const linalg::Vector& upstream, const linalg::Vector& upstreamR,
                               DeviceGradient& rgrad, DeviceGradient* grad) {
  if ((int64_t)upstream.size() != out.len() || (int64_t)upstreamR.size() != out.len())
    throw Panic("invalid upstream size");
  Session s;
  rgrad.SeedInto(s);
  if (grad) grad->SeedInto(s);
  out.bwdR(s, DevVec::Upload(upstream), DevVec::Upload(upstreamR), rgrad.Keys(), grad ? &grad->Keys() : nullptr);
  s.Flush();
}

inline bool InGrad(const Gradient* g, Variable* v) { return g != nullptr && g->find(v) != g->end(); }
inline Dev AccOrNull(Session& s, Gradient* g, Variable* v) {
  return InGrad(g, v) ? s.Accumulator(*g, v, v->len()) : nullptr;
}

}  // namespace b200

// =================================================================================================
// LinTran (lin_tran.go:17-425)
// =================================================================================================
class LinTran : public RFunc, public RBatcher {
 public:
  VariablePtr Data;
  int Rows, Cols;

  LinTran(VariablePtr data, int rows, int cols) : Data(std::move(data)), Rows(rows), Cols(cols) {}

  ResultPtr Apply(ResultPtr in) override {  // lin_tran.go:24-30
    checkLen(in->len(), 1);
    return Batch(in, 1);
  }
  RResultPtr ApplyR(const RVector& v, RResultPtr in) override {  // lin_tran.go:33-39
    checkLen(in->len(), 1);
    return BatchR(v, in, 1);
  }
  ResultPtr Batch(ResultPtr in, int n) override;                       // lin_tran.go:50-62
  RResultPtr BatchR(const RVector& v, RResultPtr in, int n) override;  // lin_tran.go:64-77
  // BatchR with the matrix's RVariable already built (MatMulVecsR: RVector{v: mat.ROutput()} with the R-vector in HBM)
  RResultPtr BatchRWith(RVariablePtr rdata, RResultPtr in, int n);

 private:
  void checkLen(int64_t got, int n) const {
    if ((int64_t)Cols * n != got)  // lin_tran.go:26-27,52-54
      throw Panic(Sprintf("input length should be %lld but got %lld", (long long)Cols * n, got));
  }
};

namespace b200 {

class LinTranResult : public DeviceResult {
 public:
  LinTranResult(LinTran m, ResultPtr in, Dev x, int n, Dev y)
      : DeviceResult(std::move(y)), m_(std::move(m)), in_(std::move(in)), x_(std::move(x)), n_(n) {}
  bool Constant(const Gradient& g) override { return m_.Data->Constant(g) && in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // lin_tran.go:373-382
    if (!m_.Data->Constant(grad)) {
      Dev acc = s.Accumulator(grad, m_.Data.get(), (int64_t)m_.Rows * m_.Cols);
      check(af_lintran_wgrad(ctx(), u->p, x_->p, acc->p, m_.Rows, m_.Cols, n_));
    }
    if (!in_->Constant(grad)) {
      Dev d = DevVec::Empty((int64_t)n_ * m_.Cols);
      check(af_lintran_igrad(ctx(), u->p, m_.Data->dev()->p, d->p, m_.Rows, m_.Cols, n_));
      in_->bwd(s, d, grad);
    }
  }

 private:
  LinTran m_;
  ResultPtr in_;
  Dev x_;
  int n_;
};

class LinTranRResult : public DeviceRResult {
 public:
  LinTranRResult(LinTran m, RResultPtr in, RVariablePtr rdata, Dev x, Dev xr, int n, Dev y, Dev ry)
      : DeviceRResult(std::move(y), std::move(ry)), m_(std::move(m)), in_(std::move(in)), rdata_(std::move(rdata)),
        x_(std::move(x)), xr_(std::move(xr)), n_(n) {}
  bool Constant(const RGradient& rg, const Gradient* g) override {
    return in_->Constant(rg, g) && rdata_->Constant(rg, g);
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // lin_tran.go:408-425
    Variable* w = m_.Data.get();
    Dev gacc = AccOrNull(s, grad, w), racc = AccOrNull(s, &rgrad, w);
    const bool down = !in_->Constant(rgrad, grad);
    if (!down) {
      if (gacc || racc)
        check(af_lintran_wgrad_r(ctx(), u->p, uR->p, x_->p, ptr(xr_), mptr(gacc), mptr(racc), m_.Rows, m_.Cols, n_));
    } else {
      // both gradients from one call: with 9..16 samples the library streams W once for the two of them
      Dev d = DevVec::Empty((int64_t)n_ * m_.Cols), dr = DevVec::Empty((int64_t)n_ * m_.Cols);
      check(af_lintran_backward_r(ctx(), u->p, uR->p, rdata_->dev()->p, ptr(rdata_->devR()), x_->p, ptr(xr_), mptr(gacc),
                                  mptr(racc), d->p, dr->p, m_.Rows, m_.Cols, n_, 0));
      in_->bwdR(s, d, dr, rgrad, grad);
    }
  }

 private:
  LinTran m_;
  RResultPtr in_;
  RVariablePtr rdata_;
  Dev x_, xr_;
  int n_;
};

}  // namespace b200

inline ResultPtr LinTran::Batch(ResultPtr in, int n) {
  checkLen(in->len(), n);
  b200::Dev x = in->dev();
  b200::Dev y = b200::DevVec::Empty((int64_t)n * Rows);
  b200::check(af_lintran_fwd(b200::ctx(), Data->dev()->p, x->p, y->p, Rows, Cols, n));
  return std::make_shared<b200::LinTranResult>(*this, in, x, n, y);
}

inline RResultPtr LinTran::BatchR(const RVector& v, RResultPtr in, int n) {
  return BatchRWith(NewRVariable(Data, v), std::move(in), n);
}

inline RResultPtr LinTran::BatchRWith(RVariablePtr rdata, RResultPtr in, int n) {
  checkLen(in->len(), n);
  b200::Dev x = in->dev(), xr = in->devR();
  b200::Dev y = b200::DevVec::Empty((int64_t)n * Rows), ry = b200::DevVec::Empty((int64_t)n * Rows);
  b200::check(af_lintran_fwd_r(b200::ctx(), rdata->dev()->p, b200::ptr(rdata->devR()), x->p, b200::ptr(xr), y->p,
                               ry->p, Rows, Cols, n));
  return std::make_shared<b200::LinTranRResult>(*this, in, rdata, x, xr, n, y, ry);
}

// =================================================================================================
// LinAdd (lin_add.go:13-109); as a Batcher the packed form is Add(x, Repeat(b, n))
// =================================================================================================
class LinAdd : public RFunc, public RBatcher {
 public:
  VariablePtr Var;
  explicit LinAdd(VariablePtr v) : Var(std::move(v)) {}
  ResultPtr Apply(ResultPtr in) override { return Batch(in, 1); }
  RResultPtr ApplyR(const RVector& v, RResultPtr in) override { return BatchR(v, in, 1); }
  ResultPtr Batch(ResultPtr in, int n) override;
  RResultPtr BatchR(const RVector& v, RResultPtr in, int n) override;

 private:
  void checkLen(int64_t got, int n) const {
    int64_t want = Var->len() * n;
    if (want != got) throw Panic(Sprintf("input length should be %lld but got %lld", want, got));
  }
};

namespace b200 {

class LinAddResult : public DeviceResult {
 public:
  LinAddResult(VariablePtr var, ResultPtr in, int n, Dev y)
      : DeviceResult(std::move(y)), var_(std::move(var)), in_(std::move(in)), n_(n) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g) && var_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // lin_add.go:66-73
    if (!var_->Constant(grad)) {
      Dev acc = s.Accumulator(grad, var_.get(), var_->len());
      check(af_bias_grad(ctx(), u->p, acc->p, acc->n, n_));
    }
    if (!in_->Constant(grad)) in_->bwd(s, u, grad);
  }

 private:
  VariablePtr var_;
  ResultPtr in_;
  int n_;
};

class LinAddRResult : public DeviceRResult {
 public:
  LinAddRResult(RVariablePtr rvar, RResultPtr in, int n, Dev y, Dev ry)
      : DeviceRResult(std::move(y), std::move(ry)), rvar_(std::move(rvar)), in_(std::move(in)), n_(n) {}
  bool Constant(const RGradient& rg, const Gradient* g) override {
    return rvar_->Constant(rg, g) && in_->Constant(rg, g);
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // lin_add.go:94-109
    Variable* v = rvar_->Var.get();
    Dev gacc = AccOrNull(s, grad, v), racc = AccOrNull(s, &rgrad, v);
    if (gacc || racc)
      check(af_bias_grad_r(ctx(), u->p, uR->p, mptr(gacc), mptr(racc), v->len(), n_));
    if (!in_->Constant(rgrad, grad)) in_->bwdR(s, u, uR, rgrad, grad);
  }

 private:
  RVariablePtr rvar_;
  RResultPtr in_;
  int n_;
};

}  // namespace b200

inline ResultPtr LinAdd::Batch(ResultPtr in, int n) {
  checkLen(in->len(), n);
  b200::Dev x = in->dev();
  b200::Dev y = b200::DevVec::Empty(x->n);
  b200::check(af_bias_fwd(b200::ctx(), x->p, Var->dev()->p, y->p, Var->len(), n));
  return std::make_shared<b200::LinAddResult>(Var, in, n, y);
}

inline RResultPtr LinAdd::BatchR(const RVector& v, RResultPtr in, int n) {
  checkLen(in->len(), n);
  RVariablePtr rvar = NewRVariable(Var, v);
  b200::Dev x = in->dev(), xr = in->devR();
  b200::Dev y = b200::DevVec::Empty(x->n), ry = b200::DevVec::Empty(x->n);
  b200::check(af_bias_fwd_r(b200::ctx(), x->p, b200::ptr(xr), rvar->dev()->p, b200::ptr(rvar->devR()), y->p, ry->p,
                            Var->len(), n));
  return std::make_shared<b200::LinAddRResult>(rvar, in, n, y, ry);
}

// =================================================================================================
// Elementwise RFuncs (math_funcs.go); each is also an RBatcher (batching an elementwise function is applying it)
// =================================================================================================
namespace b200 {

// The kernels behind one elementwise function.  `saved` is what the backward pass reads besides the upstream:
// the forward OUTPUT (and R-output) for Sigmoid / Exp / Tanh, the forward INPUT (and R-input) for the others.
struct UnaryKernels {
  int (*fwd)(af_ctx*, const double* x, const double* xr, double* o, double* ro, int64_t len);
  int (*bwd)(af_ctx*, const double* saved, const double* savedR, double* u, double* uR, int64_t len);
  bool saves_output;
};

class UnaryResult : public DeviceResult {
 public:
  UnaryResult(const UnaryKernels* k, ResultPtr in, Dev saved, Dev out)
      : DeviceResult(std::move(out)), k_(k), in_(std::move(in)), saved_(std::move(saved)) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {
    if (in_->Constant(grad)) return;
    check(k_->bwd(ctx(), saved_->p, nullptr, u->p, nullptr, u->n));  // in place on the upstream
    in_->bwd(s, u, grad);
  }

 private:
  const UnaryKernels* k_;
  ResultPtr in_;
  Dev saved_;
};

class UnaryRResult : public DeviceRResult {
 public:
  UnaryRResult(const UnaryKernels* k, RResultPtr in, Dev saved, Dev savedR, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), k_(k), in_(std::move(in)), saved_(std::move(saved)),
        savedR_(std::move(savedR)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {
    if (in_->Constant(rgrad, grad)) return;
    check(k_->bwd(ctx(), saved_->p, ptr(savedR_), u->p, uR->p, u->n));
    in_->bwdR(s, u, uR, rgrad, grad);
  }

 private:
  const UnaryKernels* k_;
  RResultPtr in_;
  Dev saved_, savedR_;
};

class Elementwise : public RFunc, public RBatcher {
 public:
  explicit Elementwise(const UnaryKernels* k) : k_(k) {}
  ResultPtr Apply(ResultPtr in) override {
    Dev x = in->dev();
    Dev o = DevVec::Empty(x->n);
    check(k_->fwd(ctx(), x->p, nullptr, o->p, nullptr, x->n));
    return std::make_shared<UnaryResult>(k_, in, k_->saves_output ? o : x, o);
  }
  RResultPtr ApplyR(const RVector&, RResultPtr in) override {
    Dev x = in->dev(), xr = in->devR();
    Dev o = DevVec::Empty(x->n), ro = DevVec::Empty(x->n);
    check(k_->fwd(ctx(), x->p, ptr(xr), o->p, ro->p, x->n));
    return std::make_shared<UnaryRResult>(k_, in, k_->saves_output ? o : x, k_->saves_output ? ro : xr, o, ro);
  }
  ResultPtr Batch(ResultPtr in, int n) override {
    if (n <= 0 || in->len() % n != 0) throw Panic("input count does not divide input length");  // batcher.go:41
    return Apply(in);
  }
  RResultPtr BatchR(const RVector& v, RResultPtr in, int n) override {
    if (n <= 0 || in->len() % n != 0) throw Panic("input count does not divide input length");
    return ApplyR(v, in);
  }

 private:
  const UnaryKernels* k_;
};

// Sigmoid and Exp have dedicated non-R entry points (the hot ones); the rest take NULL R pointers.
inline int sigmoid_fwd(af_ctx* c, const double* x, const double* xr, double* o, double* ro, int64_t n) {
  return ro ? af_sigmoid_fwd_r(c, x, xr, o, ro, n) : af_sigmoid_fwd(c, x, o, n);
}
inline int sigmoid_bwd(af_ctx* c, const double* s, const double* sr, double* u, double* ur, int64_t n) {
  return ur ? af_sigmoid_bwd_r(c, s, sr, u, ur, n) : af_sigmoid_bwd(c, s, u, n);
}
inline int exp_fwd(af_ctx* c, const double* x, const double* xr, double* o, double* ro, int64_t n) {
  return ro ? af_exp_fwd_r(c, x, xr, o, ro, n) : af_exp_fwd(c, x, o, n);
}
inline int exp_bwd(af_ctx* c, const double* e, const double* er, double* u, double* ur, int64_t n) {
  return ur ? af_exp_bwd_r(c, e, er, u, ur, n) : af_exp_bwd(c, e, u, n);
}
inline const UnaryKernels* SigmoidKernels() {
  static const UnaryKernels k{sigmoid_fwd, sigmoid_bwd, true};
  return &k;
}
inline const UnaryKernels* ExpKernels() {
  static const UnaryKernels k{exp_fwd, exp_bwd, true};
  return &k;
}
inline const UnaryKernels* TanhKernels() {
  static const UnaryKernels k{af_tanh_fwd_r, af_tanh_bwd_r, true};
  return &k;
}
inline const UnaryKernels* LogKernels() {
  static const UnaryKernels k{af_log_fwd_r, af_log_bwd_r, false};
  return &k;
}
inline const UnaryKernels* LogSigmoidKernels() {
  static const UnaryKernels k{af_logsigmoid_fwd_r, af_logsigmoid_bwd_r, false};
  return &k;
}
inline const UnaryKernels* SinKernels() {
  static const UnaryKernels k{af_sin_fwd_r, af_sin_bwd_r, false};
  return &k;
}
inline const UnaryKernels* SquareKernels() {
  static const UnaryKernels k{af_square_fwd_r, af_square_bwd_r, false};
  return &k;
}

}  // namespace b200

struct Sigmoid : b200::Elementwise {  // math_funcs.go:286-374
  Sigmoid() : Elementwise(b200::SigmoidKernels()) {}
};
struct Exp : b200::Elementwise {  // math_funcs.go:12-97
  Exp() : Elementwise(b200::ExpKernels()) {}
};
struct Log : b200::Elementwise {  // math_funcs.go:99-189
  Log() : Elementwise(b200::LogKernels()) {}
};
struct LogSigmoid : b200::Elementwise {  // math_funcs.go:376-477
  LogSigmoid() : Elementwise(b200::LogSigmoidKernels()) {}
};
struct Sin : b200::Elementwise {  // math_funcs.go:511-597
  Sin() : Elementwise(b200::SinKernels()) {}
};
struct Tanh : b200::Elementwise {  // extension named by the north star; Sigmoid's contract
  Tanh() : Elementwise(b200::TanhKernels()) {}
};

// arithmetic.go:696-770
inline ResultPtr Square(ResultPtr in) { return b200::Elementwise(b200::SquareKernels()).Apply(std::move(in)); }
inline RResultPtr SquareR(RResultPtr in) {
  static const RVector none;
  return b200::Elementwise(b200::SquareKernels()).ApplyR(none, std::move(in));
}

// =================================================================================================
// arithmetic.go: Add, Sub, Mul, Scale, AddScaler, SumAll
// =================================================================================================
namespace b200 {

inline void SameLen(int64_t a, int64_t b) {
  if (a != b) throw Panic("vector sizes do not match");  // arithmetic.go:265
}

enum class BinOp { kAdd, kSub, kMul };

class BinaryResult : public DeviceResult {
 public:
  BinaryResult(BinOp op, ResultPtr a, ResultPtr b, Dev out)
      : DeviceResult(std::move(out)), op_(op), a_(std::move(a)), b_(std::move(b)) {}
  bool Constant(const Gradient& g) override { return a_->Constant(g) && b_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {
    bool ac = a_->Constant(grad), bc = b_->Constant(grad);
    switch (op_) {
      case BinOp::kAdd:  // arithmetic.go:34-49
        if (!ac) a_->bwd(s, bc ? u : u->Copy(), grad);
        if (!bc) b_->bwd(s, u, grad);
        break;
      case BinOp::kSub:  // arithmetic.go:128-135
        if (!ac) a_->bwd(s, bc ? u : u->Copy(), grad);
        if (!bc) {
          check(af_scale(ctx(), u->p, -1.0, u->n));
          b_->bwd(s, u, grad);
        }
        break;
      case BinOp::kMul: {  // arithmetic.go:288-306
        ResultPtr me[2] = {a_, b_}, other[2] = {b_, a_};
        for (int i = 0; i < 2; i++) {
          if (me[i]->Constant(grad)) continue;
          Dev d = DevVec::Empty(u->n);
          check(af_mul_bwd(ctx(), u->p, nullptr, other[i]->dev()->p, nullptr, d->p, nullptr, u->n));
          me[i]->bwd(s, d, grad);
        }
        break;
      }
    }
  }

 private:
  BinOp op_;
  ResultPtr a_, b_;
};

class BinaryRResult : public DeviceRResult {
 public:
  BinaryRResult(BinOp op, RResultPtr a, RResultPtr b, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), op_(op), a_(std::move(a)), b_(std::move(b)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return a_->Constant(rg, g) && b_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {
    bool ac = a_->Constant(rgrad, grad), bc = b_->Constant(rgrad, grad);
    switch (op_) {
      case BinOp::kAdd:  // arithmetic.go:79-96
        if (!ac) a_->bwdR(s, bc ? u : u->Copy(), bc ? uR : uR->Copy(), rgrad, grad);
        if (!bc) b_->bwdR(s, u, uR, rgrad, grad);
        break;
      case BinOp::kSub:  // arithmetic.go:175-182
        if (!ac) a_->bwdR(s, bc ? u : u->Copy(), bc ? uR : uR->Copy(), rgrad, grad);
        if (!bc) {
          check(af_scale(ctx(), u->p, -1.0, u->n));
          check(af_scale(ctx(), uR->p, -1.0, uR->n));
          b_->bwdR(s, u, uR, rgrad, grad);
        }
        break;
      case BinOp::kMul: {  // arithmetic.go:349-376
        RResultPtr me[2] = {a_, b_}, other[2] = {b_, a_};
        for (int i = 0; i < 2; i++) {
          if (me[i]->Constant(rgrad, grad)) continue;
          Dev d = DevVec::Empty(u->n), dr = DevVec::Empty(u->n);
          check(af_mul_bwd(ctx(), u->p, uR->p, other[i]->dev()->p, ptr(other[i]->devR()), d->p, dr->p, u->n));
          me[i]->bwdR(s, d, dr, rgrad, grad);
        }
        break;
      }
    }
  }

 private:
  BinOp op_;
  RResultPtr a_, b_;
};

inline Dev BinaryForward(BinOp op, const Dev& a, const Dev& b) {
  Dev out = DevVec::Empty(a->n);
  auto fn = op == BinOp::kAdd ? af_add : op == BinOp::kSub ? af_sub : af_mul;
  check(fn(ctx(), a->p, b->p, out->p, a->n));
  return out;
}

inline ResultPtr Binary(BinOp op, ResultPtr a, ResultPtr b) {
  SameLen(a->len(), b->len());
  Dev out = BinaryForward(op, a->dev(), b->dev());
  return std::make_shared<BinaryResult>(op, std::move(a), std::move(b), out);
}

inline RResultPtr BinaryR(BinOp op, RResultPtr a, RResultPtr b) {
  SameLen(a->len(), b->len());
  Dev x = a->dev(), y = b->dev();
  Dev out = BinaryForward(op, x, y), outR;
  if (op == BinOp::kMul) {  // arithmetic.go:325-329: a*bR + aR*b
    outR = DevVec::Empty(x->n);
    check(af_mul_r(ctx(), x->p, ptr(a->devR()), y->p, ptr(b->devR()), outR->p, x->n));
  } else {
    outR = BinaryForward(op, ZerosIfNull(a->devR(), x->n), ZerosIfNull(b->devR(), x->n));
  }
  return std::make_shared<BinaryRResult>(op, std::move(a), std::move(b), out, outR);
}

class ScaledResult : public DeviceResult {
 public:
  ScaledResult(ResultPtr in, double f, Dev out) : DeviceResult(std::move(out)), in_(std::move(in)), f_(f) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // arithmetic.go:444-456
    if (in_->Constant(grad)) return;
    if (f_ != 1.0) check(af_scale(ctx(), u->p, f_, u->n));
    in_->bwd(s, u, grad);
  }

 private:
  ResultPtr in_;
  double f_;
};

class ScaledRResult : public DeviceRResult {
 public:
  ScaledRResult(RResultPtr in, double f, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)), f_(f) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // arithmetic.go:476-493
    if (in_->Constant(rgrad, grad)) return;
    if (f_ != 1.0) {
      check(af_scale(ctx(), u->p, f_, u->n));
      check(af_scale(ctx(), uR->p, f_, uR->n));
    }
    in_->bwdR(s, u, uR, rgrad, grad);
  }

 private:
  RResultPtr in_;
  double f_;
};

class SumAllResult : public DeviceResult {
 public:
  SumAllResult(ResultPtr in, Dev out) : DeviceResult(std::move(out)), in_(std::move(in)) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // arithmetic.go:987-995
    if (in_->Constant(grad)) return;
    Dev d = DevVec::Empty(in_->len());
    check(af_fill_dev(ctx(), d->p, u->p, d->n));  // the upstream scalar never leaves the device
    in_->bwd(s, d, grad);
  }

 private:
  ResultPtr in_;
};

class SumAllRResult : public DeviceRResult {
 public:
  SumAllRResult(RResultPtr in, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // arithmetic.go:1035-1047
    if (in_->Constant(rgrad, grad)) return;
    Dev d = DevVec::Empty(in_->len()), dr = DevVec::Empty(in_->len());
    check(af_fill_dev(ctx(), d->p, u->p, d->n));
    check(af_fill_dev(ctx(), dr->p, uR->p, dr->n));
    in_->bwdR(s, d, dr, rgrad, grad);
  }

 private:
  RResultPtr in_;
};

}  // namespace b200

inline ResultPtr Add(ResultPtr a, ResultPtr b) { return b200::Binary(b200::BinOp::kAdd, std::move(a), std::move(b)); }
inline RResultPtr AddR(RResultPtr a, RResultPtr b) { return b200::BinaryR(b200::BinOp::kAdd, std::move(a), std::move(b)); }
inline ResultPtr Sub(ResultPtr a, ResultPtr b) { return b200::Binary(b200::BinOp::kSub, std::move(a), std::move(b)); }
inline RResultPtr SubR(RResultPtr a, RResultPtr b) { return b200::BinaryR(b200::BinOp::kSub, std::move(a), std::move(b)); }
inline ResultPtr Mul(ResultPtr a, ResultPtr b) { return b200::Binary(b200::BinOp::kMul, std::move(a), std::move(b)); }
inline RResultPtr MulR(RResultPtr a, RResultPtr b) { return b200::BinaryR(b200::BinOp::kMul, std::move(a), std::move(b)); }

// arithmetic.go:436-493
inline ResultPtr Scale(ResultPtr in, double f) {
  b200::Dev out = in->dev()->Copy();
  b200::check(af_scale(b200::ctx(), out->p, f, out->n));
  return std::make_shared<b200::ScaledResult>(std::move(in), f, out);
}
inline RResultPtr ScaleR(RResultPtr in, double f) {
  b200::Dev out = in->dev()->Copy();
  b200::check(af_scale(b200::ctx(), out->p, f, out->n));
  b200::Dev xr = in->devR();
  b200::Dev outR = xr ? xr->Copy() : b200::DevVec::Zeros(out->n);
  if (xr) b200::check(af_scale(b200::ctx(), outR->p, f, outR->n));
  return std::make_shared<b200::ScaledRResult>(std::move(in), f, out, outR);
}
// arithmetic.go:184-252: the backward pass is the identity
inline ResultPtr AddScaler(ResultPtr in, double f) {
  b200::Dev x = in->dev();
  b200::Dev out = b200::DevVec::Empty(x->n);
  b200::check(af_add_scaler(b200::ctx(), x->p, f, out->p, x->n));
  return std::make_shared<b200::ScaledResult>(std::move(in), 1.0, out);
}
inline RResultPtr AddScalerR(RResultPtr in, double f) {
  b200::Dev x = in->dev(), xr = in->devR();
  b200::Dev out = b200::DevVec::Empty(x->n);
  b200::check(af_add_scaler(b200::ctx(), x->p, f, out->p, x->n));
  return std::make_shared<b200::ScaledRResult>(std::move(in), 1.0, out, xr ? xr : b200::DevVec::Zeros(x->n));
}
// arithmetic.go:972-1047
inline ResultPtr SumAll(ResultPtr in) {
  b200::Dev x = in->dev();
  b200::Dev out = b200::DevVec::Empty(1);
  b200::check(af_sum_all(b200::ctx(), x->p, nullptr, out->p, nullptr, x->n));
  return std::make_shared<b200::SumAllResult>(std::move(in), out);
}
inline RResultPtr SumAllR(RResultPtr in) {
  b200::Dev x = in->dev();
  b200::Dev xr = b200::ZerosIfNull(in->devR(), x->n);
  b200::Dev out = b200::DevVec::Empty(1), outR = b200::DevVec::Empty(1);
  b200::check(af_sum_all(b200::ctx(), x->p, xr->p, out->p, outR->p, x->n));
  return std::make_shared<b200::SumAllRResult>(std::move(in), out, outR);
}
// math_funcs.go:191-199
struct SquaredNorm : RFunc {
  ResultPtr Apply(ResultPtr in) override { return SumAll(Square(std::move(in))); }
  RResultPtr ApplyR(const RVector&, RResultPtr in) override { return SumAllR(SquareR(std::move(in))); }
};

// =================================================================================================
// Softmax (math_funcs.go:479-509).  As a Batcher every row of the packed batch is one softmax, in ONE kernel per
// direction (the reference goes through RFuncBatcher, batcher.go:57-84).
// =================================================================================================
namespace b200 {
class SoftmaxResult : public DeviceResult {
 public:
  SoftmaxResult(ResultPtr in, double t, int n, Dev y) : DeviceResult(std::move(y)), in_(std::move(in)), t_(t), n_(n) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {
    if (in_->Constant(grad)) return;
    check(af_softmax_bwd_r(ctx(), d_->p, nullptr, t_, u->p, nullptr, d_->n / n_, n_));
    in_->bwd(s, u, grad);
  }

 private:
  ResultPtr in_;
  double t_;
  int n_;
};
class SoftmaxRResult : public DeviceRResult {
 public:
  SoftmaxRResult(RResultPtr in, double t, int n, Dev y, Dev ry)
      : DeviceRResult(std::move(y), std::move(ry)), in_(std::move(in)), t_(t), n_(n) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {
    if (in_->Constant(rgrad, grad)) return;
    check(af_softmax_bwd_r(ctx(), d_->p, dr_->p, t_, u->p, uR->p, d_->n / n_, n_));
    in_->bwdR(s, u, uR, rgrad, grad);
  }

 private:
  RResultPtr in_;
  double t_;
  int n_;
};
}  // namespace b200

class Softmax : public RFunc, public RBatcher {
 public:
  double Temperature;  // 0 means 1 (math_funcs.go:481-485)
  explicit Softmax(double temperature = 0) : Temperature(temperature) {}
  ResultPtr Apply(ResultPtr in) override { return Batch(std::move(in), 1); }
  RResultPtr ApplyR(const RVector& v, RResultPtr in) override { return BatchR(v, std::move(in), 1); }
  ResultPtr Batch(ResultPtr in, int n) override {
    if (n <= 0 || in->len() % n != 0) throw Panic("input count does not divide input length");
    b200::Dev x = in->dev();
    b200::Dev y = b200::DevVec::Empty(x->n);
    b200::check(af_softmax_fwd_r(b200::ctx(), x->p, nullptr, Temperature, y->p, nullptr, x->n / n, n));
    return std::make_shared<b200::SoftmaxResult>(std::move(in), Temperature, n, y);
  }
  RResultPtr BatchR(const RVector&, RResultPtr in, int n) override {
    if (n <= 0 || in->len() % n != 0) throw Panic("input count does not divide input length");
    b200::Dev x = in->dev(), xr = in->devR();
    b200::Dev y = b200::DevVec::Empty(x->n), ry = b200::DevVec::Empty(x->n);
    b200::check(af_softmax_fwd_r(b200::ctx(), x->p, b200::ptr(xr), Temperature, y->p, ry->p, x->n / n, n));
    return std::make_shared<b200::SoftmaxRResult>(std::move(in), Temperature, n, y, ry);
  }
};

// =================================================================================================
// slices.go: Concat, Slice, Repeat
// =================================================================================================
namespace b200 {

class ConcatResult : public DeviceResult {
 public:
  ConcatResult(std::vector<ResultPtr> ins, Dev out) : DeviceResult(std::move(out)), ins_(std::move(ins)) {}
  bool Constant(const Gradient& g) override {
    for (auto& r : ins_)
      if (!r->Constant(g)) return false;
    return true;
  }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // slices.go:40-49: each input gets its slice
    int64_t at = 0;
    for (auto& r : ins_) {
      int64_t l = r->len();
      if (!r->Constant(grad)) r->bwd(s, DevVec::View(u, at, at + l), grad);
      at += l;
    }
  }

 private:
  std::vector<ResultPtr> ins_;
};

class ConcatRResult : public DeviceRResult {
 public:
  ConcatRResult(std::vector<RResultPtr> ins, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), ins_(std::move(ins)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override {
    for (auto& r : ins_)
      if (!r->Constant(rg, g)) return false;
    return true;
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // slices.go:104-119
    int64_t at = 0;
    for (auto& r : ins_) {
      int64_t l = r->len();
      if (!r->Constant(rgrad, grad)) r->bwdR(s, DevVec::View(u, at, at + l), DevVec::View(uR, at, at + l), rgrad, grad);
      at += l;
    }
  }

 private:
  std::vector<RResultPtr> ins_;
};

class SliceResult : public DeviceResult {
 public:
  SliceResult(ResultPtr in, int64_t start, int64_t end, Dev view)
      : DeviceResult(std::move(view)), in_(std::move(in)), start_(start), end_(end) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // slices.go:148-163: zero-padded upstream
    if (in_->Constant(grad)) return;
    Dev d = DevVec::Zeros(in_->len());
    check(af_copy(ctx(), d->p + start_, u->p, end_ - start_));
    in_->bwd(s, d, grad);
  }

 private:
  ResultPtr in_;
  int64_t start_, end_;
};

class SliceRResult : public DeviceRResult {
 public:
  SliceRResult(RResultPtr in, int64_t start, int64_t end, Dev view, Dev viewR)
      : DeviceRResult(std::move(view), std::move(viewR)), in_(std::move(in)), start_(start), end_(end) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // slices.go:186-206
    if (in_->Constant(rgrad, grad)) return;
    Dev d = DevVec::Zeros(in_->len()), dr = DevVec::Zeros(in_->len());
    check(af_copy(ctx(), d->p + start_, u->p, end_ - start_));
    check(af_copy(ctx(), dr->p + start_, uR->p, end_ - start_));
    in_->bwdR(s, d, dr, rgrad, grad);
  }

 private:
  RResultPtr in_;
  int64_t start_, end_;
};

class RepeatResult : public DeviceResult {
 public:
  RepeatResult(ResultPtr in, int n, Dev out) : DeviceResult(std::move(out)), in_(std::move(in)), n_(n) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // slices.go:244-250: the n parts are summed
    if (in_->Constant(grad)) return;
    Dev d = DevVec::Zeros(in_->len());
    check(af_bias_grad(ctx(), u->p, d->p, d->n, n_));
    in_->bwd(s, d, grad);
  }

 private:
  ResultPtr in_;
  int n_;
};

class RepeatRResult : public DeviceRResult {
 public:
  RepeatRResult(RResultPtr in, int n, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)), n_(n) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // slices.go:294-308
    if (in_->Constant(rgrad, grad)) return;
    Dev d = DevVec::Zeros(in_->len()), dr = DevVec::Zeros(in_->len());
    check(af_bias_grad_r(ctx(), u->p, uR->p, d->p, dr->p, d->n, n_));
    in_->bwdR(s, d, dr, rgrad, grad);
  }

 private:
  RResultPtr in_;
  int n_;
};

inline Dev RepeatDev(const Dev& x, int n) {
  Dev out = DevVec::Empty(x->n * n);
  // n is small in the reference's uses (batch-size repeats of a bias vector)
  for (int i = 0; i < n && x->n > 0; i++) check(af_copy(ctx(), out->p + (int64_t)i * x->n, x->p, x->n));
  return out;
}

}  // namespace b200

// slices.go:13-63
inline ResultPtr Concat(std::vector<ResultPtr> ins) {
  int64_t total = 0;
  for (auto& r : ins) total += r->len();
  b200::Dev out = b200::DevVec::Empty(total);
  int64_t at = 0;
  for (auto& r : ins) {
    b200::Dev x = r->dev();
    if (x->n > 0) b200::check(af_copy(b200::ctx(), out->p + at, x->p, x->n));
    at += x->n;
  }
  return std::make_shared<b200::ConcatResult>(std::move(ins), out);
}
// slices.go:65-119
inline RResultPtr ConcatR(std::vector<RResultPtr> ins) {
  int64_t total = 0;
  for (auto& r : ins) total += r->len();
  b200::Dev out = b200::DevVec::Empty(total), outR = b200::DevVec::Empty(total);
  int64_t at = 0;
  for (auto& r : ins) {
    b200::Dev x = r->dev(), xr = r->devR();
    if (x->n > 0) {
      b200::check(af_copy(b200::ctx(), out->p + at, x->p, x->n));
      if (xr) b200::check(af_copy(b200::ctx(), outR->p + at, xr->p, x->n));
      else b200::check(af_zero(b200::ctx(), outR->p + at, x->n));
    }
    at += x->n;
  }
  return std::make_shared<b200::ConcatRResult>(std::move(ins), out, outR);
}
// slices.go:130-163: a view, no copy
inline ResultPtr Slice(ResultPtr in, int64_t start, int64_t end) {
  if (start < 0 || end < start || end > in->len()) throw Panic("slice bounds out of range");
  b200::Dev v = b200::DevVec::View(in->dev(), start, end);
  return std::make_shared<b200::SliceResult>(std::move(in), start, end, v);
}
// slices.go:165-206
inline RResultPtr SliceR(RResultPtr in, int64_t start, int64_t end) {
  if (start < 0 || end < start || end > in->len()) throw Panic("slice bounds out of range");
  b200::Dev v = b200::DevVec::View(in->dev(), start, end);
  b200::Dev xr = in->devR();
  b200::Dev vr = xr ? b200::DevVec::View(xr, start, end) : nullptr;
  return std::make_shared<b200::SliceRResult>(std::move(in), start, end, v, vr);
}
// slices.go:216-259
inline ResultPtr Repeat(ResultPtr in, int n) {
  b200::Dev out = b200::RepeatDev(in->dev(), n);
  return std::make_shared<b200::RepeatResult>(std::move(in), n, out);
}
// slices.go:261-308
inline RResultPtr RepeatR(RResultPtr in, int n) {
  b200::Dev out = b200::RepeatDev(in->dev(), n);
  b200::Dev xr = in->devR();
  b200::Dev outR = xr ? b200::RepeatDev(xr, n) : b200::DevVec::Zeros(out->n);
  return std::make_shared<b200::RepeatRResult>(std::move(in), n, out, outR);
}

// Per-sample concatenation of packed batches: sample i of the result is [a_i | b_i | ...] -- what Concat produces under
// a batcher (batcher.go:57-84) and the layout bench/lstm.go:140,188 builds per lane; one strided copy per part.
namespace b200 {
inline std::vector<int64_t> BatchedWidths(const std::vector<int64_t>& lens, int n) {
  std::vector<int64_t> w;
  for (int64_t l : lens) {
    if (n <= 0 || l % n != 0) throw Panic("input count does not divide input length");  // batcher.go:41
    w.push_back(l / n);
  }
  return w;
}
inline Dev BatchedPart(const Dev& u, int64_t off, int64_t w, int64_t total, int n) {
  Dev out = DevVec::Empty((int64_t)n * w);
  if (out->n > 0) check(af_copy_2d(ctx(), out->p, w, u->p + off, total, n, w));
  return out;
}
class ConcatBatchedResult : public DeviceResult {
 public:
  ConcatBatchedResult(std::vector<ResultPtr> ins, std::vector<int64_t> widths, int n, Dev out)
      : DeviceResult(std::move(out)), ins_(std::move(ins)), widths_(std::move(widths)), n_(n) {}
  bool Constant(const Gradient& g) override {
    for (auto& r : ins_)
      if (!r->Constant(g)) return false;
    return true;
  }
  void bwd(Session& s, Dev u, Gradient& grad) override {
    int64_t total = d_->n / (n_ > 0 ? n_ : 1), off = 0;
    for (size_t i = 0; i < ins_.size(); off += widths_[i], i++)
      if (!ins_[i]->Constant(grad)) ins_[i]->bwd(s, BatchedPart(u, off, widths_[i], total, n_), grad);
  }

 private:
  std::vector<ResultPtr> ins_;
  std::vector<int64_t> widths_;
  int n_;
};
class ConcatBatchedRResult : public DeviceRResult {
 public:
  ConcatBatchedRResult(std::vector<RResultPtr> ins, std::vector<int64_t> widths, int n, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), ins_(std::move(ins)), widths_(std::move(widths)), n_(n) {}
  bool Constant(const RGradient& rg, const Gradient* g) override {
    for (auto& r : ins_)
      if (!r->Constant(rg, g)) return false;
    return true;
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {
    int64_t total = d_->n / (n_ > 0 ? n_ : 1), off = 0;
    for (size_t i = 0; i < ins_.size(); off += widths_[i], i++)
      if (!ins_[i]->Constant(rgrad, grad))
        ins_[i]->bwdR(s, BatchedPart(u, off, widths_[i], total, n_), BatchedPart(uR, off, widths_[i], total, n_), rgrad, grad);
  }

 private:
  std::vector<RResultPtr> ins_;
  std::vector<int64_t> widths_;
  int n_;
};
}  // namespace b200

inline ResultPtr ConcatBatched(int n, std::vector<ResultPtr> ins) {
  std::vector<int64_t> lens;
  for (auto& r : ins) lens.push_back(r->len());
  std::vector<int64_t> widths = b200::BatchedWidths(lens, n);
  int64_t total = 0;
  for (int64_t w : widths) total += w;
  b200::Dev out = b200::DevVec::Empty((int64_t)n * total);
  int64_t off = 0;
  for (size_t i = 0; i < ins.size(); off += widths[i], i++)
    if (widths[i] > 0) b200::check(af_copy_2d(b200::ctx(), out->p + off, total, ins[i]->dev()->p, widths[i], n, widths[i]));
  return std::make_shared<b200::ConcatBatchedResult>(std::move(ins), std::move(widths), n, out);
}
inline RResultPtr ConcatBatchedR(int n, std::vector<RResultPtr> ins) {
  std::vector<int64_t> lens;
  for (auto& r : ins) lens.push_back(r->len());
  std::vector<int64_t> widths = b200::BatchedWidths(lens, n);
  int64_t total = 0;
  for (int64_t w : widths) total += w;
  b200::Dev out = b200::DevVec::Empty((int64_t)n * total), outR = b200::DevVec::Zeros((int64_t)n * total);
  int64_t off = 0;
  for (size_t i = 0; i < ins.size(); off += widths[i], i++) {
    if (widths[i] == 0) continue;
    b200::check(af_copy_2d(b200::ctx(), out->p + off, total, ins[i]->dev()->p, widths[i], n, widths[i]));
    b200::Dev xr = ins[i]->devR();
    if (xr) b200::check(af_copy_2d(b200::ctx(), outR->p + off, total, xr->p, widths[i], n, widths[i]));
  }
  return std::make_shared<b200::ConcatBatchedRResult>(std::move(ins), std::move(widths), n, out, outR);
}

// batcher.go:34-84: the naive adaptor -- split the packed input, apply F per sample, concatenate.
class FuncBatcher : public Batcher {
 public:
  std::shared_ptr<Func> F;
  explicit FuncBatcher(std::shared_ptr<Func> f) : F(std::move(f)) {}
  ResultPtr Batch(ResultPtr in, int n) override {
    if (n <= 0 || in->len() % n != 0) throw Panic("input count does not divide input length");  // batcher.go:41
    int64_t each = in->len() / n;
    std::vector<ResultPtr> outs;
    for (int i = 0; i < n; i++) outs.push_back(F->Apply(Slice(in, i * each, (i + 1) * each)));
    return Concat(std::move(outs));
  }
};
class RFuncBatcher : public RBatcher {
 public:
  std::shared_ptr<RFunc> F;
  explicit RFuncBatcher(std::shared_ptr<RFunc> f) : F(std::move(f)) {}
  ResultPtr Batch(ResultPtr in, int n) override { return FuncBatcher(F).Batch(std::move(in), n); }
  RResultPtr BatchR(const RVector& v, RResultPtr in, int n) override {
    if (n <= 0 || in->len() % n != 0) throw Panic("input count does not divide input length");  // batcher.go:68
    int64_t each = in->len() / n;
    std::vector<RResultPtr> outs;
    for (int i = 0; i < n; i++) outs.push_back(F->ApplyR(v, SliceR(in, i * each, (i + 1) * each)));
    return ConcatR(std::move(outs));
  }
};

// =================================================================================================
// SURVEY 8(f) rank 3, either side of the path: Cos, Norm, Inverse, Pow, Div, ScaleFirst, AddFirst, the log-domain
// sums, the squared-error head (math_funcs.go:201-284,599-607; arithmetic.go:378-427,495-694,772-966,1049-1090)
// =================================================================================================
namespace b200 {

// Inverse's backward pass reads the forward OUTPUT and the INPUT's R-vector (arithmetic.go:800-807,851-867).
inline const UnaryKernels* InverseKernels() {
  static const UnaryKernels k{af_inverse_fwd_r, af_inverse_bwd_r, true};
  return &k;
}

class PowResult : public DeviceResult {
 public:
  PowResult(ResultPtr in, double p, Dev x, Dev out) : DeviceResult(std::move(out)), in_(std::move(in)), x_(std::move(x)), p_(p) {}
  bool Constant(const Gradient& g) override { return p_ != 0 && in_->Constant(g); }  // arithmetic.go:892-894
  void bwd(Session& s, Dev u, Gradient& grad) override {                             // arithmetic.go:896-906
    if (Constant(grad)) return;
    check(af_pow_bwd_r(ctx(), x_->p, nullptr, p_, u->p, nullptr, u->n));
    in_->bwd(s, u, grad);
  }

 private:
  ResultPtr in_;
  Dev x_;
  double p_;
};
class PowRResult : public DeviceRResult {
 public:
  PowRResult(RResultPtr in, double p, Dev x, Dev xr, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)), x_(std::move(x)), xr_(std::move(xr)), p_(p) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return p_ != 0 && in_->Constant(rg, g); }  // :944-946
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {                            // :948-966
    if (Constant(rgrad, grad)) return;
    check(af_pow_bwd_r(ctx(), x_->p, ptr(xr_), p_, u->p, uR->p, u->n));
    in_->bwdR(s, u, uR, rgrad, grad);
  }

 private:
  RResultPtr in_;
  Dev x_, xr_;
  double p_;
};

class QuotientResult : public DeviceResult {
 public:
  QuotientResult(ResultPtr num, ResultPtr denom, Dev b, Dev out)
      : DeviceResult(std::move(out)), num_(std::move(num)), denom_(std::move(denom)), b_(std::move(b)) {}
  bool Constant(const Gradient& g) override { return num_->Constant(g) && denom_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // arithmetic.go:407-423
    bool na = !num_->Constant(grad), nb = !denom_->Constant(grad);
    if (!na && !nb) return;
    Dev da = na ? DevVec::Empty(u->n) : nullptr, db = nb ? DevVec::Empty(u->n) : nullptr;
    check(af_div_bwd(ctx(), u->p, d_->p, b_->p, mptr(da), mptr(db), u->n));
    if (na) num_->bwd(s, da, grad);
    if (nb) denom_->bwd(s, db, grad);
  }

 private:
  ResultPtr num_, denom_;
  Dev b_;
};

// ScaleFirst (arithmetic.go:495-593): the scaler is read where it lives, on the device.
class ScaleFirstResult : public DeviceResult {
 public:
  ScaleFirstResult(ResultPtr in, ResultPtr scaler, Dev x, Dev f, Dev out)
      : DeviceResult(std::move(out)), in_(std::move(in)), scaler_(std::move(scaler)), x_(std::move(x)), f_(std::move(f)) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g) && scaler_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // arithmetic.go:520-533
    if (!scaler_->Constant(grad)) {
      Dev down = DevVec::Zeros(scaler_->len());
      check(af_dot(ctx(), u->p, x_->p, nullptr, nullptr, down->p, u->n));
      scaler_->bwd(s, down, grad);
    }
    if (!in_->Constant(grad)) {  // after the scaler's branch, so the upstream can be scaled in place
      check(af_scale_dev(ctx(), u->p, nullptr, f_->p, nullptr, u->p, nullptr, u->n));
      in_->bwd(s, u, grad);
    }
  }

 private:
  ResultPtr in_, scaler_;
  Dev x_, f_;
};
class ScaleFirstRResult : public DeviceRResult {
 public:
  ScaleFirstRResult(RResultPtr in, RResultPtr scaler, Dev x, Dev xr, Dev f, Dev fr, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)), scaler_(std::move(scaler)), x_(std::move(x)),
        xr_(std::move(xr)), f_(std::move(f)), fr_(std::move(fr)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g) && scaler_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // arithmetic.go:567-593
    if (!scaler_->Constant(rgrad, grad)) {
      Dev down = DevVec::Zeros(scaler_->len()), downR = DevVec::Zeros(scaler_->len());
      check(af_dot(ctx(), u->p, x_->p, nullptr, nullptr, down->p, u->n));
      check(af_dot(ctx(), uR->p, x_->p, xr_ ? u->p : nullptr, ptr(xr_), downR->p, u->n));
      scaler_->bwdR(s, down, downR, rgrad, grad);
    }
    if (!in_->Constant(rgrad, grad)) {
      check(af_scale_dev(ctx(), u->p, uR->p, f_->p, ptr(fr_), u->p, uR->p, u->n));
      in_->bwdR(s, u, uR, rgrad, grad);
    }
  }

 private:
  RResultPtr in_, scaler_;
  Dev x_, xr_, f_, fr_;
};

// AddFirst (arithmetic.go:595-694)
class AddFirstResult : public DeviceResult {
 public:
  AddFirstResult(ResultPtr in, ResultPtr scaler, Dev out)
      : DeviceResult(std::move(out)), in_(std::move(in)), scaler_(std::move(scaler)) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g) && scaler_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // arithmetic.go:627-639
    if (!scaler_->Constant(grad)) {
      Dev su = DevVec::Zeros(scaler_->len());
      check(af_sum_all(ctx(), u->p, nullptr, su->p, nullptr, u->n));
      scaler_->bwd(s, su, grad);
    }
    if (!in_->Constant(grad)) in_->bwd(s, u, grad);
  }

 private:
  ResultPtr in_, scaler_;
};
class AddFirstRResult : public DeviceRResult {
 public:
  AddFirstRResult(RResultPtr in, RResultPtr scaler, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)), scaler_(std::move(scaler)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g) && scaler_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // arithmetic.go:680-694
    if (!scaler_->Constant(rgrad, grad)) {
      Dev su = DevVec::Zeros(scaler_->len()), suR = DevVec::Zeros(scaler_->len());
      check(af_sum_all(ctx(), u->p, uR->p, su->p, suR->p, u->n));
      scaler_->bwdR(s, su, suR, rgrad, grad);
    }
    if (!in_->Constant(rgrad, grad)) in_->bwdR(s, u, uR, rgrad, grad);
  }

 private:
  RResultPtr in_, scaler_;
};

class NormResult : public DeviceResult {
 public:
  NormResult(ResultPtr in, Dev x, Dev out) : DeviceResult(std::move(out)), in_(std::move(in)), x_(std::move(x)) {}
  bool Constant(const Gradient& g) override { return in_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // math_funcs.go:238-247
    if (in_->Constant(grad)) return;
    Dev d = DevVec::Empty(x_->n);
    check(af_norm_bwd_r(ctx(), x_->p, nullptr, d_->p, nullptr, u->p, nullptr, d->p, nullptr, x_->n));
    in_->bwd(s, d, grad);
  }

 private:
  ResultPtr in_;
  Dev x_;
};
class NormRResult : public DeviceRResult {
 public:
  NormRResult(RResultPtr in, Dev x, Dev xr, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), in_(std::move(in)), x_(std::move(x)), xr_(std::move(xr)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return in_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // math_funcs.go:267-284
    if (in_->Constant(rgrad, grad)) return;
    Dev d = DevVec::Empty(x_->n), dr = DevVec::Empty(x_->n);
    check(af_norm_bwd_r(ctx(), x_->p, ptr(xr_), d_->p, dr_->p, u->p, uR->p, d->p, dr->p, x_->n));
    in_->bwdR(s, d, dr, rgrad, grad);
  }

 private:
  RResultPtr in_;
  Dev x_, xr_;
};

class SquaredErrorResult : public DeviceResult {
 public:
  SquaredErrorResult(ResultPtr a, ResultPtr y, Dev da, Dev dy, Dev out)
      : DeviceResult(std::move(out)), a_(std::move(a)), y_(std::move(y)), da_(std::move(da)), dy_(std::move(dy)) {}
  bool Constant(const Gradient& g) override { return a_->Constant(g) && y_->Constant(g); }
  void bwd(Session& s, Dev u, Gradient& grad) override {
    ResultPtr who[2] = {a_, y_};
    for (int i = 0; i < 2; i++) {
      if (who[i]->Constant(grad)) continue;
      Dev g = DevVec::Empty(da_->n);
      check(af_sqerr_bwd_r(ctx(), da_->p, nullptr, dy_->p, nullptr, u->p, nullptr, i == 0 ? 1.0 : -1.0, g->p, nullptr, da_->n));
      who[i]->bwd(s, g, grad);
    }
  }

 private:
  ResultPtr a_, y_;
  Dev da_, dy_;
};
class SquaredErrorRResult : public DeviceRResult {
 public:
  SquaredErrorRResult(RResultPtr a, RResultPtr y, Dev da, Dev dar, Dev dy, Dev dyr, Dev out, Dev outR)
      : DeviceRResult(std::move(out), std::move(outR)), a_(std::move(a)), y_(std::move(y)), da_(std::move(da)),
        dar_(std::move(dar)), dy_(std::move(dy)), dyr_(std::move(dyr)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return a_->Constant(rg, g) && y_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {
    RResultPtr who[2] = {a_, y_};
    for (int i = 0; i < 2; i++) {
      if (who[i]->Constant(rgrad, grad)) continue;
      Dev g = DevVec::Empty(da_->n), gr = DevVec::Empty(da_->n);
      check(af_sqerr_bwd_r(ctx(), da_->p, ptr(dar_), dy_->p, ptr(dyr_), u->p, uR->p, i == 0 ? 1.0 : -1.0, g->p, gr->p,
                           da_->n));
      who[i]->bwdR(s, g, gr, rgrad, grad);
    }
  }

 private:
  RResultPtr a_, y_;
  Dev da_, dar_, dy_, dyr_;
};

// out = x + sign * m[0] with the upstream passing through untouched: the "- maxVal" / "+ maxVal" steps of the log-domain
// sums (arithmetic.go:1056-1090; maxVal is a constant of the graph, kept on the device).
inline ResultPtr Shift(ResultPtr in, const Dev& m, double sign) {
  Dev x = in->dev();
  Dev out = DevVec::Empty(x->n);
  check(af_add_dev(ctx(), x->p, m->p, sign, out->p, x->n));
  return std::make_shared<ScaledResult>(std::move(in), 1.0, out);
}
inline RResultPtr ShiftR(RResultPtr in, const Dev& m, double sign) {
  Dev x = in->dev(), xr = in->devR();
  Dev out = DevVec::Empty(x->n);
  check(af_add_dev(ctx(), x->p, m->p, sign, out->p, x->n));
  return std::make_shared<ScaledRResult>(std::move(in), 1.0, out, xr ? xr : DevVec::Zeros(x->n));
}
inline Dev MaxAbs(std::initializer_list<Dev> xs) {  // linalg.Vector.MaxAbs over all the vectors
  Dev m = DevVec::Zeros(1);
  int i = 0;
  for (const Dev& x : xs) check(af_max_abs(ctx(), x->p, m->p, i++ > 0, x->n));
  return m;
}

}  // namespace b200

// math_funcs.go:599-607: Cos(x) = Sin(x + pi/2)
struct Cos : RFunc {
  ResultPtr Apply(ResultPtr in) override { return Sin().Apply(AddScaler(std::move(in), 3.14159265358979323846 / 2)); }
  RResultPtr ApplyR(const RVector& v, RResultPtr in) override {
    return Sin().ApplyR(v, AddScalerR(std::move(in), 3.14159265358979323846 / 2));
  }
};

// math_funcs.go:201-284
struct Norm : RFunc {
  ResultPtr Apply(ResultPtr in) override {
    b200::Dev x = in->dev(), o = b200::DevVec::Empty(1);
    b200::check(af_norm_fwd_r(b200::ctx(), x->p, nullptr, o->p, nullptr, x->n));
    return std::make_shared<b200::NormResult>(std::move(in), x, o);
  }
  RResultPtr ApplyR(const RVector&, RResultPtr in) override {
    b200::Dev x = in->dev(), xr = in->devR(), o = b200::DevVec::Empty(1), ro = b200::DevVec::Empty(1);
    b200::check(af_norm_fwd_r(b200::ctx(), x->p, b200::ptr(xr), o->p, ro->p, x->n));
    return std::make_shared<b200::NormRResult>(std::move(in), x, xr, o, ro);
  }
};

// arithmetic.go:772-867
inline ResultPtr Inverse(ResultPtr in) {
  b200::Dev x = in->dev(), o = b200::DevVec::Empty(x->n);
  b200::check(af_inverse_fwd_r(b200::ctx(), x->p, nullptr, o->p, nullptr, x->n));
  return std::make_shared<b200::UnaryResult>(b200::InverseKernels(), std::move(in), o, o);
}
inline RResultPtr InverseR(RResultPtr in) {
  b200::Dev x = in->dev(), xr = in->devR(), o = b200::DevVec::Empty(x->n), ro = b200::DevVec::Empty(x->n);
  b200::check(af_inverse_fwd_r(b200::ctx(), x->p, b200::ptr(xr), o->p, ro->p, x->n));
  return std::make_shared<b200::UnaryRResult>(b200::InverseKernels(), std::move(in), o, xr, o, ro);
}
// arithmetic.go:869-966
inline ResultPtr Pow(ResultPtr in, double p) {
  b200::Dev x = in->dev(), o = b200::DevVec::Empty(x->n);
  b200::check(af_pow_fwd_r(b200::ctx(), x->p, nullptr, p, o->p, nullptr, x->n));
  return std::make_shared<b200::PowResult>(std::move(in), p, x, o);
}
inline RResultPtr PowR(RResultPtr in, double p) {
  b200::Dev x = in->dev(), xr = in->devR(), o = b200::DevVec::Empty(x->n), ro = b200::DevVec::Empty(x->n);
  b200::check(af_pow_fwd_r(b200::ctx(), x->p, b200::ptr(xr), p, o->p, ro->p, x->n));
  return std::make_shared<b200::PowRResult>(std::move(in), p, x, xr, o, ro);
}
// arithmetic.go:378-427
inline ResultPtr Div(ResultPtr a, ResultPtr b) {
  b200::SameLen(a->len(), b->len());
  b200::Dev x = a->dev(), y = b->dev(), o = b200::DevVec::Empty(x->n);
  b200::check(af_div(b200::ctx(), x->p, y->p, o->p, x->n));
  return std::make_shared<b200::QuotientResult>(std::move(a), std::move(b), y, o);
}
inline RResultPtr DivR(RResultPtr a, RResultPtr b) { return MulR(std::move(a), InverseR(std::move(b))); }
// arithmetic.go:495-593: `in` scaled by the first (only) component of `scaler`
inline ResultPtr ScaleFirst(ResultPtr in, ResultPtr scaler) {
  b200::Dev x = in->dev(), f = scaler->dev(), o = b200::DevVec::Empty(x->n);
  b200::check(af_scale_dev(b200::ctx(), x->p, nullptr, f->p, nullptr, o->p, nullptr, x->n));
  return std::make_shared<b200::ScaleFirstResult>(std::move(in), std::move(scaler), x, f, o);
}
inline RResultPtr ScaleFirstR(RResultPtr in, RResultPtr scaler) {
  b200::Dev x = in->dev(), xr = in->devR(), f = scaler->dev(), fr = scaler->devR();
  b200::Dev o = b200::DevVec::Empty(x->n), ro = b200::DevVec::Empty(x->n);
  b200::check(af_scale_dev(b200::ctx(), x->p, b200::ptr(xr), f->p, b200::ptr(fr), o->p, ro->p, x->n));
  return std::make_shared<b200::ScaleFirstRResult>(std::move(in), std::move(scaler), x, xr, f, fr, o, ro);
}
// arithmetic.go:595-694: the first component of v2 added to every component of v1
inline ResultPtr AddFirst(ResultPtr v1, ResultPtr v2) {
  b200::Dev x = v1->dev(), o = b200::DevVec::Empty(x->n);
  b200::check(af_add_dev(b200::ctx(), x->p, v2->dev()->p, 1.0, o->p, x->n));
  return std::make_shared<b200::AddFirstResult>(std::move(v1), std::move(v2), o);
}
inline RResultPtr AddFirstR(RResultPtr v1, RResultPtr v2) {
  b200::Dev x = v1->dev(), xr = b200::ZerosIfNull(v1->devR(), x->n), fr = v2->devR();
  b200::Dev o = b200::DevVec::Empty(x->n), ro;
  b200::check(af_add_dev(b200::ctx(), x->p, v2->dev()->p, 1.0, o->p, x->n));
  if (fr) {
    ro = b200::DevVec::Empty(x->n);
    b200::check(af_add_dev(b200::ctx(), xr->p, fr->p, 1.0, ro->p, x->n));
  } else {
    ro = xr->Copy();
  }
  return std::make_shared<b200::AddFirstRResult>(std::move(v1), std::move(v2), o, ro);
}
// arithmetic.go:1049-1072: log(exp(v1) + exp(v2)) without overflow
inline ResultPtr AddLogDomain(ResultPtr v1, ResultPtr v2) {
  b200::SameLen(v1->len(), v2->len());
  b200::Dev m = b200::MaxAbs({v1->dev(), v2->dev()});
  ResultPtr e1 = Exp().Apply(b200::Shift(std::move(v1), m, -1.0)), e2 = Exp().Apply(b200::Shift(std::move(v2), m, -1.0));
  return b200::Shift(Log().Apply(Add(e1, e2)), m, 1.0);
}
inline RResultPtr AddLogDomainR(RResultPtr v1, RResultPtr v2) {
  static const RVector none;
  b200::SameLen(v1->len(), v2->len());
  b200::Dev m = b200::MaxAbs({v1->dev(), v2->dev()});
  RResultPtr e1 = Exp().ApplyR(none, b200::ShiftR(std::move(v1), m, -1.0));
  RResultPtr e2 = Exp().ApplyR(none, b200::ShiftR(std::move(v2), m, -1.0));
  return b200::ShiftR(Log().ApplyR(none, AddR(e1, e2)), m, 1.0);
}
// arithmetic.go:1074-1090: log(sum(exp(v)))
inline ResultPtr SumAllLogDomain(ResultPtr v) {
  b200::Dev m = b200::MaxAbs({v->dev()});
  return b200::Shift(Log().Apply(SumAll(Exp().Apply(b200::Shift(std::move(v), m, -1.0)))), m, 1.0);
}
inline RResultPtr SumAllLogDomainR(RResultPtr v) {
  static const RVector none;
  b200::Dev m = b200::MaxAbs({v->dev()});
  return b200::ShiftR(Log().ApplyR(none, SumAllR(Exp().ApplyR(none, b200::ShiftR(std::move(v), m, -1.0)))), m, 1.0);
}
// north_star's "SquaredError": .5 ||actual - expected||^2 = Scale(SquaredNorm(Sub(a, y)), .5) of the reference in one
// reduction forward and one elementwise pass backward.
inline ResultPtr SquaredError(ResultPtr actual, ResultPtr expected) {
  b200::SameLen(actual->len(), expected->len());
  b200::Dev a = actual->dev(), y = expected->dev(), o = b200::DevVec::Empty(1);
  b200::check(af_sqerr_fwd_r(b200::ctx(), a->p, nullptr, y->p, nullptr, o->p, nullptr, a->n));
  return std::make_shared<b200::SquaredErrorResult>(std::move(actual), std::move(expected), a, y, o);
}
inline RResultPtr SquaredErrorR(RResultPtr actual, RResultPtr expected) {
  b200::SameLen(actual->len(), expected->len());
  b200::Dev a = actual->dev(), ar = actual->devR(), y = expected->dev(), yr = expected->devR();
  b200::Dev o = b200::DevVec::Empty(1), ro = b200::DevVec::Empty(1);
  b200::check(af_sqerr_fwd_r(b200::ctx(), a->p, b200::ptr(ar), y->p, b200::ptr(yr), o->p, ro->p, a->n));
  return std::make_shared<b200::SquaredErrorRResult>(std::move(actual), std::move(expected), a, ar, y, yr, o, ro);
}

// slices.go:310-335
inline std::vector<ResultPtr> Split(int n, ResultPtr in) {
  if (n <= 0 || in->len() % n != 0) throw Panic("count does not divide input length");
  int64_t each = in->len() / n;
  std::vector<ResultPtr> parts;
  for (int i = 0; i < n; i++) parts.push_back(Slice(in, i * each, (i + 1) * each));
  return parts;
}
inline std::vector<RResultPtr> SplitR(int n, RResultPtr in) {
  if (n <= 0 || in->len() % n != 0) throw Panic("count does not divide input length");
  int64_t each = in->len() / n;
  std::vector<RResultPtr> parts;
  for (int i = 0; i < n; i++) parts.push_back(SliceR(in, i * each, (i + 1) * each));
  return parts;
}

// =================================================================================================
// Pool / PoolAll / PoolSplit (pool.go), Fold (fold.go), MatMulVecs (lin_alg.go:14-133): graph logic around the
// temp-variable protocol -- insert grad[tmp], propagate, read, delete.  The pool variable shares the producing node's
// device buffer (Variable::Wrap) and what accumulates for it is taken from the backward pass's device accumulators
// (Session::Take), so neither direction leaves HBM; a foreign host node in the chain still works (it adds into the
// map's host vector, which Take folds in).
// =================================================================================================
namespace b200 {

inline VariablePtr PoolVar(Result& r) { return Variable::Wrap(r.dev()); }
inline RVariablePtr PoolRVar(RResult& r) { return RVariable::Wrap(Variable::Wrap(r.dev()), r.devR()); }
// Gradient{v: linalg.Vector{}} at pool.go:107,117,159,174
inline Gradient Marker(Variable* v) {
  Gradient g;
  g[v] = linalg::Vector();
  return g;
}

class PooledResult : public DeviceResult {
 public:
  PooledResult(std::vector<ResultPtr> ins, std::vector<VariablePtr> vars, ResultPtr out)
      : DeviceResult(out->dev()), ins_(std::move(ins)), vars_(std::move(vars)), out_(std::move(out)) {}
  bool Constant(const Gradient& g) override {  // pool.go:102-112
    if (!out_->Constant(g)) return false;
    for (size_t i = 0; i < vars_.size(); i++)
      if (!out_->Constant(Marker(vars_[i].get())) && !ins_[i]->Constant(g)) return false;
    return true;
  }
  void bwd(Session& s, Dev u, Gradient& grad) override {  // pool.go:114-136
    std::vector<bool> constant(vars_.size());
    for (size_t i = 0; i < vars_.size(); i++) {
      constant[i] = out_->Constant(Marker(vars_[i].get())) || ins_[i]->Constant(grad);
      if (!constant[i]) grad[vars_[i].get()] = linalg::Vector();
    }
    out_->bwd(s, std::move(u), grad);
    std::vector<Dev> ups(vars_.size());
    for (size_t i = 0; i < vars_.size(); i++)
      if (!constant[i]) ups[i] = s.Take(grad, vars_[i].get(), vars_[i]->len());
    for (size_t i = 0; i < vars_.size(); i++)
      if (!constant[i]) ins_[i]->bwd(s, ups[i], grad);
  }

 private:
  std::vector<ResultPtr> ins_;
  std::vector<VariablePtr> vars_;
  ResultPtr out_;
};

class PooledRResult : public DeviceRResult {
 public:
  PooledRResult(std::vector<RResultPtr> ins, std::vector<VariablePtr> vars, RResultPtr out)
      : DeviceRResult(out->dev(), out->devR()), ins_(std::move(ins)), vars_(std::move(vars)), out_(std::move(out)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override {  // pool.go:150-160
    if (!out_->Constant(rg, g)) return false;
    for (size_t i = 0; i < vars_.size(); i++)
      if (!out_->Constant(Marker(vars_[i].get()), nullptr) && !ins_[i]->Constant(rg, g)) return false;
    return true;
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // pool.go:162-195
    Gradient temp;
    if (grad == nullptr) grad = &temp;  // pool.go:164-166
    std::vector<bool> constant(vars_.size());
    for (size_t i = 0; i < vars_.size(); i++) {
      constant[i] = out_->Constant(Marker(vars_[i].get()), nullptr) || ins_[i]->Constant(rgrad, grad);
      if (!constant[i]) (*grad)[vars_[i].get()] = linalg::Vector(), rgrad[vars_[i].get()] = linalg::Vector();
    }
    out_->bwdR(s, std::move(u), std::move(uR), rgrad, grad);
    std::vector<Dev> ups(vars_.size()), upsR(vars_.size());
    for (size_t i = 0; i < vars_.size(); i++)
      if (!constant[i]) {
        ups[i] = s.Take(*grad, vars_[i].get(), vars_[i]->len());
        upsR[i] = s.Take(rgrad, vars_[i].get(), vars_[i]->len());
      }
    for (size_t i = 0; i < vars_.size(); i++)
      if (!constant[i]) ins_[i]->bwdR(s, ups[i], upsR[i], rgrad, grad);
    if (grad == &temp) s.Forget(temp);
  }

 private:
  std::vector<RResultPtr> ins_;
  std::vector<VariablePtr> vars_;
  RResultPtr out_;
};

class FoldResult : public DeviceResult {
 public:
  FoldResult(std::vector<VariablePtr> pool, std::vector<ResultPtr> inter, ResultPtr final)
      : DeviceResult(final->dev()), pool_(std::move(pool)), inter_(std::move(inter)), final_(std::move(final)) {}
  bool Constant(const Gradient& cg) override {  // fold.go:35-54 (the map is borrowed for the temporary, as in Go)
    Gradient& g = const_cast<Gradient&>(cg);
    if (!final_->Constant(g)) return false;
    for (size_t i = pool_.size(); i-- > 0;) {
      if (!inter_[i]->Constant(g)) return false;
      if (i > 0) {
        Variable* p = pool_[i - 1].get();
        g[p] = linalg::Vector();
        bool c = inter_[i]->Constant(g);
        g.erase(p);
        if (c) return true;
      }
    }
    return true;
  }
  void bwd(Session& s, Dev u, Gradient& g) override {  // fold.go:56-84
    if (pool_.empty()) return final_->bwd(s, std::move(u), g);
    Dev stateUp = std::move(u);
    for (size_t i = pool_.size(); i > 0; i--) {
      Result& res = i == pool_.size() ? *final_ : *inter_[i];
      Variable* p = pool_[i - 1].get();
      g[p] = linalg::Vector();
      if (res.Constant(g)) {
        g.erase(p);
        break;
      }
      res.bwd(s, stateUp, g);
      stateUp = s.Take(g, p, p->len());
    }
    inter_[0]->bwd(s, stateUp, g);
  }

 private:
  std::vector<VariablePtr> pool_;
  std::vector<ResultPtr> inter_;
  ResultPtr final_;
};

class FoldRResult : public DeviceRResult {
 public:
  FoldRResult(std::vector<VariablePtr> pool, std::vector<RResultPtr> inter, RResultPtr final)
      : DeviceRResult(final->dev(), final->devR()), pool_(std::move(pool)), inter_(std::move(inter)),
        final_(std::move(final)) {}
  bool Constant(const RGradient& crg, const Gradient* cg) override {  // fold.go:113-139
    RGradient& rg = const_cast<RGradient&>(crg);
    Gradient* g = const_cast<Gradient*>(cg);
    if (!final_->Constant(rg, g)) return false;
    for (size_t i = pool_.size(); i-- > 0;) {
      if (!inter_[i]->Constant(rg, g)) return false;
      if (i > 0) {
        Variable* p = pool_[i - 1].get();
        if (g != nullptr) (*g)[p] = linalg::Vector();
        rg[p] = linalg::Vector();
        bool c = inter_[i]->Constant(rg, g);
        if (g != nullptr) g->erase(p);
        rg.erase(p);
        if (c) return true;
      }
    }
    return true;
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rg, Gradient* g) override {  // fold.go:141-178
    if (pool_.empty()) return final_->bwdR(s, std::move(u), std::move(uR), rg, g);
    Gradient temp;
    if (g == nullptr) g = &temp;  // fold.go:147-149
    Dev stateUp = std::move(u), stateUpR = std::move(uR);
    for (size_t i = pool_.size(); i > 0; i--) {
      RResult& res = i == pool_.size() ? *final_ : *inter_[i];
      Variable* p = pool_[i - 1].get();
      (*g)[p] = linalg::Vector();
      rg[p] = linalg::Vector();
      if (res.Constant(rg, g)) {
        g->erase(p);
        rg.erase(p);
        break;
      }
      res.bwdR(s, stateUp, stateUpR, rg, g);
      stateUp = s.Take(*g, p, p->len());
      stateUpR = s.Take(rg, p, p->len());
    }
    inter_[0]->bwdR(s, stateUp, stateUpR, rg, g);
    if (g == &temp) s.Forget(temp);
  }

 private:
  std::vector<VariablePtr> pool_;
  std::vector<RResultPtr> inter_;
  RResultPtr final_;
};

class MatMulResult : public DeviceResult {
 public:
  MatMulResult(ResultPtr mat, VariablePtr matVar, ResultPtr res)
      : DeviceResult(res->dev()), mat_(std::move(mat)), matVar_(std::move(matVar)), res_(std::move(res)) {}
  bool Constant(const Gradient& g) override { return res_->Constant(g) && mat_->Constant(g); }  // lin_alg.go:53-55
  void bwd(Session& s, Dev u, Gradient& grad) override {                                        // lin_alg.go:57-69
    bool matConst = mat_->Constant(grad);
    if (!matConst) grad[matVar_.get()] = linalg::Vector();
    res_->bwd(s, std::move(u), grad);
    if (!matConst) mat_->bwd(s, s.Take(grad, matVar_.get(), matVar_->len()), grad);
  }

 private:
  ResultPtr mat_;
  VariablePtr matVar_;
  ResultPtr res_;
};

class MatMulRResult : public DeviceRResult {
 public:
  MatMulRResult(RResultPtr mat, VariablePtr matVar, RResultPtr res)
      : DeviceRResult(res->dev(), res->devR()), mat_(std::move(mat)), matVar_(std::move(matVar)), res_(std::move(res)) {}
  bool Constant(const RGradient& rg, const Gradient* g) override { return res_->Constant(rg, g) && mat_->Constant(rg, g); }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {  // lin_alg.go:113-133
    bool matConst = mat_->Constant(rgrad, grad);
    Gradient temp;
    if (!matConst) {
      if (grad == nullptr) grad = &temp;
      (*grad)[matVar_.get()] = linalg::Vector();
      rgrad[matVar_.get()] = linalg::Vector();
    }
    res_->bwdR(s, std::move(u), std::move(uR), rgrad, grad);
    if (!matConst) {
      Dev d = s.Take(*grad, matVar_.get(), matVar_->len()), dr = s.Take(rgrad, matVar_.get(), matVar_->len());
      mat_->bwdR(s, d, dr, rgrad, grad);
    }
    if (grad == &temp) s.Forget(temp);
  }

 private:
  RResultPtr mat_;
  VariablePtr matVar_;
  RResultPtr res_;
};

}  // namespace b200

// pool.go:38-68
inline ResultPtr PoolAll(std::vector<ResultPtr> ins, const std::function<ResultPtr(const std::vector<ResultPtr>&)>& f) {
  std::vector<VariablePtr> vars;
  std::vector<ResultPtr> pooled;
  for (auto& r : ins) {
    vars.push_back(b200::PoolVar(*r));
    pooled.push_back(vars.back());
  }
  ResultPtr out = f(pooled);
  return std::make_shared<b200::PooledResult>(std::move(ins), std::move(vars), std::move(out));
}
inline RResultPtr PoolAllR(std::vector<RResultPtr> ins,
                           const std::function<RResultPtr(const std::vector<RResultPtr>&)>& f) {
  std::vector<VariablePtr> vars;
  std::vector<RResultPtr> pooled;
  for (auto& r : ins) {
    RVariablePtr rv = b200::PoolRVar(*r);
    vars.push_back(rv->Var);
    pooled.push_back(rv);
  }
  RResultPtr out = f(pooled);
  return std::make_shared<b200::PooledRResult>(std::move(ins), std::move(vars), std::move(out));
}
// pool.go:21-32
inline ResultPtr Pool(ResultPtr r, const std::function<ResultPtr(ResultPtr)>& f) {
  return PoolAll({std::move(r)}, [&](const std::vector<ResultPtr>& in) { return f(in[0]); });
}
inline RResultPtr PoolR(RResultPtr r, const std::function<RResultPtr(RResultPtr)>& f) {
  return PoolAllR({std::move(r)}, [&](const std::vector<RResultPtr>& in) { return f(in[0]); });
}
// pool.go:70-91
inline ResultPtr PoolSplit(int n, ResultPtr r, const std::function<ResultPtr(const std::vector<ResultPtr>&)>& f) {
  return Pool(std::move(r), [&](ResultPtr in) { return f(Split(n, std::move(in))); });
}
inline RResultPtr PoolSplitR(int n, RResultPtr r, const std::function<RResultPtr(const std::vector<RResultPtr>&)>& f) {
  return PoolR(std::move(r), [&](RResultPtr in) { return f(SplitR(n, std::move(in))); });
}

// fold.go:19-30,93-107: the recurrent state never leaves HBM between timesteps
inline ResultPtr Fold(ResultPtr state, const std::vector<ResultPtr>& ins,
                      const std::function<ResultPtr(ResultPtr state, ResultPtr in)>& f) {
  std::vector<VariablePtr> pool;
  std::vector<ResultPtr> inter;
  for (auto& in : ins) {
    inter.push_back(state);
    pool.push_back(b200::PoolVar(*state));
    state = f(pool.back(), in);
  }
  return std::make_shared<b200::FoldResult>(std::move(pool), std::move(inter), std::move(state));
}
inline RResultPtr FoldR(RResultPtr state, const std::vector<RResultPtr>& ins,
                        const std::function<RResultPtr(RResultPtr state, RResultPtr in)>& f) {
  std::vector<VariablePtr> pool;
  std::vector<RResultPtr> inter;
  for (auto& in : ins) {
    inter.push_back(state);
    RVariablePtr pr = b200::PoolRVar(*state);
    pool.push_back(pr->Var);
    state = f(pr, in);
  }
  return std::make_shared<b200::FoldRResult>(std::move(pool), std::move(inter), std::move(state));
}

// lin_alg.go:14-133: LinTran whose matrix is itself a Result; reuses the six LinTran kernels verbatim
inline ResultPtr MatMulVecs(ResultPtr mat, int rows, int cols, ResultPtr vecs) {
  if (mat->len() != (int64_t)rows * cols) throw Panic("invalid matrix data size");  // lin_alg.go:29-31
  if (cols <= 0 || vecs->len() % cols != 0) throw Panic("invalid vecs size");        // lin_alg.go:32-34
  int n = (int)(vecs->len() / cols);
  VariablePtr v = b200::PoolVar(*mat);
  ResultPtr res = LinTran(v, rows, cols).Batch(std::move(vecs), n);
  return std::make_shared<b200::MatMulResult>(std::move(mat), std::move(v), std::move(res));
}
inline RResultPtr MatMulVecsR(RResultPtr mat, int rows, int cols, RResultPtr vecs) {
  if (mat->len() != (int64_t)rows * cols) throw Panic("invalid matrix data size");  // lin_alg.go:85-87
  if (cols <= 0 || vecs->len() % cols != 0) throw Panic("invalid vecs size");        // lin_alg.go:88-90
  int n = (int)(vecs->len() / cols);
  RVariablePtr mr = b200::PoolRVar(*mat);
  RResultPtr res = LinTran(mr->Var, rows, cols).BatchRWith(mr, std::move(vecs), n);  // RVector{v: mat.ROutput()}
  return std::make_shared<b200::MatMulRResult>(std::move(mat), mr->Var, std::move(res));
}
inline ResultPtr MatMulVec(ResultPtr mat, int rows, int cols, ResultPtr vec) {  // lin_alg.go:16-18
  return MatMulVecs(std::move(mat), rows, cols, std::move(vec));
}
inline RResultPtr MatMulVecR(RResultPtr mat, int rows, int cols, RResultPtr vec) {  // lin_alg.go:74-76
  return MatMulVecsR(std::move(mat), rows, cols, std::move(vec));
}

// =================================================================================================
// b200 fast paths: fused dense layer, whole-network MLP plan, LSTM plan
// =================================================================================================
namespace b200 {

// Drop-in for the triple {&LinTran{W}, &LinAdd{b}, Sigmoid{}} of bench/mlp.go:75-83: one kernel per direction.
class DenseSigmoid : public RFunc, public RBatcher {
 public:
  VariablePtr Weights, Biases;
  int Rows, Cols;
  bool Activation;

  DenseSigmoid(VariablePtr w, VariablePtr b, int rows, int cols, bool activation = true)
      : Weights(std::move(w)), Biases(std::move(b)), Rows(rows), Cols(cols), Activation(activation) {
    if (Weights->len() != (int64_t)rows * cols || Biases->len() != rows)
      throw Panic("parameter sizes do not match rows x cols");
  }
  ResultPtr Apply(ResultPtr in) override { return Batch(std::move(in), 1); }
  RResultPtr ApplyR(const RVector& v, RResultPtr in) override { return BatchR(v, std::move(in), 1); }
  ResultPtr Batch(ResultPtr in, int n) override;
  RResultPtr BatchR(const RVector& v, RResultPtr in, int n) override;

 private:
  void checkLen(int64_t got, int n) const {
    if ((int64_t)Cols * n != got)
      throw Panic(Sprintf("input length should be %lld but got %lld", (long long)Cols * n, got));
  }
};

class DenseResult : public DeviceResult {
 public:
  DenseResult(DenseSigmoid l, ResultPtr in, Dev x, int n, Dev s)
      : DeviceResult(std::move(s)), l_(std::move(l)), in_(std::move(in)), x_(std::move(x)), n_(n) {}
  bool Constant(const Gradient& g) override {
    return in_->Constant(g) && l_.Weights->Constant(g) && l_.Biases->Constant(g);
  }
  void bwd(Session& s, Dev u, Gradient& grad) override { bwdFrom(s, std::move(u), grad, false); }
  // skipAct: the layer above already applied this layer's sigmoid backward in its input-gradient epilogue
  void bwdFrom(Session& s, Dev u, Gradient& grad, bool skipAct) {
    if (Constant(grad)) return;
    if (l_.Activation && !skipAct) check(af_sigmoid_bwd(ctx(), d_->p, u->p, u->n));
    if (Dev gb = AccOrNull(s, &grad, l_.Biases.get())) check(af_bias_grad(ctx(), u->p, gb->p, l_.Rows, n_));
    if (Dev gw = AccOrNull(s, &grad, l_.Weights.get()))
      check(af_lintran_wgrad(ctx(), u->p, x_->p, gw->p, l_.Rows, l_.Cols, n_));
    if (in_->Constant(grad)) return;
    auto* prev = dynamic_cast<DenseResult*>(in_.get());
    bool fuse = prev != nullptr && prev->l_.Activation;
    Dev d = DevVec::Empty((int64_t)n_ * l_.Cols);
    check(af_dense_igrad(ctx(), u->p, nullptr, l_.Weights->dev()->p, nullptr, fuse ? prev->d_->p : nullptr, nullptr,
                         d->p, nullptr, l_.Rows, l_.Cols, n_));
    if (fuse) prev->bwdFrom(s, d, grad, true);
    else in_->bwd(s, d, grad);
  }

 private:
  DenseSigmoid l_;
  ResultPtr in_;
  Dev x_;
  int n_;
};

class DenseRResult : public DeviceRResult {
 public:
  DenseRResult(DenseSigmoid l, RResultPtr in, RVariablePtr rw, RVariablePtr rb, Dev x, Dev xr, int n, Dev s, Dev rs)
      : DeviceRResult(std::move(s), std::move(rs)), l_(std::move(l)), in_(std::move(in)), rw_(std::move(rw)),
        rb_(std::move(rb)), x_(std::move(x)), xr_(std::move(xr)), n_(n) {}
  bool Constant(const RGradient& rg, const Gradient* g) override {
    return in_->Constant(rg, g) && rw_->Constant(rg, g) && rb_->Constant(rg, g);
  }
  void bwdR(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad) override {
    bwdRFrom(s, std::move(u), std::move(uR), rgrad, grad, false);
  }
  void bwdRFrom(Session& s, Dev u, Dev uR, RGradient& rgrad, Gradient* grad, bool skipAct) {
    if (Constant(rgrad, grad)) return;
    if (l_.Activation && !skipAct)  // math_funcs.go:357-374, in place
      check(af_sigmoid_bwd_r(ctx(), d_->p, dr_->p, u->p, uR->p, u->n));
    Variable *b = l_.Biases.get(), *w = l_.Weights.get();
    Dev gb = AccOrNull(s, grad, b), rgb = AccOrNull(s, &rgrad, b);
    if (gb || rgb)  // lin_add.go:94-104, batched
      check(af_bias_grad_r(ctx(), u->p, uR->p, mptr(gb), mptr(rgb), l_.Rows, n_));
    Dev gw = AccOrNull(s, grad, w), rgw = AccOrNull(s, &rgrad, w);
    if (gw || rgw)  // lin_tran.go:410-418
      check(af_lintran_wgrad_r(ctx(), u->p, uR->p, x_->p, ptr(xr_), mptr(gw), mptr(rgw), l_.Rows, l_.Cols, n_));
    if (in_->Constant(rgrad, grad)) return;
    // lin_tran.go:420-423 (+ the sigmoid backward of the layer below, fused into the epilogue)
    auto* prev = dynamic_cast<DenseRResult*>(in_.get());
    bool fuse = prev != nullptr && prev->l_.Activation;
    Dev d = DevVec::Empty((int64_t)n_ * l_.Cols), dr = DevVec::Empty((int64_t)n_ * l_.Cols);
    check(af_dense_igrad(ctx(), u->p, uR->p, rw_->dev()->p, ptr(rw_->devR()), fuse ? prev->d_->p : nullptr,
                         fuse ? prev->dr_->p : nullptr, d->p, dr->p, l_.Rows, l_.Cols, n_));
    if (fuse) prev->bwdRFrom(s, d, dr, rgrad, grad, true);
    else in_->bwdR(s, d, dr, rgrad, grad);
  }

 private:
  DenseSigmoid l_;
  RResultPtr in_;
  RVariablePtr rw_, rb_;
  Dev x_, xr_;
  int n_;
};

inline ResultPtr DenseSigmoid::Batch(ResultPtr in, int n) {
  checkLen(in->len(), n);
  Dev x = in->dev();
  Dev s = DevVec::Empty((int64_t)n * Rows);
  check(af_dense_fwd(ctx(), Weights->dev()->p, nullptr, Biases->dev()->p, nullptr, x->p, nullptr, s->p, nullptr, Rows,
                     Cols, n, Activation ? 1 : 0));
  return std::make_shared<DenseResult>(*this, in, x, n, s);
}

inline RResultPtr DenseSigmoid::BatchR(const RVector& v, RResultPtr in, int n) {
  checkLen(in->len(), n);
  RVariablePtr rw = NewRVariable(Weights, v), rb = NewRVariable(Biases, v);
  Dev x = in->dev(), xr = in->devR();
  Dev s = DevVec::Empty((int64_t)n * Rows), rs = DevVec::Empty((int64_t)n * Rows);
  check(af_dense_fwd(ctx(), rw->dev()->p, ptr(rw->devR()), rb->dev()->p, ptr(rb->devR()), x->p, ptr(xr), s->p, rs->p,
                     Rows, Cols, n, Activation ? 1 : 0));
  return std::make_shared<DenseRResult>(*this, in, rw, rb, x, xr, n, s, rs);
}

// Rewrite {LinTran, LinAdd, Sigmoid} runs of a composed batcher into DenseSigmoid layers.
inline ComposedRBatcher FuseDense(const ComposedRBatcher& layers) {
  ComposedRBatcher out;
  size_t i = 0;
  while (i < layers.size()) {
    auto* lt = dynamic_cast<LinTran*>(layers[i].get());
    auto* la = i + 1 < layers.size() ? dynamic_cast<LinAdd*>(layers[i + 1].get()) : nullptr;
    if (lt && la && la->Var->len() == lt->Rows) {
      bool act = i + 2 < layers.size() && dynamic_cast<Sigmoid*>(layers[i + 2].get()) != nullptr;
      out.push_back(std::make_shared<DenseSigmoid>(lt->Data, la->Var, lt->Rows, lt->Cols, act));
      i += act ? 3 : 2;
    } else {
      out.push_back(layers[i]);
      i += 1;
    }
  }
  return out;
}

// The whole-network plan (bench/mlp.go:24-126, batched): parameters, R-vectors, activations and accumulators stay
// resident in HBM; one call per iteration.
class MLP {
 public:
  std::vector<int64_t> Sizes;
  int64_t Batch;

  MLP(std::vector<int64_t> sizes, int64_t batch) : Sizes(std::move(sizes)), Batch(batch) {
    check(af_mlp_create(ctx(), Sizes.data(), (int)Sizes.size() - 1, batch, &h_));
  }
  MLP(const MLP&) = delete;
  MLP& operator=(const MLP&) = delete;
  ~MLP() {
    if (h_) af_mlp_destroy(h_);
  }
  af_mlp* Handle() { return h_; }
  int NumParams() const { return 2 * ((int)Sizes.size() - 1); }
  int64_t ParamLen(int idx) const {
    int l = idx / 2;
    return (idx & 1) ? Sizes[l + 1] : Sizes[l] * Sizes[l + 1];
  }
  // vars in bench/mlp.go:84-85 order (W0, b0, W1, b1, ...); a variable absent from rv has a zero R-vector.
  void SetParams(const std::vector<VariablePtr>& vars, const RVector& rv) {
    if ((int)vars.size() != NumParams()) throw Panic("parameter count does not match the network", AF_EINVAL);
    for (int i = 0; i < NumParams(); i++) {
      check(af_mlp_set_param(h_, i, vars[i]->Vector.data(), (int64_t)vars[i]->Vector.size()));
      auto it = rv.find(vars[i].get());
      if (it != rv.end()) check(af_mlp_set_rvector(h_, i, it->second.data(), (int64_t)it->second.size()));
      else check(af_mlp_set_rvector(h_, i, nullptr, 0));
    }
  }
  // One iteration of bench/mlp.go:52-54 from HOST buffers: gradients are ADDED (variable.go:45) into grad / rgrad
  // for the variables present as keys (grad may be nullptr).  outGrad / outRGrad empty => ones / zeros
  // (bench/mlp.go:118-126).  Returns Output and ROutput.
  std::pair<linalg::Vector, linalg::Vector> StepR(const std::vector<VariablePtr>& vars, const linalg::Vector& input,
                                                  const linalg::Vector& outGrad, const linalg::Vector& outRGrad,
                                                  RGradient& rgrad, Gradient* grad) {
    int64_t nOut = Batch * Sizes.back();
    if ((int64_t)input.size() != Batch * Sizes[0])
      throw Panic(Sprintf("input length should be %lld but got %lld", Batch * Sizes[0], (long long)input.size()));
    if ((!outGrad.empty() && (int64_t)outGrad.size() != nOut) || (!outRGrad.empty() && (int64_t)outRGrad.size() != nOut))
      throw Panic("invalid upstream size");
    linalg::Vector out((size_t)nOut), outR((size_t)nOut);
    int P = NumParams();
    std::vector<linalg::Vector> tg(P), trg(P);
    std::vector<double*> pg(P, nullptr), prg(P, nullptr);
    for (int i = 0; i < P; i++) {
      if (InGrad(grad, vars[i].get())) {
        tg[i].resize((size_t)ParamLen(i));
        pg[i] = tg[i].data();
      }
      if (InGrad(&rgrad, vars[i].get())) {
        trg[i].resize((size_t)ParamLen(i));
        prg[i] = trg[i].data();
      }
    }
    check(af_mlp_step_host(h_, 1, input.data(), outGrad.empty() ? nullptr : outGrad.data(),
                           outRGrad.empty() ? nullptr : outRGrad.data(), out.data(), outR.data(), pg.data(),
                           prg.data()));
    for (int i = 0; i < P; i++) {
      if (pg[i]) {
        linalg::Vector& d = (*grad)[vars[i].get()];
        for (size_t k = 0; k < d.size(); k++) d[k] += tg[i][k];
      }
      if (prg[i]) {
        linalg::Vector& d = rgrad[vars[i].get()];
        for (size_t k = 0; k < d.size(); k++) d[k] += trg[i][k];
      }
    }
    return {std::move(out), std::move(outR)};
  }
  // Resident form: one bench iteration with fresh accumulators, nothing crosses PCIe.
  void StepFresh(bool withR) { check(af_mlp_step_fresh(h_, withR ? 1 : 0)); }
  // Data parallel (Gradient.Add across ranks, gradient.go:28-30; the context joined a communicator with b200::DPInit):
  // SetDP picks the schedule of the reduce (AF_DP_*) for the host-buffer steps; the *DP forms always reduce.
  void SetDP(int mode) { check(af_mlp_set_dp(h_, mode)); }
  void StepFreshDP(bool withR) { check(af_mlp_step_fresh_dp(h_, withR ? 1 : 0)); }
  void StepDP(bool withR) { check(af_mlp_step_dp(h_, withR ? 1 : 0)); }  // accumulators zeroed by the caller
  void DPJoin() { check(af_mlp_dp_join(h_)); }
  void SetHostShard(bool on) { check(af_mlp_set_host_shard(h_, on ? 1 : 0)); }
  void AddToVars(double scale) { check(af_mlp_add_to_vars(h_, scale)); }  // gradient.go:40-52 in HBM
  Dev Input() { return DevVec::Wrap(af_mlp_input_ptr(h_), Batch * Sizes[0]); }
  Dev Output() { return DevVec::Wrap(af_mlp_output_ptr(h_), Batch * Sizes.back()); }
  Dev ROutput() { return DevVec::Wrap(af_mlp_routput_ptr(h_), Batch * Sizes.back()); }
  Dev Grad(int idx) { return DevVec::Wrap(af_mlp_grad_ptr(h_, idx), ParamLen(idx)); }
  Dev RGrad(int idx) { return DevVec::Wrap(af_mlp_rgrad_ptr(h_, idx), ParamLen(idx)); }

 private:
  af_mlp* h_ = nullptr;
};

// The LSTM plan (bench/lstm.go:25-241 batched over `lanes` sequences, R-operator carried through).  Parameter order
// is the reference's (bench/lstm.go:75-99): InWeights, InBiases, InGateWeights, InGateBiases, ForgetGateWeights,
// ForgetGateBiases, OutputGate, OutputGateBiases, OutputWeights, OutputBiases.
class LSTM {
 public:
  int64_t InputSize, HiddenSize, OutputSize, TimeSteps, Lanes;
  static constexpr int kNumParams = 10;

  LSTM(int64_t in, int64_t hidden, int64_t out, int64_t steps, int64_t lanes)
      : InputSize(in), HiddenSize(hidden), OutputSize(out), TimeSteps(steps), Lanes(lanes) {
    check(af_lstm_create(ctx(), in, hidden, out, steps, lanes, &h_));
  }
  LSTM(const LSTM&) = delete;
  LSTM& operator=(const LSTM&) = delete;
  ~LSTM() {
    if (h_) af_lstm_destroy(h_);
  }
  int64_t ParamLen(int idx) const {
    int64_t rows = idx < 8 ? HiddenSize : OutputSize;
    return (idx & 1) ? rows : rows * (HiddenSize + InputSize);
  }
  void SetParams(const std::vector<VariablePtr>& vars, const RVector& rv) {
    if ((int)vars.size() != kNumParams) throw Panic("parameter count does not match the network", AF_EINVAL);
    for (int i = 0; i < kNumParams; i++) {
      check(af_lstm_set_param(h_, i, vars[i]->Vector.data(), (int64_t)vars[i]->Vector.size()));
      auto it = rv.find(vars[i].get());
      if (it != rv.end()) check(af_lstm_set_rvector(h_, i, it->second.data(), (int64_t)it->second.size()));
      else check(af_lstm_set_rvector(h_, i, nullptr, 0));
    }
  }
  // Forward over all time steps, then back-propagation through time with the R-operator (bench/lstm.go:42-71),
  // from HOST buffers laid out T x lanes x size.  Gradients are ADDED into the maps for the keys present.
  std::pair<linalg::Vector, linalg::Vector> RunR(const std::vector<VariablePtr>& vars, const linalg::Vector& inputs,
                                                 const linalg::Vector& upstreams, const linalg::Vector& upstreamsR,
                                                 RGradient& rgrad, Gradient* grad) {
    int64_t nOut = TimeSteps * Lanes * OutputSize;
    if ((int64_t)inputs.size() != TimeSteps * Lanes * InputSize)
      throw Panic(Sprintf("input length should be %lld but got %lld", TimeSteps * Lanes * InputSize,
                          (long long)inputs.size()));
    if ((int64_t)upstreams.size() != nOut || (int64_t)upstreamsR.size() != nOut) throw Panic("invalid upstream size");
    linalg::Vector out((size_t)nOut), outR((size_t)nOut);
    std::vector<linalg::Vector> tg(kNumParams), trg(kNumParams);
    std::vector<double*> pg(kNumParams, nullptr), prg(kNumParams, nullptr);
    for (int i = 0; i < kNumParams; i++) {
      if (InGrad(grad, vars[i].get())) {
        tg[i].resize((size_t)ParamLen(i));
        pg[i] = tg[i].data();
      }
      if (InGrad(&rgrad, vars[i].get())) {
        trg[i].resize((size_t)ParamLen(i));
        prg[i] = trg[i].data();
      }
    }
    check(af_lstm_run_host(h_, 1, inputs.data(), upstreams.data(), upstreamsR.data(), out.data(), outR.data(),
                           pg.data(), prg.data()));
    for (int i = 0; i < kNumParams; i++) {
      if (pg[i]) {
        linalg::Vector& d = (*grad)[vars[i].get()];
        for (size_t k = 0; k < d.size(); k++) d[k] += tg[i][k];
      }
      if (prg[i]) {
        linalg::Vector& d = rgrad[vars[i].get()];
        for (size_t k = 0; k < d.size(); k++) d[k] += trg[i][k];
      }
    }
    return {std::move(out), std::move(outR)};
  }
  // Data parallel over the lanes: fresh accumulators, forward, BPTT on resident inputs / upstreams, then the sum of
  // the ranks' {grad | rgrad} blocks (af_lstm_step_dp).
  void StepDP(bool withR) { check(af_lstm_step_dp(h_, withR ? 1 : 0)); }

 private:
  af_lstm* h_ = nullptr;
};

}  // namespace b200

}  // namespace autofunc

#endif  // AUTOFUNC_B200_HPP_
