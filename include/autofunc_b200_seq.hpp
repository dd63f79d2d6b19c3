// autofunc_b200_seq.hpp -- seqfunc.MapBatcher / MapRBatcher (seqfunc/map_batcher.go:14-140) and the sequence-list
// Results they consume (seqfunc/seqfunc.go:10-53, construct.go), in C++ over include/autofunc_b200.hpp.
//
// A list of vector sequences is `Seqs` = [][]linalg.Vector.  A MapBatcher feeds its Batcher all the vectors of one
// timestep at once (joinTime, seqfunc/map.go:262-270), so T timesteps over B lanes are T batched device calls instead
// of the T*B single-vector calls of MapFunc.  Here the list is packed time-major in ONE HBM buffer (PackedSeqs), so a
// timestep is a view, mapping moves no data, and the pool-variable protocol of the backward pass (insert
// grad[poolVar], propagate, read, delete -- map_batcher.go:50-61,124-139) runs on one device session for all
// timesteps.  Host-level sequence Results (ConstResult, VarResult, any user type) still work on either side: they
// are packed / unpacked at the boundary.  The resident form of a whole recurrent network is b200::LSTM.
#ifndef AUTOFUNC_B200_SEQ_HPP_
#define AUTOFUNC_B200_SEQ_HPP_

#include "autofunc_b200.hpp"

namespace autofunc {
namespace seqfunc {

using Seqs = std::vector<std::vector<linalg::Vector>>;

class PackedSeqs;
using Packed = std::shared_ptr<PackedSeqs>;

// The lower-case members are the device protocol (as in autofunc::Result): a sequence list packed time-major in ONE
// HBM buffer.  Their defaults make any host-only seqfunc Result usable as the input of a device map.
class Result {  // seqfunc/seqfunc.go:24-35
 public:
  virtual ~Result() = default;
  virtual const Seqs& OutputSeqs() = 0;
  virtual void PropagateGradient(const Seqs& upstream, Gradient& g) = 0;  // must not modify the upstream
  virtual Packed packed();
  virtual void bwdPacked(b200::Session& s, const PackedSeqs& upstream, Gradient& g);
};
using ResultPtr = std::shared_ptr<Result>;

class RResult {  // seqfunc/seqfunc.go:39-53
 public:
  virtual ~RResult() = default;
  virtual const Seqs& OutputSeqs() = 0;
  virtual const Seqs& ROutputSeqs() = 0;
  virtual void PropagateRGradient(const Seqs& upstream, const Seqs& upstreamR, RGradient& rg, Gradient* g) = 0;
  virtual Packed packed();
  virtual Packed packedR();  // nullptr == identically zero
  virtual void bwdPackedR(b200::Session& s, const PackedSeqs& upstream, const PackedSeqs& upstreamR, RGradient& rg,
                          Gradient* g);
};
using RResultPtr = std::shared_ptr<RResult>;

class Func {  // seqfunc/seqfunc.go:12-14
 public:
  virtual ~Func() = default;
  virtual ResultPtr ApplySeqs(ResultPtr r) = 0;
};
class RFunc : public virtual Func {  // seqfunc/seqfunc.go:17-20
 public:
  virtual RResultPtr ApplySeqsR(const RVector& rv, RResultPtr r) = 0;
};

// ---- seqfunc/map.go:245-284 -------------------------------------------------------------------------------------
inline size_t maxSequenceLen(const Seqs& s) {
  size_t m = 0;
  for (auto& x : s) m = x.size() > m ? x.size() : m;
  return m;
}
inline linalg::Vector joinTime(const Seqs& seqs, size_t t, int* n) {
  linalg::Vector joined;
  *n = 0;
  for (auto& s : seqs)
    if (s.size() > t) {
      joined.insert(joined.end(), s[t].begin(), s[t].end());
      ++*n;
    }
  return joined;
}
inline void appendSplit(const Seqs& seqs, size_t t, Seqs& dest, const linalg::Vector& joined) {
  int n = 0;
  for (auto& s : seqs) n += s.size() > t;
  if (n == 0) return;
  if (joined.size() % (size_t)n != 0) throw Panic("cannot evenly split vector");
  size_t each = joined.size() / (size_t)n, at = 0;
  for (size_t lane = 0; lane < seqs.size(); lane++)
    if (seqs[lane].size() > t) {
      dest[lane].emplace_back(joined.begin() + (long)at, joined.begin() + (long)(at + each));
      at += each;
    }
}

// ---- device layout ----------------------------------------------------------------------------
// A `[][]linalg.Vector` list laid out the way the batcher reads it: timestep-major, and within a timestep the vectors
// of the lanes that are still alive, in lane order -- exactly the vector joinTime builds on the host for every step
// (seqfunc/map.go:262-270).  Packing happens once (one upload); after that step t is a VIEW of the buffer, so mapping
// a Batcher over the list moves no data, and splitting back (appendSplit, map.go:272-284) waits until a host caller
// asks for the lists.
class PackedSeqs {
 public:
  std::vector<size_t> lens;     // lane lengths
  std::vector<int> counts;      // live lanes per step
  std::vector<int64_t> widths;  // vector size per step
  std::vector<int64_t> offsets; // offsets[t] .. offsets[t+1] = step t
  b200::Dev data;

  PackedSeqs(std::vector<size_t> lens_, std::vector<int64_t> widths_, b200::Dev data_)
      : lens(std::move(lens_)), widths(std::move(widths_)), data(std::move(data_)) {
    size_t T = 0;
    for (size_t n : lens) T = n > T ? n : T;
    if (widths.size() != T) throw Panic("packed buffer does not match the lane lengths");
    offsets.push_back(0);
    for (size_t t = 0; t < T; t++) {
      int c = 0;
      for (size_t n : lens) c += n > t;
      counts.push_back(c);
      offsets.push_back(offsets.back() + c * widths[t]);
    }
    if (!data || data->n != offsets.back()) throw Panic("packed buffer does not match the lane lengths");
  }
  size_t T() const { return counts.size(); }
  static void PackHost(const Seqs& seqs, std::vector<size_t>* lens, std::vector<int64_t>* widths, linalg::Vector* flat) {
    lens->clear(), widths->clear(), flat->clear();
    for (auto& sq : seqs) lens->push_back(sq.size());
    size_t T = maxSequenceLen(seqs);
    for (size_t t = 0; t < T; t++) {
      int n;
      linalg::Vector joined = joinTime(seqs, t, &n);
      if (joined.size() % (size_t)n != 0) throw Panic("cannot evenly split vector");
      size_t w = joined.size() / (size_t)n;
      for (auto& sq : seqs)
        if (sq.size() > t && sq[t].size() != w) throw Panic("cannot evenly split vector");
      widths->push_back((int64_t)w);
      flat->insert(flat->end(), joined.begin(), joined.end());
    }
  }
  static Packed FromHost(const Seqs& seqs) {
    std::vector<size_t> lens;
    std::vector<int64_t> widths;
    linalg::Vector flat;
    PackHost(seqs, &lens, &widths, &flat);
    return std::make_shared<PackedSeqs>(std::move(lens), std::move(widths), b200::DevVec::Upload(flat));
  }
  Packed Like(std::vector<int64_t> w, b200::Dev d) const { return std::make_shared<PackedSeqs>(lens, std::move(w), std::move(d)); }
  b200::Dev Step(size_t t) const { return b200::DevVec::View(data, offsets[t], offsets[t + 1]); }
  bool SameShape(const PackedSeqs& o) const { return lens == o.lens && widths == o.widths; }
  Seqs ToHost() const {
    linalg::Vector flat = data->Download();
    Seqs out(lens.size());
    for (size_t t = 0; t < T(); t++) {
      int64_t at = offsets[t];
      for (size_t lane = 0; lane < lens.size(); lane++)
        if (lens[lane] > t) {
          out[lane].emplace_back(flat.begin() + at, flat.begin() + at + widths[t]);
          at += widths[t];
        }
    }
    return out;
  }
};

inline Packed Result::packed() { return PackedSeqs::FromHost(OutputSeqs()); }
inline void Result::bwdPacked(b200::Session& s, const PackedSeqs& u, Gradient& g) {
  s.Flush();  // a host-only Result: what the device nodes accumulated so far must be in the maps it sees
  PropagateGradient(u.ToHost(), g);
}
inline Packed RResult::packed() { return PackedSeqs::FromHost(OutputSeqs()); }
inline Packed RResult::packedR() { return PackedSeqs::FromHost(ROutputSeqs()); }
inline void RResult::bwdPackedR(b200::Session& s, const PackedSeqs& u, const PackedSeqs& uR, RGradient& rg, Gradient* g) {
  s.Flush();
  PropagateRGradient(u.ToHost(), uR.ToHost(), rg, g);
}
// out.PropagateGradient / out.PropagateRGradient with packed upstreams, on the host maps ...
inline void PropagateGradient(Result& out, const PackedSeqs& u, Gradient& g) {
  b200::Session s;
  out.bwdPacked(s, u, g);
  s.Flush();
}
inline void PropagateRGradient(RResult& out, const PackedSeqs& u, const PackedSeqs& uR, RGradient& rg, Gradient* g) {
  b200::Session s;
  out.bwdPackedR(s, u, uR, rg, g);
  s.Flush();
}
// ... and with the maps in HBM (b200::DeviceGradient): a sequence backward pass that never visits the host
inline void PropagateGradient(Result& out, const PackedSeqs& u, b200::DeviceGradient& grad) {
  b200::Session s;
  grad.SeedInto(s);
  out.bwdPacked(s, u, grad.Keys());
  s.Flush();
}
inline void PropagateRGradient(RResult& out, const PackedSeqs& u, const PackedSeqs& uR, b200::DeviceGradient& rgrad,
                               b200::DeviceGradient* grad) {
  b200::Session s;
  rgrad.SeedInto(s);
  if (grad) grad->SeedInto(s);
  out.bwdPackedR(s, u, uR, rgrad.Keys(), grad ? &grad->Keys() : nullptr);
  s.Flush();
}

// ---- seqfunc/construct.go ---------------------------------------------------------------------------------------
class ConstResult : public Result {
 public:
  explicit ConstResult(Seqs v) : out_(std::move(v)) {}
  const Seqs& OutputSeqs() override { return out_; }
  void PropagateGradient(const Seqs&, Gradient&) override {}
  void bwdPacked(b200::Session&, const PackedSeqs&, Gradient&) override {}  // a gradient sink: nothing to download

 private:
  Seqs out_;
};
class ConstRResult : public RResult {  // zero r-outputs (construct.go:33-44)
 public:
  explicit ConstRResult(Seqs v) : out_(std::move(v)), rout_(out_) {
    for (auto& s : rout_)
      for (auto& x : s) std::fill(x.begin(), x.end(), 0.0);
  }
  const Seqs& OutputSeqs() override { return out_; }
  const Seqs& ROutputSeqs() override { return rout_; }
  void PropagateRGradient(const Seqs&, const Seqs&, RGradient&, Gradient*) override {}
  Packed packedR() override { return nullptr; }  // identically zero: the Batcher skips the products
  void bwdPackedR(b200::Session&, const PackedSeqs&, const PackedSeqs&, RGradient&, Gradient*) override {}

 private:
  Seqs out_, rout_;
};
using VarSeqs = std::vector<std::vector<VariablePtr>>;
class VarResult : public Result {  // construct.go:63-103
 public:
  explicit VarResult(VarSeqs vars) : vars_(std::move(vars)) {
    for (auto& seq : vars_) {
      out_.emplace_back();
      for (auto& v : seq) out_.back().push_back(v->Vector);
    }
  }
  const Seqs& OutputSeqs() override { return out_; }
  void PropagateGradient(const Seqs& u, Gradient& g) override {
    for (size_t i = 0; i < vars_.size(); i++)
      for (size_t j = 0; j < vars_[i].size(); j++) vars_[i][j]->PropagateGradient(u.at(i).at(j), g);
  }

 private:
  VarSeqs vars_;
  Seqs out_;
};
class VarRResult : public RResult {  // construct.go:105-160
 public:
  VarRResult(const RVector& rv, VarSeqs vars) : vars_(std::move(vars)) {
    for (auto& seq : vars_) {
      out_.emplace_back();
      rout_.emplace_back();
      rvars_.emplace_back();
      for (auto& v : seq) {
        rvars_.back().push_back(NewRVariable(v, rv));
        out_.back().push_back(v->Vector);
        rout_.back().push_back(rvars_.back().back()->ROutput());
      }
    }
  }
  const Seqs& OutputSeqs() override { return out_; }
  const Seqs& ROutputSeqs() override { return rout_; }
  void PropagateRGradient(const Seqs& u, const Seqs& uR, RGradient& rg, Gradient* g) override {
    for (size_t i = 0; i < vars_.size(); i++)
      for (size_t j = 0; j < vars_[i].size(); j++) rvars_[i][j]->PropagateRGradient(u.at(i).at(j), uR.at(i).at(j), rg, g);
  }

 private:
  VarSeqs vars_;
  std::vector<std::vector<RVariablePtr>> rvars_;
  Seqs out_, rout_;
};

// seqfunc.VarResult for a list that is already packed: ONE Variable holds the whole time-major buffer, so the sequence
// gradient arrives as one vector with the same layout and (with a b200::DeviceGradient) stays in HBM.
class PackedVarResult : public Result {
 public:
  PackedVarResult(VariablePtr var, std::vector<size_t> lens, std::vector<int64_t> widths)
      : var_(std::move(var)), p_(std::make_shared<PackedSeqs>(std::move(lens), std::move(widths), var_->dev())) {}
  const Seqs& OutputSeqs() override {
    if (out_.empty()) out_ = p_->ToHost();
    return out_;
  }
  Packed packed() override { return p_; }
  void PropagateGradient(const Seqs& u, Gradient& g) override { seqfunc::PropagateGradient(*this, *PackedSeqs::FromHost(u), g); }
  void bwdPacked(b200::Session& s, const PackedSeqs& u, Gradient& g) override {
    if (!p_->SameShape(u)) throw Panic("upstream sequence shape mismatch");
    var_->bwd(s, u.data, g);
  }

 private:
  VariablePtr var_;
  Packed p_;
  Seqs out_;
};
class PackedVarRResult : public RResult {
 public:
  PackedVarRResult(const RVector& rv, VariablePtr var, std::vector<size_t> lens, std::vector<int64_t> widths)
      : rvar_(NewRVariable(std::move(var), rv)),
        p_(std::make_shared<PackedSeqs>(std::move(lens), std::move(widths), rvar_->dev())) {}
  const Seqs& OutputSeqs() override {
    if (out_.empty()) out_ = p_->ToHost();
    return out_;
  }
  const Seqs& ROutputSeqs() override {
    if (rout_.empty()) {
      Packed r = packedR();
      rout_ = r ? r->ToHost() : p_->Like(p_->widths, b200::DevVec::Zeros(p_->data->n))->ToHost();
    }
    return rout_;
  }
  Packed packed() override { return p_; }
  Packed packedR() override {
    b200::Dev r = rvar_->devR();
    return r ? p_->Like(p_->widths, r) : nullptr;
  }
  void PropagateRGradient(const Seqs& u, const Seqs& uR, RGradient& rg, Gradient* g) override {
    seqfunc::PropagateRGradient(*this, *PackedSeqs::FromHost(u), *PackedSeqs::FromHost(uR), rg, g);
  }
  void bwdPackedR(b200::Session& s, const PackedSeqs& u, const PackedSeqs& uR, RGradient& rg, Gradient* g) override {
    if (!p_->SameShape(u) || !p_->SameShape(uR)) throw Panic("upstream sequence shape mismatch");
    rvar_->bwdR(s, u.data, uR.data, rg, g);
  }

 private:
  RVariablePtr rvar_;
  Packed p_;
  Seqs out_, rout_;
};

// ---- seqfunc/map_batcher.go ---------------------------------------------------------------------------------------
// The input list is packed once (Result::packed: one upload for a host list, nothing for a device one), timestep t is a
// view of that buffer wrapped as the pool variable of map_batcher.go:28 / :85-88, the Batcher's outputs are gathered
// into one packed output buffer, and the backward pass runs the insert / propagate / read / delete protocol of
// :50-61,124-139 for ALL timesteps on one device session: parameter gradients accumulate in HBM across the whole
// sequence and reach the host maps once.
class MapBatcherResult : public Result {
 public:
  ResultPtr Input;
  Packed Shape, Out;
  std::vector<VariablePtr> Pool;
  std::vector<autofunc::ResultPtr> Results;
  const Seqs& OutputSeqs() override {
    if (!have_) host_ = Out->ToHost(), have_ = true;
    return host_;
  }
  Packed packed() override { return Out; }
  void PropagateGradient(const Seqs& u, Gradient& g) override {  // :50-63
    seqfunc::PropagateGradient(*this, *PackedSeqs::FromHost(u), g);
  }
  void bwdPacked(b200::Session& s, const PackedSeqs& u, Gradient& g) override {
    if (!Out->SameShape(u)) throw Panic("upstream sequence shape mismatch");
    b200::Dev down = b200::DevVec::Empty(Shape->data->n);
    for (size_t t = 0; t < Shape->T(); t++) {
      Variable* poolVar = Pool[t].get();
      g[poolVar] = linalg::Vector();
      Results[t]->bwd(s, u.Step(t), g);
      b200::Dev d = s.Take(g, poolVar, poolVar->len());
      if (d->n > 0) b200::check(af_copy(b200::ctx(), down->p + Shape->offsets[t], d->p, d->n));
    }
    Input->bwdPacked(s, *Shape->Like(Shape->widths, down), g);
  }

 private:
  Seqs host_;
  bool have_ = false;
};

class MapBatcher : public virtual Func {
 public:
  std::shared_ptr<autofunc::Batcher> B;
  explicit MapBatcher(std::shared_ptr<autofunc::Batcher> b) : B(std::move(b)) {}
  ResultPtr ApplySeqs(ResultPtr r) override {  // map_batcher.go:20-38
    auto res = std::make_shared<MapBatcherResult>();
    res->Input = r;
    res->Shape = r->packed();
    const PackedSeqs& in = *res->Shape;
    std::vector<int64_t> widths;
    int64_t total = 0;
    for (size_t t = 0; t < in.T(); t++) {
      VariablePtr inVar = Variable::Wrap(in.Step(t));
      autofunc::ResultPtr output = B->Batch(inVar, in.counts[t]);
      if (output->len() % in.counts[t] != 0) throw Panic("cannot evenly split vector");  // map.go:277-279
      widths.push_back(output->len() / in.counts[t]);
      total += output->len();
      res->Pool.push_back(inVar);
      res->Results.push_back(output);
    }
    b200::Dev out = b200::DevVec::Empty(total);
    int64_t at = 0;
    for (auto& o : res->Results) {
      b200::Dev d = o->dev();
      if (d->n > 0) b200::check(af_copy(b200::ctx(), out->p + at, d->p, d->n));
      at += d->n;
    }
    res->Out = in.Like(std::move(widths), out);
    return res;
  }
};

class MapBatcherRResult : public RResult {
 public:
  RResultPtr Input;
  Packed Shape, Out, OutR;
  std::vector<VariablePtr> Pool;
  std::vector<autofunc::RResultPtr> Results;
  const Seqs& OutputSeqs() override {
    if (!have_) host_ = Out->ToHost(), have_ = true;
    return host_;
  }
  const Seqs& ROutputSeqs() override {
    if (!have_r_) host_r_ = OutR->ToHost(), have_r_ = true;
    return host_r_;
  }
  Packed packed() override { return Out; }
  Packed packedR() override { return OutR; }
  void PropagateRGradient(const Seqs& u, const Seqs& uR, RGradient& rg, Gradient* g) override {  // :118-140
    seqfunc::PropagateRGradient(*this, *PackedSeqs::FromHost(u), *PackedSeqs::FromHost(uR), rg, g);
  }
  void bwdPackedR(b200::Session& s, const PackedSeqs& u, const PackedSeqs& uR, RGradient& rg, Gradient* g) override {
    if (!Out->SameShape(u) || !Out->SameShape(uR)) throw Panic("upstream sequence shape mismatch");
    Gradient temp;  // "We use g for temporary gradients." (map_batcher.go:120-123)
    Gradient* gg = g == nullptr ? &temp : g;
    b200::Dev down = b200::DevVec::Empty(Shape->data->n), downR = b200::DevVec::Empty(Shape->data->n);
    for (size_t t = 0; t < Shape->T(); t++) {
      Variable* poolVar = Pool[t].get();
      rg[poolVar] = linalg::Vector();
      (*gg)[poolVar] = linalg::Vector();
      Results[t]->bwdR(s, u.Step(t), uR.Step(t), rg, gg);
      b200::Dev d = s.Take(*gg, poolVar, poolVar->len()), dr = s.Take(rg, poolVar, poolVar->len());
      if (d->n > 0) {
        b200::check(af_copy(b200::ctx(), down->p + Shape->offsets[t], d->p, d->n));
        b200::check(af_copy(b200::ctx(), downR->p + Shape->offsets[t], dr->p, dr->n));
      }
    }
    s.Forget(temp);
    Input->bwdPackedR(s, *Shape->Like(Shape->widths, down), *Shape->Like(Shape->widths, downR), rg, g);
  }

 private:
  Seqs host_, host_r_;
  bool have_ = false, have_r_ = false;
};

class MapRBatcher : public RFunc {
 public:
  std::shared_ptr<autofunc::RBatcher> B;
  explicit MapRBatcher(std::shared_ptr<autofunc::RBatcher> b) : B(std::move(b)) {}
  ResultPtr ApplySeqs(ResultPtr r) override { return MapBatcher(B).ApplySeqs(std::move(r)); }  // :71-74
  RResultPtr ApplySeqsR(const RVector& rv, RResultPtr r) override {                            // :77-100
    auto res = std::make_shared<MapBatcherRResult>();
    res->Input = r;
    res->Shape = r->packed();
    Packed shapeR = r->packedR();
    const PackedSeqs& in = *res->Shape;
    if (shapeR && !in.SameShape(*shapeR)) throw Panic("r-output sequence shape mismatch");
    std::vector<int64_t> widths;
    int64_t total = 0;
    for (size_t t = 0; t < in.T(); t++) {
      VariablePtr v = Variable::Wrap(in.Step(t));
      RVariablePtr inVar = RVariable::Wrap(v, shapeR ? shapeR->Step(t) : nullptr);  // RVariable{joined, ROutputVec: joinedR}
      autofunc::RResultPtr output = B->BatchR(rv, inVar, in.counts[t]);
      if (output->len() % in.counts[t] != 0) throw Panic("cannot evenly split vector");
      widths.push_back(output->len() / in.counts[t]);
      total += output->len();
      res->Pool.push_back(v);
      res->Results.push_back(output);
    }
    b200::Dev out = b200::DevVec::Empty(total), outR = b200::DevVec::Zeros(total);
    int64_t at = 0;
    for (auto& o : res->Results) {
      b200::Dev d = o->dev(), dr = o->devR();
      if (d->n > 0) {
        b200::check(af_copy(b200::ctx(), out->p + at, d->p, d->n));
        if (dr) b200::check(af_copy(b200::ctx(), outR->p + at, dr->p, d->n));
      }
      at += d->n;
    }
    res->Out = in.Like(widths, out);
    res->OutR = in.Like(widths, outR);
    return res;
  }
};

}  // namespace seqfunc
}  // namespace autofunc

#endif  // AUTOFUNC_B200_SEQ_HPP_
