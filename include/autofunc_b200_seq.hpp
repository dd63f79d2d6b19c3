// autofunc_b200_seq.hpp -- seqfunc.MapBatcher / MapRBatcher (seqfunc/map_batcher.go:14-140) and the sequence-list
// Results they consume (seqfunc/seqfunc.go:10-53, construct.go), in C++ over include/autofunc_b200.hpp.
//
// A list of vector sequences is `Seqs` = [][]linalg.Vector.  A MapBatcher feeds its Batcher all the vectors of one
// timestep at once (joinTime, seqfunc/map.go:262-270), so T timesteps over B lanes are T batched device calls instead
// of the T*B single-vector calls of MapFunc; the pool-variable protocol of the backward pass (insert grad[poolVar],
// propagate, read, delete -- map_batcher.go:50-61,124-139) runs unchanged on top of the device nodes, because their
// accumulators are flushed into the host maps before each Propagate* call returns.  Sequence-level Results live on the
// host like the reference's (the resident form of a whole recurrent network is b200::LSTM).
#ifndef AUTOFUNC_B200_SEQ_HPP_
#define AUTOFUNC_B200_SEQ_HPP_

#include "autofunc_b200.hpp"

namespace autofunc {
namespace seqfunc {

using Seqs = std::vector<std::vector<linalg::Vector>>;

class Result {  // seqfunc/seqfunc.go:24-35
 public:
  virtual ~Result() = default;
  virtual const Seqs& OutputSeqs() = 0;
  virtual void PropagateGradient(const Seqs& upstream, Gradient& g) = 0;  // must not modify the upstream
};
using ResultPtr = std::shared_ptr<Result>;

class RResult {  // seqfunc/seqfunc.go:39-53
 public:
  virtual ~RResult() = default;
  virtual const Seqs& OutputSeqs() = 0;
  virtual const Seqs& ROutputSeqs() = 0;
  virtual void PropagateRGradient(const Seqs& upstream, const Seqs& upstreamR, RGradient& rg, Gradient* g) = 0;
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

// ---- seqfunc/construct.go ---------------------------------------------------------------------------------------
class ConstResult : public Result {
 public:
  explicit ConstResult(Seqs v) : out_(std::move(v)) {}
  const Seqs& OutputSeqs() override { return out_; }
  void PropagateGradient(const Seqs&, Gradient&) override {}

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

// ---- seqfunc/map_batcher.go ---------------------------------------------------------------------------------------
class MapBatcherResult : public Result {
 public:
  ResultPtr Input;
  std::vector<VariablePtr> Pool;
  std::vector<autofunc::ResultPtr> Results;
  Seqs Output;
  const Seqs& OutputSeqs() override { return Output; }
  void PropagateGradient(const Seqs& u, Gradient& g) override {  // map_batcher.go:50-63
    size_t maxLen = maxSequenceLen(u);
    Seqs downstream(u.size());
    int n;
    for (size_t t = 0; t < maxLen; t++) {
      linalg::Vector upstream = joinTime(u, t, &n);
      Variable* poolVar = Pool.at(t).get();
      g[poolVar] = linalg::Vector(poolVar->Vector.size(), 0.0);
      Results[t]->PropagateGradient(std::move(upstream), g);
      appendSplit(u, t, downstream, g[poolVar]);
      g.erase(poolVar);
    }
    Input->PropagateGradient(downstream, g);
  }
};

class MapBatcher : public virtual Func {
 public:
  std::shared_ptr<autofunc::Batcher> B;
  explicit MapBatcher(std::shared_ptr<autofunc::Batcher> b) : B(std::move(b)) {}
  ResultPtr ApplySeqs(ResultPtr r) override {  // map_batcher.go:20-38
    auto res = std::make_shared<MapBatcherResult>();
    res->Input = r;
    const Seqs& in = r->OutputSeqs();
    res->Output.resize(in.size());
    size_t maxLen = maxSequenceLen(in);
    int n;
    for (size_t t = 0; t < maxLen; t++) {
      VariablePtr inVar = NewVariable(joinTime(in, t, &n));
      autofunc::ResultPtr output = B->Batch(inVar, n);
      res->Pool.push_back(inVar);
      res->Results.push_back(output);
      appendSplit(in, t, res->Output, output->Output());
    }
    return res;
  }
};

class MapBatcherRResult : public RResult {
 public:
  RResultPtr Input;
  std::vector<VariablePtr> Pool;
  std::vector<autofunc::RResultPtr> Results;
  Seqs Output, ROutput;
  const Seqs& OutputSeqs() override { return Output; }
  const Seqs& ROutputSeqs() override { return ROutput; }
  void PropagateRGradient(const Seqs& u, const Seqs& uR, RGradient& rg, Gradient* g) override {  // :118-140
    Gradient temp;  // "We use g for temporary gradients." (map_batcher.go:120-123)
    if (g == nullptr) g = &temp;
    size_t maxLen = maxSequenceLen(u);
    Seqs downstream(u.size()), downstreamR(u.size());
    int n;
    for (size_t t = 0; t < maxLen; t++) {
      linalg::Vector upstream = joinTime(u, t, &n), upstreamR = joinTime(uR, t, &n);
      Variable* poolVar = Pool.at(t).get();
      rg[poolVar] = linalg::Vector(poolVar->Vector.size(), 0.0);
      (*g)[poolVar] = linalg::Vector(poolVar->Vector.size(), 0.0);
      Results[t]->PropagateRGradient(std::move(upstream), std::move(upstreamR), rg, g);
      appendSplit(u, t, downstream, (*g)[poolVar]);
      appendSplit(uR, t, downstreamR, rg[poolVar]);
      g->erase(poolVar);
      rg.erase(poolVar);
    }
    Input->PropagateRGradient(downstream, downstreamR, rg, g);
  }
};

class MapRBatcher : public RFunc {
 public:
  std::shared_ptr<autofunc::RBatcher> B;
  explicit MapRBatcher(std::shared_ptr<autofunc::RBatcher> b) : B(std::move(b)) {}
  ResultPtr ApplySeqs(ResultPtr r) override { return MapBatcher(B).ApplySeqs(std::move(r)); }  // :71-74
  RResultPtr ApplySeqsR(const RVector& rv, RResultPtr r) override {                            // :77-100
    auto res = std::make_shared<MapBatcherRResult>();
    res->Input = r;
    const Seqs &in = r->OutputSeqs(), &inR = r->ROutputSeqs();
    res->Output.resize(in.size());
    res->ROutput.resize(in.size());
    size_t maxLen = maxSequenceLen(in);
    int n;
    for (size_t t = 0; t < maxLen; t++) {
      VariablePtr v = NewVariable(joinTime(in, t, &n));
      auto inVar = std::make_shared<RVariable>(v, joinTime(inR, t, &n));  // RVariable{joined, ROutputVec: joinedR}
      autofunc::RResultPtr output = B->BatchR(rv, inVar, n);
      res->Pool.push_back(v);
      res->Results.push_back(output);
      appendSplit(in, t, res->Output, output->Output());
      appendSplit(in, t, res->ROutput, output->ROutput());
    }
    return res;
  }
};

}  // namespace seqfunc
}  // namespace autofunc

#endif  // AUTOFUNC_B200_SEQ_HPP_
