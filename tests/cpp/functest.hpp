// The reference's acceptance procedure, functest.RFuncChecker.FullCheck (functest/rfunc_checker.go:42-68,
// checker.go:33-113, util.go:13-38), restated against include/autofunc_b200.hpp.  TEST INFRASTRUCTURE: the only
// host arithmetic here is the finite-difference bookkeeping of the checker itself.
#ifndef AUTOFUNC_B200_FUNCTEST_HPP_
#define AUTOFUNC_B200_FUNCTEST_HPP_

#include <cmath>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "autofunc_b200.hpp"

namespace functest {

using namespace autofunc;

constexpr double DefaultDelta = 1e-5;  // functest/func_checker.go:13-21
constexpr double DefaultPrec = 1e-5;

struct AddTwice : RFunc {  // functest/util.go:13-21
  ResultPtr Apply(ResultPtr r) override { return Add(r, r); }
  RResultPtr ApplyR(const RVector&, RResultPtr r) override { return AddR(r, r); }
};
struct MulTwice : RFunc {  // functest/util.go:30-38
  ResultPtr Apply(ResultPtr r) override { return Mul(r, r); }
  RResultPtr ApplyR(const RVector&, RResultPtr r) override { return MulR(r, r); }
};

class RFuncChecker {
 public:
  std::shared_ptr<RFunc> F;
  std::vector<VariablePtr> Vars;
  VariablePtr Input;
  RVector RV;
  double Delta = DefaultDelta, Prec = DefaultPrec;
  std::vector<std::string> Errors;

  RFuncChecker(std::shared_ptr<RFunc> f, std::vector<VariablePtr> vars, VariablePtr input, RVector rv)
      : F(std::move(f)), Vars(std::move(vars)), Input(std::move(input)), RV(std::move(rv)) {}

  // rfunc_checker.go:42-68
  std::vector<std::string> FullCheck() {
    auto sub = [&](std::shared_ptr<RFunc> f, std::vector<VariablePtr> vars, const std::string& tag) {
      RFuncChecker c(std::move(f), std::move(vars), Input, RV);
      c.Delta = Delta;
      c.Prec = Prec;
      c.CheckR(tag);
      c.CheckConsistency(tag);
      Errors.insert(Errors.end(), c.Errors.begin(), c.Errors.end());
    };
    sub(F, Vars, "Standard");
    for (size_t i = 0; i < Vars.size(); i++) sub(F, {Vars[i]}, "Vars[" + std::to_string(i) + "]");
    sub(std::make_shared<ComposedRFunc>(ComposedRFunc{F, std::make_shared<AddTwice>()}), Vars, "Accumulation");
    sub(std::make_shared<ComposedRFunc>(ComposedRFunc{F, std::make_shared<MulTwice>()}), Vars, "Square");
    return Errors;
  }

  // checker.go:33-113
  void CheckR(const std::string& tag) {
    Jac jac, jacR;
    jacobians(&jac, &jacR);
    for (size_t vi = 0; vi < Vars.size(); vi++)
      for (size_t ei = 0; ei < Vars[vi]->Vector.size(); ei++) {
        linalg::Vector approx = approxPartials(Vars[vi], ei);
        for (size_t oi = 0; oi < jac.size(); oi++)
          if (bad(jac[oi][vi][ei], approx[oi])) fail(tag, "gradient", vi, oi, ei, approx[oi], jac[oi][vi][ei]);
      }
    linalg::Vector exp = shiftAlongRV([&] { return applyR()->Output(); });
    linalg::Vector act = applyR()->ROutput();
    for (size_t i = 0; i < act.size(); i++)
      if (bad(act[i], exp[i])) fail(tag, "r-output", 0, i, 0, exp[i], act[i]);
    for (size_t vi = 0; vi < Vars.size(); vi++)
      for (size_t ei = 0; ei < Vars[vi]->Vector.size(); ei++) {
        linalg::Vector approx = shiftAlongRV([&] {
          Jac j, jr;
          jacobians(&j, &jr);
          linalg::Vector col(j.size());
          for (size_t oi = 0; oi < j.size(); oi++) col[oi] = j[oi][vi][ei];
          return col;
        });
        for (size_t oi = 0; oi < jacR.size(); oi++)
          if (bad(jacR[oi][vi][ei], approx[oi])) fail(tag, "r-gradient", vi, oi, ei, approx[oi], jacR[oi][vi][ei]);
      }
  }

  // rfunc_checker.go:203-256: Apply and ApplyR agree on outputs and gradients
  void CheckConsistency(const std::string& tag) {
    linalg::Vector a = F->Apply(Input)->Output();
    linalg::Vector b = applyR()->Output();
    bool same = a.size() == b.size();
    for (size_t i = 0; same && i < a.size(); i++) same = !bad(a[i], b[i]);
    if (!same) Errors.push_back(tag + " output inconsistency between Apply and ApplyR");
    Jac jn = jacobianNonR(), jr, jrr;
    jacobians(&jr, &jrr);
    if (jn.size() != jr.size()) {
      Errors.push_back(tag + " jacobian counts differ");
      return;
    }
    for (size_t oi = 0; oi < jn.size(); oi++)
      for (size_t vi = 0; vi < Vars.size(); vi++)
        for (size_t ei = 0; ei < jn[oi][vi].size(); ei++)
          if (bad(jr[oi][vi][ei], jn[oi][vi][ei]))
            fail(tag, "Apply/ApplyR gradient", vi, oi, ei, jn[oi][vi][ei], jr[oi][vi][ei]);
  }

 private:
  using Jac = std::vector<std::vector<linalg::Vector>>;  // [output][variable] -> partials

  RResultPtr applyR() { return F->ApplyR(RV, NewRVariable(Input, RV)); }
  bool bad(double a, double e) const { return std::isnan(a) != std::isnan(e) || std::fabs(a - e) > Prec; }
  void fail(const std::string& tag, const char* what, size_t vi, size_t oi, size_t ei, double want, double got) {
    std::ostringstream s;
    s << tag << " " << what << ": var " << vi << " output " << oi << " entry " << ei << ": expected " << want
      << " got " << got;
    Errors.push_back(s.str());
  }
  // exact quantities (rfunc_checker.go:98-113,150-166)
  void jacobians(Jac* jac, Jac* jacR) {
    RResultPtr out = applyR();
    size_t m = out->Output().size();
    for (size_t i = 0; i < m; i++) {
      Gradient grad = NewGradient(Vars);
      RGradient rgrad = NewRGradient(Vars);
      linalg::Vector up(m, 0.0);
      up[i] = 1;
      out->PropagateRGradient(up, linalg::Vector(m, 0.0), rgrad, &grad);
      std::vector<linalg::Vector> row, rowR;
      for (auto& v : Vars) {
        row.push_back(grad[v.get()]);
        rowR.push_back(rgrad[v.get()]);
      }
      jac->push_back(row);
      jacR->push_back(rowR);
    }
  }
  Jac jacobianNonR() {
    ResultPtr out = F->Apply(Input);
    size_t m = out->Output().size();
    Jac jac;
    for (size_t i = 0; i < m; i++) {
      Gradient grad = NewGradient(Vars);
      linalg::Vector up(m, 0.0);
      up[i] = 1;
      out->PropagateGradient(up, grad);
      std::vector<linalg::Vector> row;
      for (auto& v : Vars) row.push_back(grad[v.get()]);
      jac.push_back(row);
    }
    return jac;
  }
  // finite differences (rfunc_checker.go:85-95,117-147,170-190)
  linalg::Vector approxPartials(const VariablePtr& v, size_t idx) {
    double old = v->Vector[idx];
    v->Vector[idx] = old + Delta;
    linalg::Vector v1 = applyR()->Output();
    v->Vector[idx] = old - Delta;
    linalg::Vector v2 = applyR()->Output();
    v->Vector[idx] = old;
    for (size_t i = 0; i < v1.size(); i++) v1[i] = (v1[i] - v2[i]) * (0.5 / Delta);
    return v1;
  }
  linalg::Vector shiftAlongRV(const std::function<linalg::Vector()>& fn) {
    std::vector<std::pair<Variable*, linalg::Vector>> backups;
    for (auto& kv : RV) backups.emplace_back(kv.first, kv.first->Vector);
    for (auto& kv : RV)
      for (size_t i = 0; i < kv.second.size(); i++) kv.first->Vector[i] -= Delta * kv.second[i];
    linalg::Vector a = fn();
    for (auto& b : backups) b.first->Vector = b.second;
    for (auto& kv : RV)
      for (size_t i = 0; i < kv.second.size(); i++) kv.first->Vector[i] += Delta * kv.second[i];
    linalg::Vector b = fn();
    for (auto& bk : backups) bk.first->Vector = bk.second;
    for (size_t i = 0; i < a.size(); i++) a[i] = (b[i] - a[i]) * (0.5 / Delta);
    return a;
  }
};

}  // namespace functest

#endif  // AUTOFUNC_B200_FUNCTEST_HPP_
