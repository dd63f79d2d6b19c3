// C++ host-side tests of the drop-in (include/autofunc_b200.hpp): the reference's own Go tests for the path,
// restated test for test (file:line cited at each), plus a `run` mode that pytest drives with seeded inputs to compare
// the C++ host path with the CPU oracle at 1e-12 + 1e-10|ref| (tests/test_gpu_cpp_host.py).
//
//   host_tests selftest [name-substring]      run the restated reference tests (needs a B200)
//   host_tests run <case> <in.bin> <out.bin> <args...>
//
// Files are sequences of vectors in the reference's Variable wire format (variable.go:24-37).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>

#include "autofunc_b200.hpp"
#include "autofunc_b200_seq.hpp"
#include "functest.hpp"

using namespace autofunc;
using functest::RFuncChecker;
using linalg::Vector;

namespace {

struct TestFailure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

void Fatalf(const std::string& msg) { throw TestFailure(msg); }

std::string Show(const Vector& v) {
  std::string s = "[";
  for (size_t i = 0; i < v.size(); i++) s += (i ? " " : "") + std::to_string(v[i]);
  return s + "]";
}

void ExpectFullCheck(RFuncChecker c) {
  std::vector<std::string> errs = c.FullCheck();
  if (errs.empty()) return;
  std::string all = std::to_string(errs.size()) + " FullCheck errors; first: " + errs[0];
  Fatalf(all);
}

void ExpectPanic(const std::string& want, const std::function<void()>& fn) {
  try {
    fn();
  } catch (const Panic& p) {
    if (std::string(p.what()) != want) Fatalf("expected panic \"" + want + "\" but got \"" + p.what() + "\"");
    return;
  }
  Fatalf("expected panic \"" + want + "\" but nothing was thrown");
}

template <class T, class... A>
std::shared_ptr<T> New(A&&... a) {
  return std::make_shared<T>(std::forward<A>(a)...);
}

// ---- shared helpers ------------------------------------------------------------------------------------------------
struct Net {
  std::vector<VariablePtr> vars;
  ComposedRBatcher layers;
  std::shared_ptr<ComposedRFunc> func = New<ComposedRFunc>();
  RVector rv;
};
double Uniform(uint64_t* s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2 - 1;
}
Vector RandomVector(size_t n, uint64_t* s, double scale = 1) {
  Vector v(n);
  for (double& e : v) e = scale * Uniform(s);
  return v;
}
Net MakeNet(const std::vector<int>& sizes, uint64_t seed, bool biasesInRV = true) {
  Net net;
  uint64_t s = seed;
  for (size_t l = 0; l + 1 < sizes.size(); l++) {
    int in = sizes[l], out = sizes[l + 1];
    VariablePtr w = NewVariable(RandomVector((size_t)in * out, &s, 1 / std::sqrt((double)in)));
    VariablePtr b = NewVariable(RandomVector((size_t)out, &s));
    net.vars.push_back(w);
    net.vars.push_back(b);
    auto lt = New<LinTran>(w, out, in);
    auto la = New<LinAdd>(b);
    auto sg = New<Sigmoid>();
    net.layers.push_back(lt);
    net.layers.push_back(la);
    net.layers.push_back(sg);
    net.func->push_back(lt);
    net.func->push_back(la);
    net.func->push_back(sg);
    net.rv[w.get()] = RandomVector(w->Vector.size(), &s);
    if (biasesInRV) net.rv[b.get()] = RandomVector(b->Vector.size(), &s);
  }
  return net;
}
void ExpectClose(const Vector& a, const Vector& b, const std::string& what, double rtol = 1e-10, double atol = 1e-12) {
  if (a.size() != b.size()) Fatalf(what + ": lengths differ");
  for (size_t i = 0; i < a.size(); i++)
    if (!(std::fabs(a[i] - b[i]) <= atol + rtol * std::fabs(b[i])))
      Fatalf(what + ": entry " + std::to_string(i) + ": " + std::to_string(a[i]) + " vs " + std::to_string(b[i]));
}


// A foreign host-only Result in the middle of a device chain (the drop-in must coexist with the reference's own nodes).
struct HostDouble : Result {
  ResultPtr in;
  Vector out;
  explicit HostDouble(ResultPtr r) : in(std::move(r)), out(in->Output()) {
    for (double& e : out) e *= 2;
  }
  const Vector& Output() override { return out; }
  bool Constant(const Gradient& g) override { return in->Constant(g); }
  void PropagateGradient(Vector up, Gradient& g) override {
    for (double& e : up) e *= 2;
    in->PropagateGradient(std::move(up), g);
  }
};

// ---- fixtures: test/lin_tran_test.go:12-53 ----------------------------------------------------------------
struct LinTranFixture {
  std::shared_ptr<LinTran> mat1 = New<LinTran>(NewVariable({1, 2, 3, 3, 4, 5, -6, 1, 7, 8, -10, 4}), 3, 4);
  std::shared_ptr<LinTran> mat2 = New<LinTran>(NewVariable({4, 2, 3}), 1, 3);
  VariablePtr vec = NewVariable({4, 3, 2, 1});
  VariablePtr vec2 = NewVariable({5, 1, -3, 2, 1, 2, 3, 4});
  std::vector<VariablePtr> vars{mat1->Data, mat2->Data, vec, vec2};
  RVector rv{
      {mat1->Data.get(),
       {0.40350, 0.75874, 0.87843, 0.86287, 0.97869, 0.27141, 0.84472, 0.43952, 0.49841, 0.46748, 0.71821, 0.57590}},
      {mat2->Data.get(), {0.45823, 0.66692, 0.14662}},
      {vec.get(), {0, -1, 3, -7}},
      {vec2.get(), {1, 3, -2, 3, -3, 2, 1, 7}},
  };
};

// test/lin_tran_test.go:55-63
struct LinTranBatchTest : RFunc {
  std::shared_ptr<LinTran> m1, m2;
  LinTranBatchTest(std::shared_ptr<LinTran> a, std::shared_ptr<LinTran> b) : m1(std::move(a)), m2(std::move(b)) {}
  ResultPtr Apply(ResultPtr v) override { return m2->Batch(m1->Batch(v, 2), 2); }
  RResultPtr ApplyR(const RVector& rv, RResultPtr v) override { return m2->BatchR(rv, m1->BatchR(rv, v, 2), 2); }
};

void TestLinTranOutput() {  // test/lin_tran_test.go:65-73
  LinTranFixture f;
  Vector actual = f.mat1->Apply(f.vec)->Output();
  if (actual != Vector{19, 20, 36}) Fatalf("expected [19 20 36] got " + Show(actual));
  RResultPtr r = f.mat1->ApplyR(f.rv, NewRVariable(f.vec, f.rv));
  if (r->Output() != Vector{19, 20, 36}) Fatalf("ApplyR: expected [19 20 36] got " + Show(r->Output()));
}

void TestLinTran() {  // test/lin_tran_test.go:75-83
  LinTranFixture f;
  ExpectFullCheck(RFuncChecker(New<ComposedRFunc>(ComposedRFunc{f.mat1, f.mat2}), f.vars, f.vec, f.rv));
}

void TestLinTranBatchOutput() {  // test/lin_tran_test.go:85-97
  LinTranFixture f;
  Vector actual = LinTranBatchTest(f.mat1, f.mat2).Apply(f.vec2)->Output();
  if (actual != Vector{349, 131}) Fatalf("expected [349 131] got " + Show(actual));
}

void TestLinTranBatch() {  // test/lin_tran_test.go:99-107
  LinTranFixture f;
  ExpectFullCheck(RFuncChecker(New<LinTranBatchTest>(f.mat1, f.mat2), f.vars, f.vec2, f.rv));
}

void TestLinAdd() {  // test/lin_add_test.go:11-30
  VariablePtr v1 = NewVariable({1, 3, 2, 1}), v2 = NewVariable({5, -3, 5, 4});
  RVector rv{{v1.get(), {0.5, -10, 5, 3.14}}, {v2.get(), {0, 5, -15, -30}}};
  ExpectFullCheck(RFuncChecker(New<ComposedRFunc>(ComposedRFunc{New<LinAdd>(v1), New<LinAdd>(v2)}), {v1, v2}, v2, rv));
}

// ---- test/math_funcs_test.go:13-100 ----------------------------------------------------------------------------
struct MathFixture {
  VariablePtr vec = NewVariable({1, -0.5, 0.3, 0.7}), vecPos = NewVariable({1, 0.5, 0.3, 0.7});
  std::vector<VariablePtr> vars{vec, vecPos};
  RVector rv{{vec.get(), {0.5, -10, 5, 3.14}}};
  void Check(std::shared_ptr<RFunc> f, bool positive = false) {
    ExpectFullCheck(RFuncChecker(std::move(f), vars, positive ? vecPos : vec, rv));
  }
};
void TestExp() { MathFixture().Check(New<Exp>()); }
void TestLog() { MathFixture().Check(New<Log>(), true); }
void TestSquaredNorm() { MathFixture().Check(New<SquaredNorm>()); }
void TestSigmoid() {
  MathFixture().Check(New<ComposedRFunc>(ComposedRFunc{New<Sigmoid>(), New<Sigmoid>(), New<Sigmoid>()}));
}
void TestLogSigmoid() {
  MathFixture().Check(New<ComposedRFunc>(ComposedRFunc{New<LogSigmoid>(), New<LogSigmoid>(), New<LogSigmoid>()}));
}
void TestSoftmax() { MathFixture().Check(New<Softmax>()); }
void TestSoftmax3() { MathFixture().Check(New<Softmax>(3)); }
void TestSin() { MathFixture().Check(New<Sin>()); }
void TestTanh() { MathFixture().Check(New<Tanh>()); }  // extension; Sigmoid's contract

// ---- test/slices_test.go:11-120 --------------------------------------------------------------------------------
struct SlicesFixture {
  VariablePtr v1 = NewVariable({1, 2, -4, 3, -2}), v2 = NewVariable({4, -2, 3, -0.5, 0.3}),
              v3 = NewVariable({0.99196, 0.48826, 0.51066, 0.66715, 0.44423});
  std::vector<VariablePtr> vars{v1, v2, v3};
  // vector 2 intentionally left out (slices_test.go:30-31)
  RVector rv{{v1.get(), {0.340162, -0.325063, 0.179612, 0.056463, -0.812274}},
             {v3.get(), {0.59824, -0.63322, 0.13379, 0.99559, 0.53748}}};
};
struct ConcatTestFunc : RFunc {
  SlicesFixture* f;
  explicit ConcatTestFunc(SlicesFixture* fx) : f(fx) {}
  ResultPtr Apply(ResultPtr r) override { return Concat({f->v1, r, f->v3}); }
  RResultPtr ApplyR(const RVector&, RResultPtr r) override {
    return ConcatR({NewRVariable(f->v1, f->rv), r, NewRVariable(f->v3, f->rv)});
  }
};
struct SliceTestFunc : RFunc {
  bool WrapInput;
  explicit SliceTestFunc(bool w) : WrapInput(w) {}
  ResultPtr Apply(ResultPtr r) override {
    if (WrapInput) r = Add(r, Scale(r, 0));
    return Slice(r, 1, 3);
  }
  RResultPtr ApplyR(const RVector&, RResultPtr r) override {
    if (WrapInput) r = AddR(r, ScaleR(r, 0));
    return SliceR(r, 1, 3);
  }
};
struct RepeatTestFunc : RFunc {
  ResultPtr Apply(ResultPtr r) override { return Repeat(Add(Add(r, Scale(r, 0)), r), 3); }
  RResultPtr ApplyR(const RVector&, RResultPtr r) override { return RepeatR(AddR(AddR(r, ScaleR(r, 0)), r), 3); }
};
void TestConcat() {
  SlicesFixture f;
  ExpectFullCheck(RFuncChecker(New<ConcatTestFunc>(&f), f.vars, f.v2, f.rv));
}
void TestSlice() {
  SlicesFixture f;
  ExpectFullCheck(RFuncChecker(New<SliceTestFunc>(false), f.vars, f.v1, f.rv));
}
void TestWrappedSlice() {
  SlicesFixture f;
  ExpectFullCheck(RFuncChecker(New<SliceTestFunc>(true), f.vars, f.v1, f.rv));
}
void TestRepeat() {
  SlicesFixture f;
  ExpectFullCheck(RFuncChecker(New<RepeatTestFunc>(), f.vars, f.v1, f.rv));
}

// Concat under a batcher (batcher.go:57-84): per-sample concatenation of packed batches, by the reference's checker and
// against Slice / Concat applied sample by sample
struct ConcatBatchedTestFunc : RFunc {
  VariablePtr other;
  const RVector* rv;
  ResultPtr Apply(ResultPtr r) override { return Mul(ConcatBatched(2, {r, other}), ConcatBatched(2, {other, r})); }
  RResultPtr ApplyR(const RVector&, RResultPtr r) override {
    RResultPtr o = NewRVariable(other, *rv);
    return MulR(ConcatBatchedR(2, {r, o}), ConcatBatchedR(2, {o, r}));
  }
};
void TestConcatBatched() {
  VariablePtr x = NewVariable({1, -2, 0.5, 3}), other = NewVariable({0.3, -0.7, 2, 1.5, -1, 0.25});
  RVector rv{{x.get(), {0.5, 1, -1, 2}}};  // `other` absent: zero R
  auto f = New<ConcatBatchedTestFunc>();
  f->other = other;
  f->rv = &rv;
  ExpectFullCheck(RFuncChecker(f, {x, other}, x, rv));
  ExpectClose(ConcatBatched(2, {x, other})->Output(), {1, -2, 0.3, -0.7, 2, 0.5, 3, 1.5, -1, 0.25}, "layout");
  ExpectClose(ConcatBatched(2, {x, other})->Output(),
              Concat({Slice(x, 0, 2), Slice(other, 0, 3), Slice(x, 2, 4), Slice(other, 3, 6)})->Output(), "vs Concat of Slices");
  ExpectPanic("input count does not divide input length", [&] { ConcatBatched(4, {x, other}); });
}

// ---- test/arithmetic_test.go:12-110 (vector 2 intentionally absent from the RVector) -----------------------------
struct ArithmeticFixture {
  VariablePtr v1 = NewVariable({1, 2, -4, 3, -2}), v2 = NewVariable({4, -2, 3, -0.5, 0.3}),
              v3 = NewVariable({0.99196, 0.48826, 0.51066, 0.66715, 0.44423}),
              v4 = NewVariable({0.509893, 0.874112, -0.080468, 0.372740, -0.172097});
  std::vector<VariablePtr> vars{v1, v2, v3, v4};
  RVector rv{{v1.get(), {0.340162, -0.325063, 0.179612, 0.056463, -0.812274}},
             {v3.get(), {0.59824, -0.63322, 0.13379, 0.99559, 0.53748}},
             {v4.get(), {4.2222, 5.2762, -7.5762, 2.3420, -1.8927}}};
};
struct ArithmeticTestFunc : RFunc {  // test/arithmetic_test.go:45-69
  ArithmeticFixture* f;
  explicit ArithmeticTestFunc(ArithmeticFixture* fx) : f(fx) {}
  ResultPtr Apply(ResultPtr r) override {
    ResultPtr sq1 = Square(Mul(Scale(Add(r, f->v1), -2), f->v2));
    ResultPtr sum1 = AddScaler(Add(Mul(sq1, f->v3), Scale(f->v4, -0.5)), 2);
    ResultPtr powed = Div(Pow(Pow(Inverse(sum1), 2), 1 / 3.0), sum1);
    ResultPtr allSum = SumAll(AddFirst(powed, f->v1));
    allSum = Sub(allSum, SumAllLogDomain(f->v2));
    return ScaleFirst(AddLogDomain(f->v1, f->v2), allSum);
  }
  RResultPtr ApplyR(const RVector& v, RResultPtr r) override {
    RResultPtr r1 = NewRVariable(f->v1, v), r2 = NewRVariable(f->v2, v), r3 = NewRVariable(f->v3, v),
               r4 = NewRVariable(f->v4, v);
    RResultPtr sq1 = SquareR(MulR(ScaleR(AddR(r, r1), -2), r2));
    RResultPtr sum1 = AddScalerR(AddR(MulR(sq1, r3), ScaleR(r4, -0.5)), 2);
    RResultPtr powed = DivR(PowR(PowR(InverseR(sum1), 2), 1 / 3.0), sum1);
    RResultPtr allSum = SumAllR(AddFirstR(powed, r1));
    allSum = SubR(allSum, SumAllLogDomainR(r2));
    return ScaleFirstR(AddLogDomainR(r1, r2), allSum);
  }
};
void TestArithmetic() {  // test/arithmetic_test.go:71-79
  ArithmeticFixture f;
  ExpectFullCheck(RFuncChecker(New<ArithmeticTestFunc>(&f), f.vars, f.v4, f.rv));
}
void TestAddLogDomainOutput() {  // test/arithmetic_test.go:81-97
  ArithmeticFixture f;
  Vector expected;
  for (size_t i = 0; i < f.v1->Vector.size(); i++) expected.push_back(std::log(std::exp(f.v1->Vector[i]) + std::exp(f.v2->Vector[i])));
  ExpectClose(AddLogDomain(f.v1, f.v2)->Output(), expected, "AddLogDomain", 0, 1e-5);
}
void TestSumAllLogDomainOutput() {  // test/arithmetic_test.go:99-110
  VariablePtr in = NewVariable({3, 3.5, -1, 2});
  double sum = 0;
  for (double x : in->Vector) sum += std::exp(x);
  ExpectClose(SumAllLogDomain(in)->Output(), {std::log(sum)}, "SumAllLogDomain", 0, 1e-8);
}
void TestCosOutput() {  // test/math_funcs_test.go:102-111
  uint64_t s = 1234;
  for (int i = 0; i < 10; i++) {
    double arg = Uniform(&s) * 10;
    ExpectClose(Cos().Apply(NewVariable({arg}))->Output(), {std::cos(arg)}, "Cos", 0, 1e-5);
  }
}
void TestNormOutput() {  // test/math_funcs_test.go:113-120
  ExpectClose(Norm().Apply(NewVariable({3, 4}))->Output(), {5.0}, "Norm", 0, 1e-5);
}
void TestNorm() { MathFixture().Check(New<Norm>()); }  // test/math_funcs_test.go:122-130
// Division and the squared-error head against their reference compositions (arithmetic.go:425-427; math_funcs.go:191-199)
struct DivTestFunc : RFunc {
  ArithmeticFixture* f;
  explicit DivTestFunc(ArithmeticFixture* fx) : f(fx) {}
  ResultPtr Apply(ResultPtr r) override { return Div(Mul(r, f->v3), AddScaler(Square(f->v2), 1)); }
  RResultPtr ApplyR(const RVector& v, RResultPtr r) override {
    return DivR(MulR(r, NewRVariable(f->v3, v)), AddScalerR(SquareR(NewRVariable(f->v2, v)), 1));
  }
};
struct SquaredErrorTestFunc : RFunc {
  ArithmeticFixture* f;
  bool composed;
  SquaredErrorTestFunc(ArithmeticFixture* fx, bool c) : f(fx), composed(c) {}
  ResultPtr Apply(ResultPtr r) override {
    ResultPtr a = Mul(r, f->v3), y = Add(f->v1, f->v4);
    return composed ? Scale(SquaredNorm().Apply(Sub(a, y)), 0.5) : SquaredError(a, y);
  }
  RResultPtr ApplyR(const RVector& v, RResultPtr r) override {
    RResultPtr a = MulR(r, NewRVariable(f->v3, v)), y = AddR(NewRVariable(f->v1, v), NewRVariable(f->v4, v));
    return composed ? ScaleR(SquaredNorm().ApplyR(v, SubR(a, y)), 0.5) : SquaredErrorR(a, y);
  }
};
void TestDiv() {
  ArithmeticFixture f;
  ExpectFullCheck(RFuncChecker(New<DivTestFunc>(&f), f.vars, f.v4, f.rv));
}
void TestSquaredError() {
  ArithmeticFixture f;
  ExpectFullCheck(RFuncChecker(New<SquaredErrorTestFunc>(&f, false), f.vars, f.v2, f.rv));
  f.rv[f.v2.get()] = {0.3, -0.1, 0.7, 0.2, -0.9};
  Gradient g1 = NewGradient(f.vars), g2 = NewGradient(f.vars);
  RGradient rg1 = NewRGradient(f.vars), rg2 = NewRGradient(f.vars);
  RResultPtr fused = SquaredErrorTestFunc(&f, false).ApplyR(f.rv, NewRVariable(f.v2, f.rv));
  RResultPtr chain = SquaredErrorTestFunc(&f, true).ApplyR(f.rv, NewRVariable(f.v2, f.rv));
  ExpectClose(fused->Output(), chain->Output(), "cost");
  ExpectClose(fused->ROutput(), chain->ROutput(), "r-cost");
  fused->PropagateRGradient({1.5}, {-0.25}, rg1, &g1);
  chain->PropagateRGradient({1.5}, {-0.25}, rg2, &g2);
  for (auto& v : f.vars) {
    ExpectClose(g1[v.get()], g2[v.get()], "squared-error gradient");
    ExpectClose(rg1[v.get()], rg2[v.get()], "squared-error r-gradient");
  }
}

// ---- test/pool_test.go:11-42 --------------------------------------------------------------------------------------
struct PoolTestFunc : RFunc {
  ResultPtr Apply(ResultPtr r) override {
    return Pool(r, [](ResultPtr pooled) { return Pow(Exp().Apply(AddScaler(pooled, 2)), 0.5); });
  }
  RResultPtr ApplyR(const RVector& v, RResultPtr r) override {
    return PoolR(r, [&](RResultPtr pooled) { return PowR(Exp().ApplyR(v, AddScalerR(pooled, 2)), 0.5); });
  }
};
void TestPool() {
  VariablePtr var = NewVariable({1, -2, 3, -4, 5});
  RVector rv{{var.get(), {1, -1, 2, -2, 3}}};
  auto f = New<ComposedRFunc>(ComposedRFunc{New<Sigmoid>(), New<Exp>(), New<PoolTestFunc>()});
  ExpectFullCheck(RFuncChecker(f, {var}, var, rv));
}
// Pool's purpose (pool.go:5-20): a pooled input used k times is back-propagated through ONCE, and on the device the
// pool variable's gradient never visits the host.
void TestPoolPropagatesOnce() {
  Net net = MakeNet({6, 5}, 41);
  uint64_t s = 43;
  VariablePtr x = NewVariable(RandomVector(6, &s));
  auto uses = [](ResultPtr p) { return Add(Mul(p, p), Add(Square(p), Scale(p, 3))); };  // p appears five times
  ResultPtr pooled = Pool(net.layers.Batch(x, 1), uses), plain = uses(net.layers.Batch(x, 1));
  ExpectClose(pooled->Output(), plain->Output(), "pooled output");
  Gradient g1 = NewGradient(net.vars), g2 = NewGradient(net.vars);
  Vector up = RandomVector(5, &s);
  int64_t before = b200::Context::Default().LaunchCount();
  pooled->PropagateGradient(up, g1);
  int64_t pooledLaunches = b200::Context::Default().LaunchCount() - before;
  before = b200::Context::Default().LaunchCount();
  plain->PropagateGradient(up, g2);
  int64_t plainLaunches = b200::Context::Default().LaunchCount() - before;
  for (auto& v : net.vars) ExpectClose(g1[v.get()], g2[v.get()], "pooled gradient");
  if (g1.size() != net.vars.size()) Fatalf("the pool variable was left in the gradient map");
  if (pooledLaunches >= plainLaunches)
    Fatalf("pooling did not save work: " + std::to_string(pooledLaunches) + " vs " + std::to_string(plainLaunches));
}
// PoolAll / PoolSplit with a constant input, an unused pool variable, a nil Gradient and a foreign host node inside f
void TestPoolAllEdgeCases() {
  ArithmeticFixture f;
  auto body = [&](const std::vector<RResultPtr>& in) { return AddR(MulR(in[0], in[1]), ScaleR(in[0], 2)); };  // in[2] unused
  auto build = [&] {
    return PoolAllR({MulR(NewRVariable(f.v1, f.rv), NewRVariable(f.v3, f.rv)), SquareR(NewRVariable(f.v2, f.rv)),
                     NewRVariable(f.v4, f.rv)},
                    body);
  };
  RResultPtr direct = body({MulR(NewRVariable(f.v1, f.rv), NewRVariable(f.v3, f.rv)), SquareR(NewRVariable(f.v2, f.rv)),
                            NewRVariable(f.v4, f.rv)});
  RResultPtr pooled = build();
  ExpectClose(pooled->Output(), direct->Output(), "PoolAllR output");
  ExpectClose(pooled->ROutput(), direct->ROutput(), "PoolAllR r-output");
  Vector up{1, -2, 0.5, 3, -1}, upR{0.1, 0.2, -0.3, 0.4, 0.5};
  Gradient g1 = NewGradient(f.vars), g2 = NewGradient(f.vars);
  RGradient rg1 = NewRGradient(f.vars), rg2 = NewRGradient(f.vars), rg3 = NewRGradient({f.v1});
  pooled->PropagateRGradient(up, upR, rg1, &g1);
  direct->PropagateRGradient(up, upR, rg2, &g2);
  for (auto& v : f.vars) {
    ExpectClose(g1[v.get()], g2[v.get()], "PoolAllR gradient");
    ExpectClose(rg1[v.get()], rg2[v.get()], "PoolAllR r-gradient");
  }
  for (double e : g1[f.v4.get()])
    if (e != 0) Fatalf("an unused pool input received gradient");
  build()->PropagateRGradient(up, upR, rg3, nullptr);  // nil Gradient (pool.go:164-166), v2's branch constant
  if (rg3.size() != 1) Fatalf("pool variables were left in the r-gradient map");
  ExpectClose(rg3[f.v1.get()], rg2[f.v1.get()], "PoolAllR nil-grad r-gradient");
  RGradient none;
  if (!pooled->Constant(none, nullptr) || pooled->Constant(rg3, nullptr) || !pooled->Constant(NewRGradient({f.v4}), nullptr))
    Fatalf("pooledRResult.Constant is wrong");
  // a host-only node inside the pooled function adds into the temporary's HOST vector; Session::Take folds it in
  ResultPtr a = Pool(Scale(f.v1, 0.5), [](ResultPtr p) { return Add(New<HostDouble>(p), Square(p)); });
  ResultPtr b = Add(Scale(Scale(f.v1, 0.5), 2), Square(Scale(f.v1, 0.5)));
  Gradient ga = NewGradient(f.vars), gb = NewGradient(f.vars);
  a->PropagateGradient(up, ga);
  b->PropagateGradient(up, gb);
  ExpectClose(a->Output(), b->Output(), "pool with host node: output");
  ExpectClose(ga[f.v1.get()], gb[f.v1.get()], "pool with host node: gradient");
  ResultPtr sp = PoolSplit(5, Scale(f.v2, 2), [](const std::vector<ResultPtr>& parts) { return Mul(parts[1], parts[3]); });
  ExpectClose(sp->Output(), {-4 * -1.0}, "PoolSplit output");
  Gradient gs = NewGradient(f.vars);
  sp->PropagateGradient({1}, gs);
  ExpectClose(gs[f.v2.get()], {0, -2, 0, -8, 0}, "PoolSplit gradient");
}

// ---- test/fold_test.go:12-65 --------------------------------------------------------------------------------------
struct FoldTestFunc : RFunc {
  VariablePtr initState;
  explicit FoldTestFunc(VariablePtr s) : initState(std::move(s)) {}
  ResultPtr Apply(ResultPtr in) override {
    return Fold(initState, Split((int)(in->len() / 2), in), [](ResultPtr state, ResultPtr x) { return Mul(Inverse(state), x); });
  }
  RResultPtr ApplyR(const RVector& rv, RResultPtr in) override {
    return FoldR(NewRVariable(initState, rv), SplitR((int)(in->len() / 2), in),
                 [](RResultPtr state, RResultPtr x) { return MulR(InverseR(state), x); });
  }
};
void TestFoldOut() {
  VariablePtr in = NewVariable({1, 2, 2.5, 1.5, 3, 0.5}), init = NewVariable({2, 1});
  ExpectClose(FoldTestFunc(init).Apply(in)->Output(), {0.6, 2.0 / 3.0}, "Fold output", 0, 1e-5);
}
void TestFold() {
  std::vector<VariablePtr> vars{NewVariable({1, 2, 2.5, 1.5, 3, 0.5}), NewVariable({2, 1})};
  RVector rv;
  uint64_t s = 77;
  for (auto& v : vars) rv[v.get()] = RandomVector(v->Vector.size(), &s);
  ExpectFullCheck(RFuncChecker(New<FoldTestFunc>(vars[1]), vars, vars[0], rv));
  // an empty input list returns the state itself (fold.go:57-60)
  Gradient g = NewGradient(vars);
  ResultPtr r = Fold(vars[1], {}, [](ResultPtr st, ResultPtr) { return st; });
  r->PropagateGradient({1, 2}, g);
  ExpectClose(g[vars[1].get()], {1, 2}, "empty fold");
  // a fold whose state stops depending on the past (fold.go:70-73): the loop breaks, nothing older is touched
  ResultPtr cut = Fold(vars[1], Split(3, vars[0]), [](ResultPtr, ResultPtr x) { return Scale(x, 2); });
  Gradient g2 = NewGradient(vars);
  cut->PropagateGradient({1, 1}, g2);
  ExpectClose(g2[vars[0].get()], {0, 0, 0, 0, 2, 2}, "fold cut off from its past");
  ExpectClose(g2[vars[1].get()], {0, 0}, "fold cut off from its past: initial state");
  if (g2.size() != 2) Fatalf("fold left a pool variable in the map");
}

// ---- test/lin_alg_test.go:12-22,56-74 -----------------------------------------------------------------------------
struct MatMulVecTest : RFunc {
  LinTranFixture* f;
  explicit MatMulVecTest(LinTranFixture* fx) : f(fx) {}
  ResultPtr Apply(ResultPtr in) override { return MatMulVec(f->mat1->Data, 3, 4, in); }
  RResultPtr ApplyR(const RVector& rv, RResultPtr in) override { return MatMulVecR(NewRVariable(f->mat1->Data, rv), 3, 4, in); }
};
void TestMatMulVecOutput() {
  LinTranFixture f;
  ExpectClose(MatMulVecTest(&f).Apply(f.vec)->Output(), {19, 20, 36}, "MatMulVec", 0, 1e-5);
}
void TestMatMulVecChecks() {
  LinTranFixture f;
  ExpectFullCheck(RFuncChecker(New<MatMulVecTest>(&f), f.vars, f.vec, f.rv));
}
// MatMulVecs with a COMPUTED matrix (the case LinTran cannot express) and several vectors, against the operator chain
struct MatMulVecsComputed : RFunc {
  LinTranFixture* f;
  explicit MatMulVecsComputed(LinTranFixture* fx) : f(fx) {}
  ResultPtr Apply(ResultPtr in) override { return MatMulVecs(Square(Scale(f->mat1->Data, 0.5)), 3, 4, in); }
  RResultPtr ApplyR(const RVector& rv, RResultPtr in) override {
    return MatMulVecsR(SquareR(ScaleR(NewRVariable(f->mat1->Data, rv), 0.5)), 3, 4, in);
  }
};
void TestMatMulVecsComputedMatrix() {
  LinTranFixture f;
  ExpectFullCheck(RFuncChecker(New<MatMulVecsComputed>(&f), f.vars, f.vec2, f.rv));
  ExpectPanic("invalid matrix data size", [&] { MatMulVecs(f.vec, 3, 4, f.vec); });
  ExpectPanic("invalid vecs size", [&] { MatMulVecs(f.mat1->Data, 3, 4, f.mat2->Data); });
  ExpectPanic("count does not divide input length", [&] { Split(3, f.vec); });
}

// ---- test/variable_test.go:12-35 ---------------------------------------------------------------------------------
void TestVariableSerialize() {
  Vector x(17);
  uint64_t s = 88172645463325252ull;
  for (double& e : x) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    e = (double)(int64_t)s / 9.2e18;
  }
  VariablePtr v = NewVariable(x);
  std::string data = v->Serialize() + NewVariable({1.5, -2.5})->Serialize();
  size_t pos = 0;
  VariablePtr a = Variable::Deserialize(data, &pos), b = Variable::Deserialize(data, &pos);
  if (a->Vector != x) Fatalf("round trip changed the vector");
  if (b->Vector != Vector{1.5, -2.5} || pos != data.size()) Fatalf("second variable did not round-trip");
  if (data.size() != 8 + 17 * 8 + 8 + 2 * 8 || (unsigned char)data[0] != 17) Fatalf("wire format is not LE u64 count + f64s");
}

// ---- batcher.go:34-111: the batched form of a Func equals the per-sample form (seqfunc TestMapBatcherOutput's idea) --
void TestRFuncBatcherMatchesBatch() {
  Net net = MakeNet({5, 7, 3}, 11);
  const int n = 4;
  uint64_t s = 99;
  VariablePtr x = NewVariable(RandomVector(5 * n, &s));
  net.rv[x.get()] = RandomVector(5 * n, &s);
  std::vector<VariablePtr> vars = net.vars;
  vars.push_back(x);
  Vector up = RandomVector(3 * n, &s), upR = RandomVector(3 * n, &s);
  auto run = [&](RBatcher& b, Vector* out, Vector* outR, Gradient* g, RGradient* rg) {
    RResultPtr r = b.BatchR(net.rv, NewRVariable(x, net.rv), n);
    *out = r->Output();
    *outR = r->ROutput();
    *g = NewGradient(vars);
    *rg = NewRGradient(vars);
    r->PropagateRGradient(up, upR, *rg, g);
  };
  Vector o1, r1, o2, r2;
  Gradient g1, g2;
  RGradient rg1, rg2;
  run(net.layers, &o1, &r1, &g1, &rg1);
  RFuncBatcher naive(net.func);  // batcher.go:57-84: Slice, ApplyR per sample, Concat
  run(naive, &o2, &r2, &g2, &rg2);
  ExpectClose(o2, o1, "output");
  ExpectClose(r2, r1, "r-output");
  for (auto& v : vars) {
    ExpectClose(g2[v.get()], g1[v.get()], "gradient");
    ExpectClose(rg2[v.get()], rg1[v.get()], "r-gradient");
  }
  // non-R: FuncBatcher vs ComposedBatcher
  ResultPtr a = net.layers.Batch(x, n), b = FuncBatcher(net.func).Batch(x, n);
  ExpectClose(b->Output(), a->Output(), "Batch output");
  Gradient ga = NewGradient(vars), gb = NewGradient(vars);
  a->PropagateGradient(up, ga);
  b->PropagateGradient(up, gb);
  for (auto& v : vars) ExpectClose(gb[v.get()], ga[v.get()], "Batch gradient");
}

// The fused layer, the whole-network plan and the operator chain are the same function (bench/mlp.go:75-83).
void TestFusedAndPlanMatchOperators() {
  std::vector<int> sizes{24, 40, 16, 10};
  const int n = 33;
  Net net = MakeNet(sizes, 5, /*biasesInRV=*/false);  // biases absent from the RVector: zero R (variable.go:85-91)
  uint64_t s = 7;
  VariablePtr x = NewVariable(RandomVector((size_t)sizes[0] * n, &s));
  Vector up = RandomVector((size_t)sizes.back() * n, &s), upR = RandomVector((size_t)sizes.back() * n, &s);
  auto run = [&](RBatcher& b, Vector* out, Vector* outR, Gradient* g, RGradient* rg) {
    RResultPtr r = b.BatchR(net.rv, NewRVariable(x, net.rv), n);
    *out = r->Output();
    *outR = r->ROutput();
    *g = NewGradient(net.vars);
    *rg = NewRGradient(net.vars);
    r->PropagateRGradient(up, upR, *rg, g);
  };
  Vector o1, r1, o2, r2;
  Gradient g1, g2;
  RGradient rg1, rg2;
  run(net.layers, &o1, &r1, &g1, &rg1);
  ComposedRBatcher fused = b200::FuseDense(net.layers);
  if (fused.size() != sizes.size() - 1) Fatalf("FuseDense did not fuse every {LinTran, LinAdd, Sigmoid} triple");
  run(fused, &o2, &r2, &g2, &rg2);
  ExpectClose(o2, o1, "fused output");
  ExpectClose(r2, r1, "fused r-output");
  for (auto& v : net.vars) {
    ExpectClose(g2[v.get()], g1[v.get()], "fused gradient");
    ExpectClose(rg2[v.get()], rg1[v.get()], "fused r-gradient");
  }
  b200::MLP plan(std::vector<int64_t>(sizes.begin(), sizes.end()), n);
  plan.SetParams(net.vars, net.rv);
  Gradient g3 = NewGradient(net.vars);
  RGradient rg3 = NewRGradient(net.vars);
  for (auto& kv : g3) kv.second.assign(kv.second.size(), 1.0);  // gradients ACCUMULATE (variable.go:45)
  auto res = plan.StepR(net.vars, x->Vector, up, upR, rg3, &g3);
  ExpectClose(res.first, o1, "plan output");
  ExpectClose(res.second, r1, "plan r-output");
  for (auto& v : net.vars) {
    Vector want = g1[v.get()];
    for (double& e : want) e += 1.0;
    ExpectClose(g3[v.get()], want, "plan gradient (accumulated onto ones)");
    ExpectClose(rg3[v.get()], rg1[v.get()], "plan r-gradient");
  }
}

// result.go:62-65: grad == nil; lin_tran.go:374-386: only map keys receive gradient; Constant prunes the recursion.
void TestNilGradientAndPruning() {
  Net net = MakeNet({6, 8, 4}, 3);
  const int n = 5;
  uint64_t s = 17;
  VariablePtr x = NewVariable(RandomVector(6 * n, &s));
  Vector up = RandomVector(4 * n, &s), upR = RandomVector(4 * n, &s);
  RResultPtr r = net.layers.BatchR(net.rv, NewRVariable(x, net.rv), n);
  Gradient g = NewGradient(net.vars);
  RGradient rg = NewRGradient(net.vars);
  r->PropagateRGradient(up, upR, rg, &g);
  RGradient rgOnly = NewRGradient({net.vars[2], net.vars[3]});  // last layer only
  r = net.layers.BatchR(net.rv, NewRVariable(x, net.rv), n);
  int64_t before = b200::Context::Default().LaunchCount();
  r->PropagateRGradient(up, upR, rgOnly, nullptr);
  int64_t launches = b200::Context::Default().LaunchCount() - before;
  if (rgOnly.size() != 2) Fatalf("PropagateRGradient added keys to the map");
  ExpectClose(rgOnly[net.vars[2].get()], rg[net.vars[2].get()], "nil-grad r-gradient W2");
  ExpectClose(rgOnly[net.vars[3].get()], rg[net.vars[3].get()], "nil-grad r-gradient b2");
  // sigmoid backward + column sums + weight gradient of layer 2, two accumulator memsets: the first layer is constant
  // w.r.t. {W2, b2}, so neither an input gradient nor anything of layer 1 may run.
  if (launches > 6) Fatalf("pruned backward pass launched " + std::to_string(launches) + " kernels");
  RGradient none;
  if (!r->Constant(none, nullptr) || r->Constant(rgOnly, nullptr) || r->Constant(none, &g) ||
      r->Constant(NewRGradient({x}), nullptr))
    Fatalf("Constant() is wrong");
}

// lin_tran.go:26-27,52-54; batcher.go:41; arithmetic.go:265: the reference's panic texts
void TestPanics() {
  LinTranFixture f;
  ExpectPanic("input length should be 4 but got 8", [&] { f.mat1->Apply(f.vec2); });
  ExpectPanic("input length should be 12 but got 8", [&] { f.mat1->Batch(f.vec2, 3); });
  ExpectPanic("input length should be 4 but got 8", [&] { f.mat1->ApplyR(f.rv, NewRVariable(f.vec2, f.rv)); });
  ExpectPanic("input count does not divide input length", [&] { Sigmoid().Batch(f.vec2, 3); });
  ExpectPanic("vector sizes do not match", [&] { Add(f.vec, f.vec2); });
  ExpectPanic("vector sizes do not match", [&] { MulR(NewRVariable(f.vec, f.rv), NewRVariable(f.vec2, f.rv)); });
  ExpectPanic("invalid upstream size", [&] {
    Gradient g = NewGradient(f.vars);
    f.mat1->Apply(f.vec)->PropagateGradient({1, 2}, g);
  });
}

void TestForeignHostResultInChain() {
  LinTranFixture f;
  Gradient g1 = NewGradient(f.vars), g2 = NewGradient(f.vars);
  ResultPtr a = f.mat2->Apply(New<HostDouble>(f.mat1->Apply(f.vec)));
  ResultPtr b = f.mat2->Apply(Scale(f.mat1->Apply(f.vec), 2));
  if (a->Output() != b->Output()) Fatalf("outputs differ: " + Show(a->Output()) + " vs " + Show(b->Output()));
  a->PropagateGradient({1}, g1);
  b->PropagateGradient({1}, g2);
  for (auto& v : f.vars) ExpectClose(g1[v.get()], g2[v.get()], "gradient through a host node");
}

// ---- seqfunc/test/map_test.go:11-51 fixtures -------------------------------------------------------------------
struct SeqFixture {
  VariablePtr tran = NewVariable({0.35771, 0.87735, 0.82339, -0.11739, -0.56811, 0.56496, 0.41425, 0.14371, -0.82190,
                                  0.24597, -0.23267, 0.50681, -0.82619, -0.59787, -0.76291, -0.96163});
  std::vector<VariablePtr> vars{NewVariable({1, -3, 2, 0.5}), NewVariable({3, -3, 0, 0.5}), NewVariable({2, -3, 2, -0.5}),
                                NewVariable({0, -3, -5, 0.5}), NewVariable({-2, 0.5, 3}), tran};
  seqfunc::VarSeqs seqs{{vars[0], vars[1], vars[2]}, {vars[2], vars[0], vars[3]}, {vars[1]}, {vars[3]},
                        {vars[0], vars[3]}};
  RVector rv{{vars[0].get(), {-1, 1, 0.33, -0.33}},
             {vars[3].get(), {2, 0.33, -5, 0.33}},
             {vars[4].get(), {1, 2, 3}},
             {tran.get(),
              {-0.31895, -0.47222, 0.58130, 0.54394, -0.99254, -0.42368, -0.56713, 0.15739, 0.28701, 0.92360, -0.33617,
               -0.42437, -0.56957, -0.97235, 0.66386, 0.22420}}};
  std::shared_ptr<LinTran> linTran = New<LinTran>(tran, 4, 4);
};

void TestMapBatcherOutput() {  // seqfunc/test/map_test.go:88-120: the batched map equals the per-vector map
  SeqFixture f;
  seqfunc::MapBatcher mb(f.linTran);
  const seqfunc::Seqs actual = mb.ApplySeqs(New<seqfunc::VarResult>(f.seqs))->OutputSeqs();
  if (actual.size() != f.seqs.size()) Fatalf("lane count differs");
  for (size_t i = 0; i < f.seqs.size(); i++) {
    if (actual[i].size() != f.seqs[i].size()) Fatalf("sequence length differs");
    for (size_t j = 0; j < f.seqs[i].size(); j++)
      ExpectClose(actual[i][j], f.linTran->Apply(f.seqs[i][j])->Output(), "mapped vector", 0, 1e-5);
  }
}

// seqfunc/test/map_test.go:122-142 (TestMapBatcher / TestMapRBatcher) by the property the checker would establish:
// outputs, R-outputs, gradients and R-gradients of the batched map equal the sums over the per-vector applications.
void TestMapRBatcher() {
  SeqFixture f;
  auto net = New<ComposedRBatcher>(ComposedRBatcher{f.linTran, New<Sigmoid>()});
  seqfunc::MapRBatcher mb(net);
  seqfunc::RResultPtr res = mb.ApplySeqsR(f.rv, New<seqfunc::VarRResult>(f.rv, f.seqs));
  uint64_t s = 5;
  seqfunc::Seqs up, upR;
  for (auto& seq : f.seqs) {
    up.emplace_back();
    upR.emplace_back();
    for (size_t j = 0; j < seq.size(); j++) {
      up.back().push_back(RandomVector(4, &s));
      upR.back().push_back(RandomVector(4, &s));
    }
  }
  Gradient g = NewGradient(f.vars), gWant = NewGradient(f.vars);
  RGradient rg = NewRGradient(f.vars), rgWant = NewRGradient(f.vars);
  res->PropagateRGradient(up, upR, rg, &g);
  if (g.size() != f.vars.size() || rg.size() != f.vars.size()) Fatalf("pool variables were left in the maps");
  for (size_t i = 0; i < f.seqs.size(); i++)
    for (size_t j = 0; j < f.seqs[i].size(); j++) {
      RResultPtr one = net->BatchR(f.rv, NewRVariable(f.seqs[i][j], f.rv), 1);
      ExpectClose(res->OutputSeqs()[i][j], one->Output(), "output");
      ExpectClose(res->ROutputSeqs()[i][j], one->ROutput(), "r-output");
      one->PropagateRGradient(up[i][j], upR[i][j], rgWant, &gWant);
    }
  for (auto& v : f.vars) {
    ExpectClose(g[v.get()], gWant[v.get()], "gradient", 1e-10, 1e-12);
    ExpectClose(rg[v.get()], rgWant[v.get()], "r-gradient", 1e-10, 1e-12);
  }
  // nil Gradient (map_batcher.go:120-123) and the non-R map
  RGradient rg2 = NewRGradient(f.vars);
  mb.ApplySeqsR(f.rv, New<seqfunc::VarRResult>(f.rv, f.seqs))->PropagateRGradient(up, upR, rg2, nullptr);
  for (auto& v : f.vars) ExpectClose(rg2[v.get()], rgWant[v.get()], "nil-grad r-gradient", 1e-10, 1e-12);
  Gradient g3 = NewGradient(f.vars);
  mb.ApplySeqs(New<seqfunc::VarResult>(f.seqs))->PropagateGradient(up, g3);
  for (auto& v : f.vars) ExpectClose(g3[v.get()], gWant[v.get()], "non-R gradient", 1e-10, 1e-12);
}

// a sequence list packed time-major in ONE variable, two maps chained device to device (the second
// map reads the first one's packed output: no host list in between), gradients for the packed variable and the shared
// parameter against the per-vector applications.
void TestPackedSeqsDevicePath() {
  SeqFixture f;
  seqfunc::Seqs host, hostR;
  for (auto& seq : f.seqs) {
    host.emplace_back();
    hostR.emplace_back();
    for (auto& v : seq) {
      host.back().push_back(v->Vector);
      auto it = f.rv.find(v.get());
      hostR.back().push_back(it != f.rv.end() ? it->second : Vector(v->Vector.size(), 0.0));
    }
  }
  std::vector<size_t> lens;
  std::vector<int64_t> widths;
  Vector flat, flatR;
  seqfunc::PackedSeqs::PackHost(hostR, &lens, &widths, &flatR);
  seqfunc::PackedSeqs::PackHost(host, &lens, &widths, &flat);
  VariablePtr packedVar = NewVariable(flat);
  RVector rv = f.rv;
  rv[packedVar.get()] = flatR;
  auto net1 = New<ComposedRBatcher>(ComposedRBatcher{f.linTran, New<Sigmoid>()});
  auto net2 = New<ComposedRBatcher>(ComposedRBatcher{f.linTran, New<Tanh>()});
  auto input = New<seqfunc::PackedVarRResult>(rv, packedVar, lens, widths);
  seqfunc::RResultPtr mid = seqfunc::MapRBatcher(net1).ApplySeqsR(rv, input);
  seqfunc::RResultPtr res = seqfunc::MapRBatcher(net2).ApplySeqsR(rv, mid);
  if (res->packed()->data->n != (int64_t)flat.size() || !res->packed()->SameShape(*input->packed())) Fatalf("packed output shape");
  uint64_t s = 9;
  seqfunc::Seqs up, upR;
  for (auto& seq : f.seqs) {
    up.emplace_back();
    upR.emplace_back();
    for (size_t j = 0; j < seq.size(); j++) {
      up.back().push_back(RandomVector(4, &s));
      upR.back().push_back(RandomVector(4, &s));
    }
  }
  std::vector<VariablePtr> vars{packedVar, f.tran};
  Gradient g = NewGradient(vars);
  RGradient rg = NewRGradient(vars);
  res->PropagateRGradient(up, upR, rg, &g);
  if (g.size() != 2 || rg.size() != 2) Fatalf("pool variables were left in the maps");
  // the same thing vector by vector
  seqfunc::Seqs wantG(f.seqs.size()), wantRG(f.seqs.size());
  Gradient gT = NewGradient({f.tran});
  RGradient rgT = NewRGradient({f.tran});
  for (size_t i = 0; i < f.seqs.size(); i++)
    for (size_t j = 0; j < f.seqs[i].size(); j++) {
      VariablePtr x = NewVariable(host[i][j]);
      RVector rvx = f.rv;
      rvx[x.get()] = hostR[i][j];
      RResultPtr one = net2->BatchR(rvx, net1->BatchR(rvx, NewRVariable(x, rvx), 1), 1);
      ExpectClose(res->OutputSeqs()[i][j], one->Output(), "output");
      ExpectClose(res->ROutputSeqs()[i][j], one->ROutput(), "r-output");
      Gradient gx = NewGradient({x, f.tran});
      RGradient rgx = NewRGradient({x, f.tran});
      one->PropagateRGradient(up[i][j], upR[i][j], rgx, &gx);
      wantG[i].push_back(gx[x.get()]);
      wantRG[i].push_back(rgx[x.get()]);
      for (size_t k = 0; k < 16; k++) gT[f.tran.get()][k] += gx[f.tran.get()][k], rgT[f.tran.get()][k] += rgx[f.tran.get()][k];
    }
  Vector wantFlat, wantFlatR;
  seqfunc::PackedSeqs::PackHost(wantG, &lens, &widths, &wantFlat);
  seqfunc::PackedSeqs::PackHost(wantRG, &lens, &widths, &wantFlatR);
  ExpectClose(g[packedVar.get()], wantFlat, "gradient of the packed list");
  ExpectClose(rg[packedVar.get()], wantFlatR, "r-gradient of the packed list");
  ExpectClose(g[f.tran.get()], gT[f.tran.get()], "parameter gradient");
  ExpectClose(rg[f.tran.get()], rgT[f.tran.get()], "parameter r-gradient");
  // the same backward pass with the maps in HBM (b200::DeviceGradient) and packed upstreams: nothing visits the host
  b200::DeviceGradient dg(vars), drg(vars);
  seqfunc::Packed pu = seqfunc::PackedSeqs::FromHost(up), puR = seqfunc::PackedSeqs::FromHost(upR);
  seqfunc::PropagateRGradient(*res, *pu, *puR, drg, &dg);
  ExpectClose(dg.Host(packedVar), wantFlat, "device gradient of the packed list");
  ExpectClose(drg.Host(f.tran), rgT[f.tran.get()], "device parameter r-gradient");
  // non-R map over the same packed variable, host upstream lists
  Gradient g2 = NewGradient(vars);
  seqfunc::ResultPtr r2 = seqfunc::MapBatcher(net2).ApplySeqs(
      seqfunc::MapBatcher(net1).ApplySeqs(New<seqfunc::PackedVarResult>(packedVar, lens, widths)));
  r2->PropagateGradient(up, g2);
  ExpectClose(g2[packedVar.get()], wantFlat, "non-R gradient of the packed list");
  ExpectClose(g2[f.tran.get()], gT[f.tran.get()], "non-R parameter gradient");
  ExpectPanic("upstream sequence shape mismatch", [&] { r2->PropagateGradient({{Vector(4, 0.0)}}, g2); });
}

// gradient.go:11-57 with the maps in HBM: same numbers as the host maps, AddToVars without a download
void TestDeviceGradient() {
  Net net = MakeNet({6, 9, 4}, 21);
  const int n = 5;
  uint64_t s = 31;
  VariablePtr x = NewVariable(RandomVector(6 * n, &s));
  Vector up = RandomVector(4 * n, &s), upR = RandomVector(4 * n, &s);
  Gradient g = NewGradient(net.vars);
  RGradient rg = NewRGradient(net.vars);
  net.layers.BatchR(net.rv, NewRVariable(x, net.rv), n)->PropagateRGradient(up, upR, rg, &g);
  b200::DeviceGradient dg(net.vars), drg(net.vars);
  RResultPtr r = net.layers.BatchR(net.rv, NewRVariable(x, net.rv), n);
  b200::PropagateRGradient(*r, up, upR, drg, &dg);
  b200::PropagateRGradient(*r, up, upR, drg, &dg);  // accumulates (variable.go:45)
  for (auto& v : net.vars) {
    Vector want = g[v.get()], wantR = rg[v.get()];
    for (double& e : want) e *= 2;
    for (double& e : wantR) e *= 2;
    ExpectClose(dg.Host(v), want, "device gradient");
    ExpectClose(drg.Host(v), wantR, "device r-gradient");
    if (!dg.Keys()[v.get()].empty()) Fatalf("a device gradient pass wrote to the host map");
  }
  dg.Scale(0.5);
  std::vector<Vector> before;
  for (auto& v : net.vars) before.push_back(v->Vector);
  dg.AddToVars(-0.1);
  ResultPtr after = net.layers.Batch(x, n);  // uses the updated device copies
  for (size_t i = 0; i < net.vars.size(); i++) {
    net.vars[i]->Sync();
    Vector want = before[i];
    for (size_t k = 0; k < want.size(); k++) want[k] += -0.1 * g[net.vars[i].get()][k];
    ExpectClose(net.vars[i]->Vector, want, "AddToVars on the device");
  }
  std::vector<VariablePtr> fresh;
  for (auto& v : net.vars) fresh.push_back(NewVariable(v->Vector));
  ComposedRBatcher again;
  for (size_t l = 0; l < 2; l++) {
    again.push_back(New<LinTran>(fresh[2 * l], l == 0 ? 9 : 4, l == 0 ? 6 : 9));
    again.push_back(New<LinAdd>(fresh[2 * l + 1]));
    again.push_back(New<Sigmoid>());
  }
  ExpectClose(after->Output(), again.Batch(x, n)->Output(), "forward pass after the device update");
  dg.Zero();
  b200::PropagateGradient(*after, up, dg);
  Gradient g2 = NewGradient(fresh);
  again.Batch(x, n)->PropagateGradient(up, g2);
  for (size_t i = 0; i < fresh.size(); i++) ExpectClose(dg.Host(net.vars[i]), g2[fresh[i].get()], "non-R device gradient");
}

// af_graph_*: an explicit-buffer ABI sequence recorded once and replayed (bench/lintran.go's step at n = 10 is bound by
// launch gaps).  The CPU restatement of the ABI records nothing (AF_ENOTSUP): the test is then vacuous.
void TestRecordedGraph() {
  using b200::Dev;
  using b200::DevVec;
  const int rows = 6, cols = 8, n = 3;
  uint64_t s = 3;
  Vector w = RandomVector(rows * cols, &s), x = RandomVector(cols * n, &s), u = RandomVector(rows * n, &s);
  Dev W = DevVec::Upload(w), X = DevVec::Upload(x), U = DevVec::Upload(u);
  Dev Y = DevVec::Zeros(rows * n), gW = DevVec::Zeros(rows * cols), D = DevVec::Zeros(cols * n);
  auto step = [&] {
    b200::check(af_lintran_fwd(b200::ctx(), W->p, X->p, Y->p, rows, cols, n));
    b200::check(af_lintran_wgrad(b200::ctx(), U->p, X->p, gW->p, rows, cols, n));
    b200::check(af_lintran_igrad(b200::ctx(), U->p, W->p, D->p, rows, cols, n));
  };
  step();  // eager warm-up; gW holds one step
  Vector y1 = Y->Download(), g1 = gW->Download(), d1 = D->Download();
  std::unique_ptr<b200::Graph> g;
  try {
    g.reset(new b200::Graph());
  } catch (const Panic& p) {
    if (p.code == AF_ENOTSUP) return;
    throw;
  }
  int64_t before = b200::Context::Default().LaunchCount();
  step();
  g->End();
  if (b200::Context::Default().LaunchCount() != before) Fatalf("recording counted launches");
  ExpectClose(gW->Download(), g1, "recording must not execute");
  g->Launch();
  g->Launch();
  if (b200::Context::Default().LaunchCount() != before + 2 * g->Launches() || g->Launches() < 3) Fatalf("replay launch count");
  Vector want = g1;
  for (double& e : want) e *= 3;
  ExpectClose(gW->Download(), want, "replayed weight gradient accumulates");
  ExpectClose(Y->Download(), y1, "replayed output");
  ExpectClose(D->Download(), d1, "replayed input gradient");
  b200::Graph g2;  // refused calls leave the recording abandonable
  ExpectPanic("af_sync: not allowed between af_graph_begin and af_graph_end", [&] { b200::Context::Default().Sync(); });
}

struct NamedTest {
  const char* name;
  void (*fn)();
};
#define T(x) {#x, x}
const NamedTest kTests[] = {
    T(TestLinTranOutput), T(TestLinTran), T(TestLinTranBatchOutput), T(TestLinTranBatch), T(TestLinAdd), T(TestExp),
    T(TestLog), T(TestSquaredNorm), T(TestSigmoid), T(TestLogSigmoid), T(TestSoftmax), T(TestSoftmax3), T(TestSin),
    T(TestTanh), T(TestConcat), T(TestSlice), T(TestWrappedSlice), T(TestRepeat), T(TestConcatBatched), T(TestArithmetic),
    T(TestAddLogDomainOutput), T(TestSumAllLogDomainOutput), T(TestCosOutput), T(TestNormOutput), T(TestNorm),
    T(TestDiv), T(TestSquaredError), T(TestPool), T(TestPoolPropagatesOnce), T(TestPoolAllEdgeCases), T(TestFoldOut),
    T(TestFold), T(TestMatMulVecOutput), T(TestMatMulVecChecks), T(TestMatMulVecsComputedMatrix),
    T(TestVariableSerialize), T(TestRFuncBatcherMatchesBatch), T(TestFusedAndPlanMatchOperators),
    T(TestNilGradientAndPruning), T(TestPanics), T(TestForeignHostResultInChain), T(TestMapBatcherOutput),
    T(TestMapRBatcher), T(TestPackedSeqsDevicePath), T(TestDeviceGradient), T(TestRecordedGraph),
};
#undef T

int SelfTest(const std::string& filter) {
  int failed = 0, ran = 0;
  for (const NamedTest& t : kTests) {
    if (!filter.empty() && std::string(t.name).find(filter) == std::string::npos) continue;
    ran++;
    try {
      t.fn();
      std::printf("PASS %s\n", t.name);
    } catch (const std::exception& e) {
      failed++;
      std::printf("FAIL %s: %s\n", t.name, e.what());
    }
  }
  std::printf("%d tests, %d failed; %lld kernel launches\n", ran, failed,
              (long long)b200::Context::Default().LaunchCount());
  return failed == 0 && ran > 0 ? 0 : 1;
}

// ---- run mode: seeded inputs from pytest, outputs back for the oracle comparison -------------------------------
std::vector<Vector> ReadVectors(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::string bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  std::vector<Vector> out;
  size_t pos = 0;
  while (pos < bytes.size()) out.push_back(Variable::Deserialize(bytes, &pos)->Vector);
  return out;
}
void WriteVectors(const std::string& path, const std::vector<Vector>& vs) {
  std::ofstream out(path, std::ios::binary);
  for (const Vector& v : vs) {
    std::string s = Variable(v).Serialize();
    out.write(s.data(), (std::streamsize)s.size());
  }
  if (!out) throw std::runtime_error("cannot write " + path);
}

// mlp <mode> <n> <s0> ... <sL>   in: W0 b0 ... | RW0 Rb0 ... | X U UR    out: out rout g... rg...
std::vector<Vector> RunMLP(const std::vector<std::string>& a, const std::vector<Vector>& in) {
  std::string mode = a.at(0);
  int n = std::stoi(a.at(1));
  std::vector<int64_t> sizes;
  for (size_t i = 2; i < a.size(); i++) sizes.push_back(std::stoll(a[i]));
  size_t P = 2 * (sizes.size() - 1);
  if (in.size() != 2 * P + 3) throw std::runtime_error("mlp: expected 2P+3 input vectors");
  std::vector<VariablePtr> vars;
  RVector rv;
  ComposedRBatcher layers;
  for (size_t i = 0; i < P; i++) {
    vars.push_back(NewVariable(in[i]));
    rv[vars[i].get()] = in[P + i];
  }
  for (size_t l = 0; l + 1 < sizes.size(); l++) {
    layers.push_back(New<LinTran>(vars[2 * l], (int)sizes[l + 1], (int)sizes[l]));
    layers.push_back(New<LinAdd>(vars[2 * l + 1]));
    layers.push_back(New<Sigmoid>());
  }
  VariablePtr x = NewVariable(in[2 * P]);
  const Vector &up = in[2 * P + 1], &upR = in[2 * P + 2];
  Gradient g = NewGradient(vars);
  RGradient rg = NewRGradient(vars);
  std::vector<Vector> out;
  if (mode == "plan") {
    b200::MLP plan(sizes, n);
    plan.SetParams(vars, rv);
    auto r = plan.StepR(vars, x->Vector, up, upR, rg, &g);
    out = {r.first, r.second};
  } else if (mode == "nonr") {
    ResultPtr r = layers.Batch(x, n);
    out = {r->Output()};
    r->PropagateGradient(up, g);
    for (auto& v : vars) out.push_back(g[v.get()]);
    return out;
  } else {
    ComposedRBatcher net = mode == "fused" ? b200::FuseDense(layers) : layers;
    RResultPtr r = net.BatchR(rv, NewRVariable(x, rv), n);
    out = {r->Output(), r->ROutput()};
    r->PropagateRGradient(up, upR, rg, mode == "nilgrad" ? nullptr : &g);
  }
  for (auto& v : vars) out.push_back(g[v.get()]);
  for (auto& v : vars) out.push_back(rg[v.get()]);
  return out;
}

// lintran <rows> <cols> <n>   (bench/lintran.go:60-90: the input is a variable too)
// in: W RW X RX U UR    out: Y RY gW rgW gX rgX
std::vector<Vector> RunLinTran(const std::vector<std::string>& a, const std::vector<Vector>& in) {
  int rows = std::stoi(a.at(0)), cols = std::stoi(a.at(1)), n = std::stoi(a.at(2));
  VariablePtr w = NewVariable(in.at(0)), x = NewVariable(in.at(2));
  RVector rv{{w.get(), in.at(1)}, {x.get(), in.at(3)}};
  LinTran lt(w, rows, cols);
  RResultPtr r = lt.BatchR(rv, NewRVariable(x, rv), n);
  Gradient g = NewGradient({w, x});
  RGradient rg = NewRGradient({w, x});
  std::vector<Vector> out{r->Output(), r->ROutput()};
  r->PropagateRGradient(in.at(4), in.at(5), rg, &g);
  out.insert(out.end(), {g[w.get()], rg[w.get()], g[x.get()], rg[x.get()]});
  return out;
}

// lstm <I> <H> <O> <T> <B>   in: 10 params | 10 R-vectors | inputs upstreams upstreamsR   out: out rout 10 g 10 rg
std::vector<Vector> RunLSTM(const std::vector<std::string>& a, const std::vector<Vector>& in) {
  b200::LSTM net(std::stoll(a.at(0)), std::stoll(a.at(1)), std::stoll(a.at(2)), std::stoll(a.at(3)), std::stoll(a.at(4)));
  std::vector<VariablePtr> vars;
  RVector rv;
  for (int i = 0; i < 10; i++) {
    vars.push_back(NewVariable(in.at(i)));
    rv[vars[i].get()] = in.at(10 + i);
  }
  net.SetParams(vars, rv);
  Gradient g = NewGradient(vars);
  RGradient rg = NewRGradient(vars);
  auto r = net.RunR(vars, in.at(20), in.at(21), in.at(22), rg, &g);
  std::vector<Vector> out{r.first, r.second};
  for (auto& v : vars) out.push_back(g[v.get()]);
  for (auto& v : vars) out.push_back(rg[v.get()]);
  return out;
}

// lstmops <I> <H> <O> <T> <B>: bench/lstm.go's network assembled from the drop-in's OPERATORS (lstmBlock.Apply
// :139-151, lstmNet.StepTime :163-196), batched over B lanes the way a Batcher sees them, R-operator carried through;
// BPTT by propagating every step's output through the whole graph.  Same files as `lstm`, so the oracle's LSTMNet and
// the fused plan are the references.   in: 10 params | 10 R-vectors | inputs upstreams upstreamsR
std::vector<Vector> RunLSTMOps(const std::vector<std::string>& a, const std::vector<Vector>& in) {
  const int I = std::stoi(a.at(0)), H = std::stoi(a.at(1)), O = std::stoi(a.at(2)), T = std::stoi(a.at(3)), B = std::stoi(a.at(4));
  const int C = H + I;
  std::vector<VariablePtr> p;  // W_in b_in W_ingate b_ingate W_forget b_forget W_outgate b_outgate W_out b_out
  RVector rv;
  for (int i = 0; i < 10; i++) {
    p.push_back(NewVariable(in.at(i)));
    rv[p[i].get()] = in.at(10 + i);
  }
  Sigmoid sig;
  auto gate = [&](int k, RResultPtr x, int rows) {
    return sig.BatchR(rv, LinAdd(p[2 * k + 1]).BatchR(rv, LinTran(p[2 * k], rows, C).BatchR(rv, x, B), B), B);
  };
  RResultPtr state = NewRVariable(NewVariable(Vector((size_t)B * H, 0.0)), rv);  // lstm.go:175, zero R
  std::vector<RResultPtr> results;
  for (int t = 0; t < T; t++) {
    Vector xt(in.at(20).begin() + (long)t * B * I, in.at(20).begin() + (long)(t + 1) * B * I);
    RResultPtr sample = NewRVariable(NewVariable(xt), rv);
    RResultPtr joined = ConcatBatchedR(B, {state, sample});                                       // :140,188
    RResultPtr inState = gate(0, joined, H), inMask = gate(1, joined, H), forgetMask = gate(2, joined, H);
    RResultPtr outState = AddR(MulR(forgetMask, state), MulR(inMask, inState));                    // :147-150
    RResultPtr masked = MulR(outState, gate(3, joined, H));                                        // :189-191
    results.push_back(gate(4, ConcatBatchedR(B, {masked, sample}), O));                            // :193-194
    state = outState;
  }
  Gradient g = NewGradient(p);
  RGradient rg = NewRGradient(p);
  Vector out, rout;
  for (int t = 0; t < T; t++) {
    out.insert(out.end(), results[t]->Output().begin(), results[t]->Output().end());
    rout.insert(rout.end(), results[t]->ROutput().begin(), results[t]->ROutput().end());
    Vector up(in.at(21).begin() + (long)t * B * O, in.at(21).begin() + (long)(t + 1) * B * O);
    Vector upR(in.at(22).begin() + (long)t * B * O, in.at(22).begin() + (long)(t + 1) * B * O);
    results[t]->PropagateRGradient(up, upR, rg, &g);
  }
  std::vector<Vector> res{out, rout};
  for (auto& v : p) res.push_back(g[v.get()]);
  for (auto& v : p) res.push_back(rg[v.get()]);
  return res;
}

// arith <n> <m> <k>: a composite over Add/Repeat/Sigmoid/Exp/Scale/Mul/Sub/Square/SumAll/Slice/Concat
// in: X (n*m) B (m) RX RB U UR (1 + k)   out: out rout gX gB rgX rgB
std::vector<Vector> RunArith(const std::vector<std::string>& a, const std::vector<Vector>& in) {
  int n = std::stoi(a.at(0)), k = std::stoi(a.at(2));
  VariablePtr x = NewVariable(in.at(0)), b = NewVariable(in.at(1));
  RVector rv{{x.get(), in.at(2)}, {b.get(), in.at(3)}};
  RResultPtr rx = NewRVariable(x, rv), rb = NewRVariable(b, rv);
  RResultPtr h = AddR(rx, RepeatR(rb, n));
  RResultPtr gte = MulR(Sigmoid().ApplyR(rv, h), Exp().ApplyR(rv, ScaleR(rx, 0.5)));
  RResultPtr o = ConcatR({SumAllR(SquareR(SubR(gte, rx))), SliceR(gte, 1, 1 + k)});
  Gradient g = NewGradient({x, b});
  RGradient rg = NewRGradient({x, b});
  std::vector<Vector> out{o->Output(), o->ROutput()};
  o->PropagateRGradient(in.at(4), in.at(5), rg, &g);
  out.insert(out.end(), {g[x.get()], g[b.get()], rg[x.get()], rg[b.get()]});
  return out;
}

// next <rows> <cols> <n>: a composite over SURVEY 8(f) rank 3 -- MatMulVecs with a computed matrix, Pool, Div, batched
// Softmax, Fold, Inverse, the squared-error head, Norm, the log-domain sums, Pow, AddFirst, ScaleFirst, Cos.
// in: M (rows*cols) V (cols*n) B (rows) RM RV RB U UR (2*rows + 1)   out: out rout gM gV gB rgM rgV rgB
std::vector<Vector> RunNext(const std::vector<std::string>& a, const std::vector<Vector>& in) {
  int rows = std::stoi(a.at(0)), cols = std::stoi(a.at(1)), n = std::stoi(a.at(2));
  VariablePtr M = NewVariable(in.at(0)), V = NewVariable(in.at(1)), B = NewVariable(in.at(2));
  RVector rv{{M.get(), in.at(3)}, {V.get(), in.at(4)}, {B.get(), in.at(5)}};
  RResultPtr rM = NewRVariable(M, rv), rV = NewRVariable(V, rv), rB = NewRVariable(B, rv);
  RResultPtr y = MatMulVecsR(AddScalerR(SquareR(rM), 1), rows, cols, rV);
  RResultPtr p = PoolR(y, [](RResultPtr q) { return DivR(q, AddScalerR(SquareR(q), 1)); });
  RResultPtr sm = Softmax().BatchR(rv, p, n);
  RResultPtr folded = FoldR(rB, SplitR(n, sm), [](RResultPtr state, RResultPtr x) {
    return MulR(InverseR(AddScalerR(SquareR(state), 1)), AddR(x, state));
  });
  RResultPtr cost = SquaredErrorR(folded, rB), nrm = Norm().ApplyR(rv, p), lse = SumAllLogDomainR(folded);
  RResultPtr pw = PowR(AddScalerR(SquareR(folded), 1), 0.3);
  RResultPtr o = ConcatR({ScaleFirstR(AddFirstR(pw, nrm), lse), cost, Cos().ApplyR(rv, AddLogDomainR(folded, rB))});
  Gradient g = NewGradient({M, V, B});
  RGradient rg = NewRGradient({M, V, B});
  std::vector<Vector> out{o->Output(), o->ROutput()};
  o->PropagateRGradient(in.at(6), in.at(7), rg, &g);
  if (g.size() != 3 || rg.size() != 3) throw std::runtime_error("temporary variables were left in the gradient maps");
  out.insert(out.end(), {g[M.get()], g[V.get()], g[B.get()], rg[M.get()], rg[V.get()], rg[B.get()]});
  return out;
}

}  // namespace

int main(int argc, char** argv) {
  std::vector<std::string> args(argv + 1, argv + argc);
  try {
    if (!args.empty() && args[0] == "selftest") return SelfTest(args.size() > 1 ? args[1] : "");
    if (args.size() >= 4 && args[0] == "run") {
      std::vector<Vector> in = ReadVectors(args[2]);
      std::vector<std::string> rest(args.begin() + 4, args.end());
      std::vector<Vector> out;
      if (args[1] == "mlp") out = RunMLP(rest, in);
      else if (args[1] == "lintran") out = RunLinTran(rest, in);
      else if (args[1] == "lstm") out = RunLSTM(rest, in);
      else if (args[1] == "lstmops") out = RunLSTMOps(rest, in);
      else if (args[1] == "arith") out = RunArith(rest, in);
      else if (args[1] == "next") out = RunNext(rest, in);
      else throw std::runtime_error("unknown case " + args[1]);
      WriteVectors(args[3], out);
      std::printf("run ok: %lld kernel launches\n", (long long)b200::Context::Default().LaunchCount());
      return 0;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "host_tests: %s\n", e.what());
    return 2;
  }
  std::fprintf(stderr, "usage: host_tests selftest [filter] | run <case> <in> <out> <args...>\n");
  return 64;
}
