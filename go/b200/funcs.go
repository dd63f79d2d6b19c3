// SURVEY 8(f) rows through cgo: the elementwise functions, the batched Softmax, the squared-error cost head and the
// device-resident parameter update, on top of the entry points declared in include/autofunc_b200.h after af_fill.
//
// Like b200.go this file is NOT compiled in the build image (no Go toolchain there); it transliterates
// autofunc_b200/funcs.py, which the GPU parity tests exercise (tests/test_gpu_next.py).  The combinators of the
// reference that are pure graph logic -- Pool (pool.go), Fold (fold.go), MatMulVecs (lin_alg.go), the log-domain sums
// (arithmetic.go:1049-1090) -- need no cgo of their own: they compose device nodes through the Result/RResult
// interfaces and work unchanged on top of these types.
package b200

/*
#include "autofunc_b200.h"
*/
import "C"

import (
	"github.com/unixpickle/autofunc"
	"github.com/unixpickle/num-analysis/linalg"
)

type ewKind int

const (
	kindLog        ewKind = iota // math_funcs.go:99-189
	kindLogSigmoid               // math_funcs.go:376-477
	kindSin                      // math_funcs.go:511-597
	kindTanh                     // extension named by the north star (the reference has no Tanh)
)

// ewFunc is an elementwise autofunc.RFunc (and, being shape-agnostic, an RBatcher).  funcs.py: _EwFunc.
type ewFunc struct{ kind ewKind }

var (
	Log        = ewFunc{kindLog}
	LogSigmoid = ewFunc{kindLogSigmoid}
	Sin        = ewFunc{kindSin}
	Tanh       = ewFunc{kindTanh}
)

func (f ewFunc) fwd(x, xR, o, oR *devVec) {
	n := C.int64_t(x.n)
	switch f.kind {
	case kindLog:
		check(C.af_log_fwd_r(context(), x.p, ptr(xR), o.p, ptr(oR), n))
	case kindLogSigmoid:
		check(C.af_logsigmoid_fwd_r(context(), x.p, ptr(xR), o.p, ptr(oR), n))
	case kindSin:
		check(C.af_sin_fwd_r(context(), x.p, ptr(xR), o.p, ptr(oR), n))
	case kindTanh:
		check(C.af_tanh_fwd_r(context(), x.p, ptr(xR), o.p, ptr(oR), n))
	}
}

// bwd overwrites the upstreams in place, as the reference does (result.go:24-29).
func (f ewFunc) bwd(src, srcR, u, uR *devVec) {
	n := C.int64_t(u.n)
	switch f.kind {
	case kindLog:
		check(C.af_log_bwd_r(context(), src.p, ptr(srcR), u.p, ptr(uR), n))
	case kindLogSigmoid:
		check(C.af_logsigmoid_bwd_r(context(), src.p, ptr(srcR), u.p, ptr(uR), n))
	case kindSin:
		check(C.af_sin_bwd_r(context(), src.p, ptr(srcR), u.p, ptr(uR), n))
	case kindTanh: // reads the forward OUTPUTS (T, RT)
		check(C.af_tanh_bwd_r(context(), src.p, ptr(srcR), u.p, ptr(uR), n))
	}
}

func (f ewFunc) ApplyR(_ autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	x, xR := inputDevR(in)
	o, oR := newDev(x.n), newDev(x.n)
	f.fwd(x, xR, o, oR)
	return &ewRResult{f: f, in: in, x: x, xR: xR, o: o, oR: oR}
}

func (f ewFunc) BatchR(rv autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	if len(in.Output())%n != 0 {
		panic("input count does not divide input length") // batcher.go:41
	}
	return f.ApplyR(rv, in)
}

type ewRResult struct {
	f            ewFunc
	in           autofunc.RResult
	x, xR, o, oR *devVec
}

func (r *ewRResult) dev() *devVec            { return r.o }
func (r *ewRResult) devR() *devVec           { return r.oR }
func (r *ewRResult) Output() linalg.Vector   { return r.o.download() }
func (r *ewRResult) ROutput() linalg.Vector  { return r.oR.download() }
func (r *ewRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g)
}
func (r *ewRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}
func (r *ewRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	if r.in.Constant(rg, g) {
		return
	}
	if r.f.kind == kindTanh {
		r.f.bwd(r.o, r.oR, u, uR)
	} else {
		r.f.bwd(r.x, r.xR, u, uR)
	}
	downR(r.in, s, u, uR, rg, g)
}

// ---------------------------------------------------------------------------------------------
// Softmax: drop-in for RFuncBatcher{&autofunc.Softmax{T}} (math_funcs.go:479-509, batcher.go:57-84):
// one softmax per sample, ONE kernel per direction.  funcs.py: Softmax, _SoftmaxResult.
// ---------------------------------------------------------------------------------------------
type Softmax struct{ Temperature float64 }

func (s *Softmax) ApplyR(rv autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	return s.BatchR(rv, in, 1)
}

func (s *Softmax) BatchR(_ autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	x, xR := inputDevR(in)
	if n < 1 || x.n%n != 0 {
		panic("input count does not divide input length")
	}
	y, yR := newDev(x.n), newDev(x.n)
	check(C.af_softmax_fwd_r(context(), x.p, ptr(xR), C.double(s.Temperature), y.p, yR.p, C.int64_t(x.n/n), C.int64_t(n)))
	return &softmaxRResult{t: s.Temperature, in: in, y: y, yR: yR, rows: n}
}

type softmaxRResult struct {
	t     float64
	in    autofunc.RResult
	y, yR *devVec
	rows  int
}

func (r *softmaxRResult) dev() *devVec           { return r.y }
func (r *softmaxRResult) devR() *devVec          { return r.yR }
func (r *softmaxRResult) Output() linalg.Vector  { return r.y.download() }
func (r *softmaxRResult) ROutput() linalg.Vector { return r.yR.download() }
func (r *softmaxRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g)
}
func (r *softmaxRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}
func (r *softmaxRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	if r.in.Constant(rg, g) {
		return
	}
	check(C.af_softmax_bwd_r(context(), r.y.p, r.yR.p, C.double(r.t), u.p, uR.p, C.int64_t(r.y.n/r.rows), C.int64_t(r.rows)))
	downR(r.in, s, u, uR, rg, g)
}

// ---------------------------------------------------------------------------------------------
// SquaredErrorR: 0.5*||a - y||^2, the reference's ScaleR(SquaredNorm{}.ApplyR(rv, SubR(a, y)), 0.5)
// (math_funcs.go:191-199, arithmetic.go:144,466) as one reduction.  funcs.py: squared_error_r.
// ---------------------------------------------------------------------------------------------
func SquaredErrorR(actual, expected autofunc.RResult) autofunc.RResult {
	a, aR := inputDevR(actual)
	y, yR := inputDevR(expected)
	if a.n != y.n {
		panic("vector sizes do not match") // arithmetic.go:265
	}
	c, cR := newDev(1), newDev(1)
	check(C.af_sqerr_fwd_r(context(), a.p, ptr(aR), y.p, ptr(yR), c.p, cR.p, C.int64_t(a.n)))
	return &sqErrRResult{actual: actual, expected: expected, a: a, aR: aR, y: y, yR: yR, c: c, cR: cR}
}

type sqErrRResult struct {
	actual, expected   autofunc.RResult
	a, aR, y, yR, c, cR *devVec
}

func (r *sqErrRResult) dev() *devVec           { return r.c }
func (r *sqErrRResult) devR() *devVec          { return r.cR }
func (r *sqErrRResult) Output() linalg.Vector  { return r.c.download() }
func (r *sqErrRResult) ROutput() linalg.Vector { return r.cR.download() }
func (r *sqErrRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.actual.Constant(rg, g) && r.expected.Constant(rg, g)
}
func (r *sqErrRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}
func (r *sqErrRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	for _, side := range []struct {
		in   autofunc.RResult
		sign float64
	}{{r.actual, 1}, {r.expected, -1}} {
		if side.in.Constant(rg, g) {
			continue
		}
		d, dR := newDev(r.a.n), newDev(r.a.n)
		check(C.af_sqerr_bwd_r(context(), r.a.p, ptr(r.aR), r.y.p, ptr(r.yR), u.p, uR.p, C.double(side.sign),
			d.p, dR.p, C.int64_t(r.a.n)))
		downR(side.in, s, d, dR, rg, g)
	}
}

// ---------------------------------------------------------------------------------------------
// MLP plan: one bench iteration in one call, and Gradient.AddToVars without leaving HBM
// ---------------------------------------------------------------------------------------------

// StepFresh is bench/mlp.go:43-57 with fresh Gradient maps (:33-35) and outGrad = 1, outRGrad = 0 (:118-126);
// results stay on the device (af_mlp_grad_ptr / af_mlp_output_ptr).  api.py: MLPPlan.step_fresh.
func (m *MLP) StepFresh(withR bool) {
	r := C.int(0)
	if withR {
		r = 1
	}
	check(C.af_mlp_step_fresh(m.h, r))
}

// AddToVars is Gradient.AddToVars (gradient.go:40-52) on the plan's own parameters: param += scale * grad, in HBM.
func (m *MLP) AddToVars(scale float64) { check(C.af_mlp_add_to_vars(m.h, C.double(scale))) }

// Param downloads parameter i (W0, b0, W1, b1, ...) after device-side updates.
func (m *MLP) Param(i int, dst linalg.Vector) {
	check(C.af_mlp_get_param(m.h, C.int(i), hostPtr(dst), C.int64_t(len(dst))))
}
