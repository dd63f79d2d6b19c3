// The operators of the dense path other than LinTran, through cgo: LinAdd as a Batcher / RBatcher (the packed
// form is Add(x, Repeat(b, n)), SURVEY F4), Sigmoid and Exp as RFunc + RBatcher, and the arithmetic the LSTM
// cell of bench/lstm.go:139-196 is made of (Add, Mul, Scale, each with its R twin).
//
// Like b200.go this file is NOT compiled in the build image (no Go toolchain there); it transliterates
// autofunc_b200/api.py (LinAdd, Sigmoid, Exp, add/add_r, mul/mul_r, scale/scale_r), which the GPU parity
// tests exercise (tests/test_gpu_parity.py, tests/test_gpu_api_ops.py).
package b200

/*
#include "autofunc_b200.h"
*/
import "C"

import (
	"fmt"

	"github.com/unixpickle/autofunc"
	"github.com/unixpickle/num-analysis/linalg"
)

// ---------------------------------------------------------------------------------------------
// LinAdd: drop-in for autofunc.LinAdd (lin_add.go:9-109) that is also a Batcher / RBatcher.
// ---------------------------------------------------------------------------------------------
type LinAdd struct {
	Var *autofunc.Variable
}

func (l LinAdd) checkLen(got, n int) {
	if want := n * len(l.Var.Vector); got != want {
		panic(fmt.Sprintf("input length should be %d but got %d", want, got))
	}
}

func (l LinAdd) Apply(in autofunc.Result) autofunc.Result { return l.Batch(in, 1) }
func (l LinAdd) ApplyR(v autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	return l.BatchR(v, in, 1)
}

func (l LinAdd) Batch(in autofunc.Result, n int) autofunc.Result {
	x := inputDev(in)
	l.checkLen(x.n, n)
	b := upload(l.Var.Vector)
	y := newDev(x.n)
	check(C.af_bias_fwd(context(), x.p, b.p, y.p, C.int64_t(b.n), C.int64_t(n)))
	return &linAddResult{v: l.Var, in: in, n: n, y: y}
}

func (l LinAdd) BatchR(v autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	x, xR := inputDevR(in)
	l.checkLen(x.n, n)
	b, bR := rOperand(l.Var, v)
	y, yR := newDev(x.n), newDev(x.n)
	check(C.af_bias_fwd_r(context(), x.p, ptr(xR), b.p, ptr(bR), y.p, yR.p, C.int64_t(b.n), C.int64_t(n)))
	return &linAddRResult{rvar: autofunc.NewRVariable(l.Var, v), in: in, n: n, y: y, yR: yR}
}

type linAddResult struct {
	hostOut
	v  *autofunc.Variable
	in autofunc.Result
	n  int
	y  *devVec
}

func (r *linAddResult) dev() *devVec          { return r.y }
func (r *linAddResult) Output() linalg.Vector { return r.output(r.y) }
func (r *linAddResult) Constant(g autofunc.Gradient) bool {
	return r.in.Constant(g) && r.v.Constant(g)
}
func (r *linAddResult) PropagateGradient(u linalg.Vector, g autofunc.Gradient) { propagate(r, u, g) }

// lin_add.go:66-73; the bias gradient of a packed batch is the column sum (Repeat backward, slices.go:244-250)
func (r *linAddResult) bwd(s *session, u *devVec, g autofunc.Gradient) {
	if gb := s.gradAcc(g, r.v); gb != nil {
		check(C.af_bias_grad(context(), u.p, gb.p, C.int64_t(gb.n), C.int64_t(r.n)))
	}
	if !r.in.Constant(g) {
		down(r.in, s, u, g)
	}
}

type linAddRResult struct {
	hostOut
	rvar  *autofunc.RVariable
	in    autofunc.RResult
	n     int
	y, yR *devVec
}

func (r *linAddRResult) dev() *devVec           { return r.y }
func (r *linAddRResult) devR() *devVec          { return r.yR }
func (r *linAddRResult) Output() linalg.Vector  { return r.output(r.y) }
func (r *linAddRResult) ROutput() linalg.Vector { return r.routput(r.yR, r.y.n) }
func (r *linAddRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.rvar.Constant(rg, g) && r.in.Constant(rg, g)
}
func (r *linAddRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}

// lin_add.go:94-109
func (r *linAddRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	v := r.rvar.Variable
	gb, rgb := s.gradAcc(g, v), s.rgradAcc(rg, v)
	if gb != nil || rgb != nil {
		check(C.af_bias_grad_r(context(), u.p, uR.p, ptr(gb), ptr(rgb), C.int64_t(len(v.Vector)), C.int64_t(r.n)))
	}
	if !r.in.Constant(rg, g) {
		downR(r.in, s, u, uR, rg, g)
	}
}

// ---------------------------------------------------------------------------------------------
// Sigmoid (math_funcs.go:286-374) and Exp (math_funcs.go:12-97): RFunc + RBatcher, value and R form in one
// pass, backward in place on the upstreams.  Both backward passes read the forward OUTPUTS.
// ---------------------------------------------------------------------------------------------
type unaryKind int

const (
	kindSigmoid unaryKind = iota
	kindExp
)

type unaryFunc struct{ kind unaryKind }

var (
	Sigmoid = unaryFunc{kindSigmoid}
	Exp     = unaryFunc{kindExp}
)

func (f unaryFunc) Apply(in autofunc.Result) autofunc.Result {
	x := inputDev(in)
	o := newDev(x.n)
	switch f.kind {
	case kindSigmoid:
		check(C.af_sigmoid_fwd(context(), x.p, o.p, C.int64_t(x.n)))
	case kindExp:
		check(C.af_exp_fwd(context(), x.p, o.p, C.int64_t(x.n)))
	}
	return &unaryResult{f: f, in: in, o: o}
}

func (f unaryFunc) ApplyR(_ autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	x, xR := inputDevR(in)
	o, oR := newDev(x.n), newDev(x.n)
	switch f.kind {
	case kindSigmoid:
		check(C.af_sigmoid_fwd_r(context(), x.p, ptr(xR), o.p, oR.p, C.int64_t(x.n)))
	case kindExp:
		check(C.af_exp_fwd_r(context(), x.p, ptr(xR), o.p, oR.p, C.int64_t(x.n)))
	}
	return &unaryRResult{f: f, in: in, o: o, oR: oR}
}

func (f unaryFunc) Batch(in autofunc.Result, n int) autofunc.Result {
	if n < 1 || len(in.Output())%n != 0 {
		panic("input count does not divide input length") // batcher.go:41
	}
	return f.Apply(in)
}

func (f unaryFunc) BatchR(rv autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	if n < 1 || len(in.Output())%n != 0 {
		panic("input count does not divide input length")
	}
	return f.ApplyR(rv, in)
}

type unaryResult struct {
	hostOut
	f  unaryFunc
	in autofunc.Result
	o  *devVec
}

func (r *unaryResult) dev() *devVec                                           { return r.o }
func (r *unaryResult) Output() linalg.Vector                                  { return r.output(r.o) }
func (r *unaryResult) Constant(g autofunc.Gradient) bool                      { return r.in.Constant(g) }
func (r *unaryResult) PropagateGradient(u linalg.Vector, g autofunc.Gradient) { propagate(r, u, g) }

// math_funcs.go:326-333 (Sigmoid), :56-63 (Exp)
func (r *unaryResult) bwd(s *session, u *devVec, g autofunc.Gradient) {
	if r.in.Constant(g) {
		return
	}
	switch r.f.kind {
	case kindSigmoid:
		check(C.af_sigmoid_bwd(context(), r.o.p, u.p, C.int64_t(u.n)))
	case kindExp:
		check(C.af_exp_bwd(context(), r.o.p, u.p, C.int64_t(u.n)))
	}
	down(r.in, s, u, g)
}

type unaryRResult struct {
	hostOut
	f     unaryFunc
	in    autofunc.RResult
	o, oR *devVec
}

func (r *unaryRResult) dev() *devVec           { return r.o }
func (r *unaryRResult) devR() *devVec          { return r.oR }
func (r *unaryRResult) Output() linalg.Vector  { return r.output(r.o) }
func (r *unaryRResult) ROutput() linalg.Vector { return r.routput(r.oR, r.o.n) }
func (r *unaryRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g)
}
func (r *unaryRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}

// math_funcs.go:357-374 (Sigmoid, reference operand order), :85-97 (Exp)
func (r *unaryRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	if r.in.Constant(rg, g) {
		return
	}
	switch r.f.kind {
	case kindSigmoid:
		check(C.af_sigmoid_bwd_r(context(), r.o.p, r.oR.p, u.p, uR.p, C.int64_t(u.n)))
	case kindExp:
		check(C.af_exp_bwd_r(context(), r.o.p, r.oR.p, u.p, uR.p, C.int64_t(u.n)))
	}
	downR(r.in, s, u, uR, rg, g)
}

// ---------------------------------------------------------------------------------------------
// Add / Mul (arithmetic.go:17-96, 261-376) and Scale (arithmetic.go:436-493), with their R twins.
// ---------------------------------------------------------------------------------------------
type binaryKind int

const (
	kindAdd binaryKind = iota
	kindMul
)

func sameLen(a, b *devVec) {
	if a.n != b.n {
		panic("vector sizes do not match") // arithmetic.go:265
	}
}

func binaryFwd(kind binaryKind, a, b *devVec) *devVec {
	o := newDev(a.n)
	switch kind {
	case kindAdd:
		check(C.af_add(context(), a.p, b.p, o.p, C.int64_t(a.n)))
	case kindMul:
		check(C.af_mul(context(), a.p, b.p, o.p, C.int64_t(a.n)))
	}
	return o
}

func Add(r1, r2 autofunc.Result) autofunc.Result { return binary(kindAdd, r1, r2) }
func Mul(r1, r2 autofunc.Result) autofunc.Result { return binary(kindMul, r1, r2) }

func binary(kind binaryKind, r1, r2 autofunc.Result) autofunc.Result {
	a, b := inputDev(r1), inputDev(r2)
	sameLen(a, b)
	return &binaryResult{kind: kind, r1: r1, r2: r2, a: a, b: b, o: binaryFwd(kind, a, b)}
}

func AddR(r1, r2 autofunc.RResult) autofunc.RResult { return binaryR(kindAdd, r1, r2) }
func MulR(r1, r2 autofunc.RResult) autofunc.RResult { return binaryR(kindMul, r1, r2) }

func binaryR(kind binaryKind, r1, r2 autofunc.RResult) autofunc.RResult {
	a, aR := inputDevR(r1)
	b, bR := inputDevR(r2)
	sameLen(a, b)
	oR := newDev(a.n)
	switch kind {
	case kindAdd: // arithmetic.go:60-61
		check(C.af_add(context(), zerosIfNil(aR, a.n).p, zerosIfNil(bR, a.n).p, oR.p, C.int64_t(a.n)))
	case kindMul: // arithmetic.go:325-329: a*bR + aR*b
		check(C.af_mul_r(context(), a.p, ptr(aR), b.p, ptr(bR), oR.p, C.int64_t(a.n)))
	}
	return &binaryRResult{kind: kind, r1: r1, r2: r2, a: a, aR: aR, b: b, bR: bR, o: binaryFwd(kind, a, b), oR: oR}
}

type binaryResult struct {
	hostOut
	kind    binaryKind
	r1, r2  autofunc.Result
	a, b, o *devVec
}

func (r *binaryResult) dev() *devVec          { return r.o }
func (r *binaryResult) Output() linalg.Vector { return r.output(r.o) }
func (r *binaryResult) Constant(g autofunc.Gradient) bool {
	return r.r1.Constant(g) && r.r2.Constant(g)
}
func (r *binaryResult) PropagateGradient(u linalg.Vector, g autofunc.Gradient) { propagate(r, u, g) }

func (r *binaryResult) bwd(s *session, u *devVec, g autofunc.Gradient) {
	c1, c2 := r.r1.Constant(g), r.r2.Constant(g)
	switch r.kind {
	case kindAdd: // arithmetic.go:34-49: the first side gets a copy when the second also needs the upstream
		if !c1 {
			if c2 {
				down(r.r1, s, u, g)
			} else {
				down(r.r1, s, u.copy(), g)
			}
		}
		if !c2 {
			down(r.r2, s, u, g)
		}
	case kindMul: // arithmetic.go:288-306: D = U * other
		if !c1 {
			d := newDev(u.n)
			check(C.af_mul_bwd(context(), u.p, nil, r.b.p, nil, d.p, nil, C.int64_t(u.n)))
			down(r.r1, s, d, g)
		}
		if !c2 {
			d := newDev(u.n)
			check(C.af_mul_bwd(context(), u.p, nil, r.a.p, nil, d.p, nil, C.int64_t(u.n)))
			down(r.r2, s, d, g)
		}
	}
}

type binaryRResult struct {
	hostOut
	kind                 binaryKind
	r1, r2               autofunc.RResult
	a, aR, b, bR, o, oR  *devVec
}

func (r *binaryRResult) dev() *devVec           { return r.o }
func (r *binaryRResult) devR() *devVec          { return r.oR }
func (r *binaryRResult) Output() linalg.Vector  { return r.output(r.o) }
func (r *binaryRResult) ROutput() linalg.Vector { return r.routput(r.oR, r.o.n) }
func (r *binaryRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.r1.Constant(rg, g) && r.r2.Constant(rg, g)
}
func (r *binaryRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}

func (r *binaryRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	c1, c2 := r.r1.Constant(rg, g), r.r2.Constant(rg, g)
	switch r.kind {
	case kindAdd: // arithmetic.go:79-96
		if !c1 {
			if c2 {
				downR(r.r1, s, u, uR, rg, g)
			} else {
				downR(r.r1, s, u.copy(), uR.copy(), rg, g)
			}
		}
		if !c2 {
			downR(r.r2, s, u, uR, rg, g)
		}
	case kindMul: // arithmetic.go:349-376: D = U*o ; DR = U*oR + UR*o
		if !c1 {
			d, dR := newDev(u.n), newDev(u.n)
			check(C.af_mul_bwd(context(), u.p, uR.p, r.b.p, ptr(r.bR), d.p, dR.p, C.int64_t(u.n)))
			downR(r.r1, s, d, dR, rg, g)
		}
		if !c2 {
			d, dR := newDev(u.n), newDev(u.n)
			check(C.af_mul_bwd(context(), u.p, uR.p, r.a.p, ptr(r.aR), d.p, dR.p, C.int64_t(u.n)))
			downR(r.r2, s, d, dR, rg, g)
		}
	}
}

// Scale / ScaleR: arithmetic.go:436-493.
func Scale(in autofunc.Result, f float64) autofunc.Result {
	o := inputDev(in).copy()
	check(C.af_scale(context(), o.p, C.double(f), C.int64_t(o.n)))
	return &scaleResult{f: f, in: in, o: o}
}

func ScaleR(in autofunc.RResult, f float64) autofunc.RResult {
	x, xR := inputDevR(in)
	o, oR := x.copy(), zerosIfNil(xR, x.n).copy()
	check(C.af_scale(context(), o.p, C.double(f), C.int64_t(o.n)))
	check(C.af_scale(context(), oR.p, C.double(f), C.int64_t(oR.n)))
	return &scaleRResult{f: f, in: in, o: o, oR: oR}
}

type scaleResult struct {
	hostOut
	f  float64
	in autofunc.Result
	o  *devVec
}

func (r *scaleResult) dev() *devVec                                           { return r.o }
func (r *scaleResult) Output() linalg.Vector                                  { return r.output(r.o) }
func (r *scaleResult) Constant(g autofunc.Gradient) bool                      { return r.in.Constant(g) }
func (r *scaleResult) PropagateGradient(u linalg.Vector, g autofunc.Gradient) { propagate(r, u, g) }
func (r *scaleResult) bwd(s *session, u *devVec, g autofunc.Gradient) { // arithmetic.go:444-456
	if !r.in.Constant(g) {
		check(C.af_scale(context(), u.p, C.double(r.f), C.int64_t(u.n)))
		down(r.in, s, u, g)
	}
}

type scaleRResult struct {
	hostOut
	f     float64
	in    autofunc.RResult
	o, oR *devVec
}

func (r *scaleRResult) dev() *devVec           { return r.o }
func (r *scaleRResult) devR() *devVec          { return r.oR }
func (r *scaleRResult) Output() linalg.Vector  { return r.output(r.o) }
func (r *scaleRResult) ROutput() linalg.Vector { return r.routput(r.oR, r.o.n) }
func (r *scaleRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g)
}
func (r *scaleRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}
func (r *scaleRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) { // arithmetic.go:476-493
	if !r.in.Constant(rg, g) {
		check(C.af_scale(context(), u.p, C.double(r.f), C.int64_t(u.n)))
		check(C.af_scale(context(), uR.p, C.double(r.f), C.int64_t(uR.n)))
		downR(r.in, s, u, uR, rg, g)
	}
}
