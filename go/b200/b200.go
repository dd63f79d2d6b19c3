// Package b200 is the cgo shim that makes libautofunc_b200.so a drop-in for autofunc's dense
// differentiation path.  Its types implement autofunc.Func / RFunc (func.go:7-27), autofunc.Batcher /
// RBatcher (batcher.go:5-29) and return autofunc.Result / RResult (result.go:10-77) nodes whose values
// live in HBM, on top of the C ABI in include/autofunc_b200.h.  Nothing here computes on the CPU:
// every forward and backward method is a call into the library.
//
// NOTE: no Go toolchain exists in the build image, so this package is NOT compiled or tested there.
// It is kept 1:1 with the header (tests/test_abi_cpu.py checks every C.af_* call against the
// prototypes: name and argument count) and with autofunc_b200/api.py, the tested Python mirror of
// the same logic; every method cites the api.py routine it transliterates.
//
// Files: b200.go (context, device vectors, backward session, LinTran, DenseSigmoid, MLP plan,
// recorded graphs), ops.go (LinAdd, Sigmoid, Exp, Add/Mul/Scale), funcs.go (SURVEY 8f functions), lstm.go (the LSTM plan).
//
// Build (on a machine with Go >= 1.21, CUDA 12.9 and a B200):
//   CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/autofunc_b200 -lautofunc_b200" go build ./go/b200
package b200

/*
#include <stdlib.h>
#include "autofunc_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"

	"github.com/unixpickle/autofunc"
	"github.com/unixpickle/num-analysis/linalg"
)

// ctx is the process-wide context (device 0, own stream); api.py: context().  Like the reference,
// the package is single-goroutine (SURVEY 8b "Threading"): calls are serialised by the caller.
var ctx *C.af_ctx

func context() *C.af_ctx {
	if ctx == nil {
		if rc := C.af_ctx_create(0, &ctx); rc != 0 {
			panic("autofunc_b200: " + C.GoString(C.af_last_error()))
		}
	}
	return ctx
}

func check(rc C.int) {
	if rc != 0 {
		// AF_ESHAPE carries the reference's own panic text (lin_tran.go:26-27, arithmetic.go:265).
		panic(C.GoString(C.af_last_error()))
	}
}

// hostPtr is &x[0] as a C pointer, or NULL for an empty vector.
func hostPtr(x linalg.Vector) *C.double {
	if len(x) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&x[0]))
}

// devVec is a float64 vector resident in HBM (api.py: DeviceVec).
type devVec struct {
	p *C.double
	n int
}

func newDev(n int) *devVec {
	v := &devVec{n: n}
	check(C.af_malloc(context(), C.int64_t(n), &v.p))
	runtime.SetFinalizer(v, func(v *devVec) { C.af_free(context(), v.p) })
	return v
}

func upload(x linalg.Vector) *devVec {
	v := newDev(len(x))
	if len(x) > 0 {
		check(C.af_upload(context(), v.p, hostPtr(x), C.int64_t(len(x))))
	}
	return v
}

func (v *devVec) download() linalg.Vector {
	out := make(linalg.Vector, v.n)
	if v.n > 0 {
		check(C.af_download(context(), hostPtr(out), v.p, C.int64_t(v.n)))
	}
	return out
}

func (v *devVec) copy() *devVec {
	c := newDev(v.n)
	check(C.af_copy(context(), c.p, v.p, C.int64_t(v.n)))
	return c
}

func ptr(v *devVec) *C.double {
	if v == nil {
		return nil // NULL R operand == identically zero (variable.go:85-91)
	}
	return v.p
}

// zerosIfNil materialises an absent (identically zero) R operand for the entry points that need one.
func zerosIfNil(v *devVec, n int) *devVec {
	if v != nil {
		return v
	}
	z := newDev(n)
	check(C.af_zero(context(), z.p, C.int64_t(n)))
	return z
}

// deviceNode / deviceRNode are implemented by every Result / RResult this package returns, so that device
// nodes recognise each other (as slices.go:148 recognises *Variable) and hand device buffers down the
// recursion instead of host vectors.
type deviceNode interface {
	dev() *devVec
	bwd(s *session, u *devVec, g autofunc.Gradient)
}
type deviceRNode interface {
	dev() *devVec
	devR() *devVec // nil == identically zero
	bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient)
}

// hostOut caches the downloaded Output / ROutput of a device node: the reference's Output() returns the same
// aliased slice on every call (result.go:12-15).
type hostOut struct {
	out, outR linalg.Vector
}

func (h *hostOut) output(d *devVec) linalg.Vector {
	if h.out == nil {
		h.out = d.download()
	}
	return h.out
}

func (h *hostOut) routput(d *devVec, n int) linalg.Vector {
	if h.outR == nil {
		if d == nil {
			h.outR = make(linalg.Vector, n)
		} else {
			h.outR = d.download()
		}
	}
	return h.outR
}

// session holds device accumulators for parameter gradients; flush adds them (+=) into the caller's
// maps before the outermost Propagate(R)Gradient returns, which keeps the temp-variable protocol of
// pool.go:114-136 working (api.py: _Session).  One backward pass sees exactly one Gradient and one
// RGradient map, so an accumulator is identified by (which map, variable).
type accKey struct {
	isR bool
	v   *autofunc.Variable
}
type sessionEntry struct {
	dst linalg.Vector
	buf *devVec
}
type session struct {
	acc map[accKey]*sessionEntry
}

func newSession() *session { return &session{acc: map[accKey]*sessionEntry{}} }

func (s *session) accumulator(key accKey, dst linalg.Vector) *devVec {
	if e, ok := s.acc[key]; ok {
		return e.buf
	}
	b := newDev(len(dst))
	check(C.af_zero(context(), b.p, C.int64_t(b.n)))
	s.acc[key] = &sessionEntry{dst: dst, buf: b}
	return b
}

// gradAcc / rgradAcc return the device accumulator for v, or nil when v receives no gradient: the map is nil
// (result.go:62-65) or v is not one of its keys (variable.go:43-52).
func (s *session) gradAcc(g autofunc.Gradient, v *autofunc.Variable) *devVec {
	if g == nil {
		return nil
	}
	dst, ok := g[v]
	if !ok {
		return nil
	}
	return s.accumulator(accKey{false, v}, dst)
}

func (s *session) rgradAcc(rg autofunc.RGradient, v *autofunc.Variable) *devVec {
	dst, ok := rg[v]
	if !ok {
		return nil
	}
	return s.accumulator(accKey{true, v}, dst)
}

func (s *session) flush() {
	for _, e := range s.acc {
		e.dst.Add(e.buf.download())
	}
	s.acc = map[accKey]*sessionEntry{}
}

// inputDev / inputDevR: the device image of an input node (api.py: _in_d, _in_dr).
func inputDev(in autofunc.Result) *devVec {
	if d, ok := in.(deviceNode); ok {
		return d.dev()
	}
	return upload(in.Output())
}

func inputDevR(in autofunc.RResult) (x, xR *devVec) {
	if d, ok := in.(deviceRNode); ok {
		return d.dev(), d.devR()
	}
	// host node (an *autofunc.RVariable or a foreign RResult); a variable absent from the RVector has a
	// fresh zero ROutput (variable.go:85-91), uploaded as such
	return upload(in.Output()), upload(in.ROutput())
}

// down / downR continue the recursion into an input: device nodes get device upstreams; a host node gets the
// accumulators flushed first (it may read the maps: pool.go:114-136) and host copies of the upstreams.
func down(in autofunc.Result, s *session, u *devVec, g autofunc.Gradient) {
	if d, ok := in.(deviceNode); ok {
		d.bwd(s, u, g)
		return
	}
	s.flush()
	in.PropagateGradient(u.download(), g)
}

func downR(in autofunc.RResult, s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	if d, ok := in.(deviceRNode); ok {
		d.bwdR(s, u, uR, rg, g)
		return
	}
	s.flush()
	in.PropagateRGradient(u.download(), uR.download(), rg, g)
}

// propagate / propagateR are the host entry points every node's PropagateGradient / PropagateRGradient
// forwards to (api.py: _DeviceResult.propagate_gradient / propagate_rgradient).
func propagate(n deviceNode, u linalg.Vector, g autofunc.Gradient) {
	s := newSession()
	n.bwd(s, upload(u), g)
	s.flush()
}

func propagateR(n deviceRNode, u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	s := newSession()
	n.bwdR(s, upload(u), upload(uR), rg, g)
	s.flush()
}

// rOperand uploads a variable and, when the RVector holds one, its R-vector (nil otherwise: the product is skipped).
func rOperand(x *autofunc.Variable, v autofunc.RVector) (d, dR *devVec) {
	d = upload(x.Vector)
	if r, ok := v[x]; ok {
		dR = upload(r)
	}
	return
}

// ---------------------------------------------------------------------------------------------
// LinTran: drop-in for autofunc.LinTran (lin_tran.go:17-425).  api.py: LinTran, _LinTranResult,
// _LinTranRResult.
// ---------------------------------------------------------------------------------------------
type LinTran struct {
	Data *autofunc.Variable
	Rows int
	Cols int
}

func (l *LinTran) checkLen(got, n int) {
	if l.Cols*n != got { // lin_tran.go:26-27,52-54
		panic(fmt.Sprintf("input length should be %d but got %d", l.Cols*n, got))
	}
}

func (l *LinTran) Apply(in autofunc.Result) autofunc.Result { return l.Batch(in, 1) }

func (l *LinTran) ApplyR(v autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	return l.BatchR(v, in, 1)
}

func (l *LinTran) Batch(in autofunc.Result, n int) autofunc.Result {
	x := inputDev(in)
	l.checkLen(x.n, n)
	w := upload(l.Data.Vector)
	y := newDev(n * l.Rows)
	check(C.af_lintran_fwd(context(), w.p, x.p, y.p, C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(n)))
	return &linTranResult{m: l, in: in, w: w, x: x, n: n, y: y}
}

func (l *LinTran) BatchR(v autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	x, xR := inputDevR(in)
	l.checkLen(x.n, n)
	w, rw := rOperand(l.Data, v)
	y, yR := newDev(n*l.Rows), newDev(n*l.Rows)
	check(C.af_lintran_fwd_r(context(), w.p, ptr(rw), x.p, ptr(xR), y.p, yR.p,
		C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(n)))
	return &linTranRResult{m: l, in: in, rdata: autofunc.NewRVariable(l.Data, v), w: w, rw: rw, x: x, xR: xR, n: n, y: y, yR: yR}
}

type linTranResult struct {
	hostOut
	m    *LinTran
	in   autofunc.Result
	w, x *devVec
	n    int
	y    *devVec
}

func (r *linTranResult) dev() *devVec          { return r.y }
func (r *linTranResult) Output() linalg.Vector { return r.output(r.y) }
func (r *linTranResult) Constant(g autofunc.Gradient) bool {
	return r.m.Data.Constant(g) && r.in.Constant(g)
}
func (r *linTranResult) PropagateGradient(u linalg.Vector, g autofunc.Gradient) { propagate(r, u, g) }

// lin_tran.go:373-382
func (r *linTranResult) bwd(s *session, u *devVec, g autofunc.Gradient) {
	m := r.m
	if gW := s.gradAcc(g, m.Data); gW != nil {
		check(C.af_lintran_wgrad(context(), u.p, r.x.p, gW.p, C.int64_t(m.Rows), C.int64_t(m.Cols), C.int64_t(r.n)))
	}
	if !r.in.Constant(g) {
		d := newDev(r.n * m.Cols)
		check(C.af_lintran_igrad(context(), u.p, r.w.p, d.p, C.int64_t(m.Rows), C.int64_t(m.Cols), C.int64_t(r.n)))
		down(r.in, s, d, g)
	}
}

type linTranRResult struct {
	hostOut
	m            *LinTran
	in           autofunc.RResult
	rdata        *autofunc.RVariable
	w, rw, x, xR *devVec
	n            int
	y, yR        *devVec
}

func (r *linTranRResult) dev() *devVec           { return r.y }
func (r *linTranRResult) devR() *devVec          { return r.yR }
func (r *linTranRResult) Output() linalg.Vector  { return r.output(r.y) }
func (r *linTranRResult) ROutput() linalg.Vector { return r.routput(r.yR, r.y.n) }
func (r *linTranRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g) && r.rdata.Constant(rg, g)
}
func (r *linTranRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}

// lin_tran.go:408-425
func (r *linTranRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	m := r.m
	gW, rgW := s.gradAcc(g, m.Data), s.rgradAcc(rg, m.Data)
	if r.in.Constant(rg, g) {
		if gW != nil || rgW != nil {
			check(C.af_lintran_wgrad_r(context(), u.p, uR.p, r.x.p, ptr(r.xR), ptr(gW), ptr(rgW),
				C.int64_t(m.Rows), C.int64_t(m.Cols), C.int64_t(r.n)))
		}
	} else {
		// both gradients from one call: with 9..16 samples the library streams W once for the two of them
		d, dR := newDev(r.n*m.Cols), newDev(r.n*m.Cols)
		check(C.af_lintran_backward_r(context(), u.p, uR.p, r.w.p, ptr(r.rw), r.x.p, ptr(r.xR), ptr(gW), ptr(rgW), d.p, dR.p,
			C.int64_t(m.Rows), C.int64_t(m.Cols), C.int64_t(r.n), C.int(0)))
		downR(r.in, s, d, dR, rg, g)
	}
}

// ---------------------------------------------------------------------------------------------
// DenseSigmoid: one Batcher / RBatcher for the triple {&LinTran{W}, &LinAdd{b}, Sigmoid{}} of
// bench/mlp.go:75-83; one fused kernel per direction.  api.py: DenseSigmoid, _DenseResult.
// ---------------------------------------------------------------------------------------------
type DenseSigmoid struct {
	Weights, Biases *autofunc.Variable
	Rows, Cols      int
}

func (l *DenseSigmoid) checkLen(got, n int) {
	if l.Cols*n != got {
		panic(fmt.Sprintf("input length should be %d but got %d", l.Cols*n, got))
	}
}

func (l *DenseSigmoid) Apply(in autofunc.Result) autofunc.Result { return l.Batch(in, 1) }
func (l *DenseSigmoid) ApplyR(v autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	return l.BatchR(v, in, 1)
}

func (l *DenseSigmoid) Batch(in autofunc.Result, n int) autofunc.Result {
	x := inputDev(in)
	l.checkLen(x.n, n)
	w, b := upload(l.Weights.Vector), upload(l.Biases.Vector)
	s := newDev(n * l.Rows)
	// d_RS == NULL selects the non-R forward
	check(C.af_dense_fwd(context(), w.p, nil, b.p, nil, x.p, nil, s.p, nil,
		C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(n), 1))
	return &denseResult{l: l, in: in, w: w, x: x, n: n, s: s}
}

func (l *DenseSigmoid) BatchR(v autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	x, xR := inputDevR(in)
	l.checkLen(x.n, n)
	w, rw := rOperand(l.Weights, v)
	b, rb := rOperand(l.Biases, v)
	s, sR := newDev(n*l.Rows), newDev(n*l.Rows)
	check(C.af_dense_fwd(context(), w.p, ptr(rw), b.p, ptr(rb), x.p, ptr(xR), s.p, sR.p,
		C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(n), 1))
	return &denseRResult{l: l, in: in, rv: v, w: w, rw: rw, x: x, xR: xR, n: n, s: s, sR: sR}
}

type denseResult struct {
	hostOut
	l    *DenseSigmoid
	in   autofunc.Result
	w, x *devVec
	n    int
	s    *devVec
}

func (r *denseResult) dev() *devVec          { return r.s }
func (r *denseResult) Output() linalg.Vector { return r.output(r.s) }
func (r *denseResult) Constant(g autofunc.Gradient) bool {
	return r.in.Constant(g) && r.l.Weights.Constant(g) && r.l.Biases.Constant(g)
}
func (r *denseResult) PropagateGradient(u linalg.Vector, g autofunc.Gradient) { propagate(r, u, g) }
func (r *denseResult) bwd(s *session, u *devVec, g autofunc.Gradient)         { r.bwdFrom(s, u, g, false) }

// skipAct: the layer above already applied this layer's sigmoid backward in its igrad epilogue.
func (r *denseResult) bwdFrom(s *session, u *devVec, g autofunc.Gradient, skipAct bool) {
	if r.Constant(g) {
		return
	}
	l, c := r.l, context()
	if !skipAct { // math_funcs.go:326-333, in place
		check(C.af_sigmoid_bwd(c, r.s.p, u.p, C.int64_t(u.n)))
	}
	if gb := s.gradAcc(g, l.Biases); gb != nil { // lin_add.go:66-69, batched
		check(C.af_bias_grad(c, u.p, gb.p, C.int64_t(l.Rows), C.int64_t(r.n)))
	}
	if gw := s.gradAcc(g, l.Weights); gw != nil { // lin_tran.go:374-376
		check(C.af_lintran_wgrad(c, u.p, r.x.p, gw.p, C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(r.n)))
	}
	if !r.in.Constant(g) { // lin_tran.go:378-381 (+ fused sigmoid backward of the layer below)
		d := newDev(r.n * l.Cols)
		prev, fuse := r.in.(*denseResult)
		var ps *C.double
		if fuse {
			ps = prev.s.p
		}
		check(C.af_dense_igrad(c, u.p, nil, r.w.p, nil, ps, nil, d.p, nil,
			C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(r.n)))
		if fuse {
			prev.bwdFrom(s, d, g, true)
		} else {
			down(r.in, s, d, g)
		}
	}
}

type denseRResult struct {
	hostOut
	l            *DenseSigmoid
	in           autofunc.RResult
	rv           autofunc.RVector
	w, rw, x, xR *devVec
	n            int
	s, sR        *devVec
}

func (r *denseRResult) dev() *devVec           { return r.s }
func (r *denseRResult) devR() *devVec          { return r.sR }
func (r *denseRResult) Output() linalg.Vector  { return r.output(r.s) }
func (r *denseRResult) ROutput() linalg.Vector { return r.routput(r.sR, r.s.n) }
func (r *denseRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g) && autofunc.NewRVariable(r.l.Weights, r.rv).Constant(rg, g) &&
		autofunc.NewRVariable(r.l.Biases, r.rv).Constant(rg, g)
}
func (r *denseRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	propagateR(r, u, uR, rg, g)
}
func (r *denseRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	r.bwdRFrom(s, u, uR, rg, g, false)
}

func (r *denseRResult) bwdRFrom(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient, skipAct bool) {
	if r.Constant(rg, g) {
		return
	}
	l, c := r.l, context()
	if !skipAct { // math_funcs.go:357-374, in place
		check(C.af_sigmoid_bwd_r(c, r.s.p, r.sR.p, u.p, uR.p, C.int64_t(u.n)))
	}
	gb, rgb := s.gradAcc(g, l.Biases), s.rgradAcc(rg, l.Biases)
	if gb != nil || rgb != nil { // lin_add.go:94-104, batched
		check(C.af_bias_grad_r(c, u.p, uR.p, ptr(gb), ptr(rgb), C.int64_t(l.Rows), C.int64_t(r.n)))
	}
	gw, rgw := s.gradAcc(g, l.Weights), s.rgradAcc(rg, l.Weights)
	if gw != nil || rgw != nil { // lin_tran.go:410-418
		check(C.af_lintran_wgrad_r(c, u.p, uR.p, r.x.p, ptr(r.xR), ptr(gw), ptr(rgw),
			C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(r.n)))
	}
	if !r.in.Constant(rg, g) { // lin_tran.go:420-423 (+ fused sigmoid backward of the layer below)
		d, dR := newDev(r.n*l.Cols), newDev(r.n*l.Cols)
		prev, fuse := r.in.(*denseRResult)
		var ps, psR *C.double
		if fuse {
			ps, psR = prev.s.p, prev.sR.p
		}
		check(C.af_dense_igrad(c, u.p, uR.p, r.w.p, ptr(r.rw), ps, psR, d.p, dR.p,
			C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(r.n)))
		if fuse {
			prev.bwdRFrom(s, d, dR, rg, g, true)
		} else {
			downR(r.in, s, d, dR, rg, g)
		}
	}
}

// ---------------------------------------------------------------------------------------------
// MLP: the whole-network plan (bench/mlp.go:24-126, batched).  api.py: MLPPlan.
// ---------------------------------------------------------------------------------------------
type MLP struct {
	h     *C.af_mlp
	Sizes []int
	Batch int
}

func NewMLP(sizes []int, batch int) *MLP {
	cs := make([]C.int64_t, len(sizes))
	for i, s := range sizes {
		cs[i] = C.int64_t(s)
	}
	m := &MLP{Sizes: sizes, Batch: batch}
	check(C.af_mlp_create(context(), &cs[0], C.int(len(sizes)-1), C.int64_t(batch), &m.h))
	runtime.SetFinalizer(m, func(m *MLP) { C.af_mlp_destroy(m.h) })
	return m
}

// SetParams uploads variables in bench/mlp.go:84-85 order (W0, b0, W1, b1, ...) and their R-vectors.
func (m *MLP) SetParams(vars []*autofunc.Variable, rv autofunc.RVector) {
	for i, v := range vars {
		check(C.af_mlp_set_param(m.h, C.int(i), hostPtr(v.Vector), C.int64_t(len(v.Vector))))
		if r, ok := rv[v]; ok {
			check(C.af_mlp_set_rvector(m.h, C.int(i), hostPtr(r), C.int64_t(len(r))))
		} else {
			check(C.af_mlp_set_rvector(m.h, C.int(i), nil, 0))
		}
	}
}

// cPtrArray is a C-allocated array of n host pointers: cgo forbids passing Go memory that itself holds Go
// pointers, so the pointer arrays of af_mlp_step_host live in C memory and the vectors they name are pinned.
type cPtrArray struct {
	base **C.double
	n    int
}

func newCPtrArray(n int) cPtrArray {
	p := C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof((*C.double)(nil))))
	return cPtrArray{base: (**C.double)(p), n: n}
}

func (a cPtrArray) set(i int, p *C.double) {
	unsafe.Slice(a.base, a.n)[i] = p
}

func (a cPtrArray) free() { C.free(unsafe.Pointer(a.base)) }

// StepR is one iteration of bench/mlp.go:52-54 from zeroed accumulators; gradients are ADDED into
// grad / rgrad (variable.go:45) in the order of vars.  outGrad / outRGrad may be nil: ones / zeros, as
// bench/mlp.go:118-126 builds them.
func (m *MLP) StepR(vars []*autofunc.Variable, input, outGrad, outRGrad linalg.Vector,
	rgrad autofunc.RGradient, grad autofunc.Gradient) (out, outR linalg.Vector) {
	nOut := m.Batch * m.Sizes[len(m.Sizes)-1]
	out, outR = make(linalg.Vector, nOut), make(linalg.Vector, nOut)
	tmpG := make([]linalg.Vector, len(vars))
	tmpRG := make([]linalg.Vector, len(vars))
	nParams := 2 * (len(m.Sizes) - 1)
	if len(vars) != nParams {
		panic(fmt.Sprintf("b200.MLP: expected %d variables (W0, b0, W1, b1, ...) but got %d", nParams, len(vars)))
	}
	pg, prg := newCPtrArray(nParams), newCPtrArray(nParams) // the library reads 2L entries; NULL entries are skipped
	defer pg.free()
	defer prg.free()
	var pin runtime.Pinner
	defer pin.Unpin()
	for i, v := range vars {
		tmpG[i] = make(linalg.Vector, len(v.Vector))
		tmpRG[i] = make(linalg.Vector, len(v.Vector))
		if len(v.Vector) > 0 {
			pin.Pin(&tmpG[i][0])
			pin.Pin(&tmpRG[i][0])
		}
		pg.set(i, hostPtr(tmpG[i]))
		prg.set(i, hostPtr(tmpRG[i]))
	}
	check(C.af_mlp_step_host(m.h, 1, hostPtr(input), hostPtr(outGrad), hostPtr(outRGrad),
		hostPtr(out), hostPtr(outR), pg.base, prg.base))
	for i, v := range vars {
		if g, ok := grad[v]; ok {
			g.Add(tmpG[i])
		}
		if g, ok := rgrad[v]; ok {
			g.Add(tmpRG[i])
		}
	}
	return
}

// Graph is a recorded launch sequence (af_graph_*): every C.af_* launch issued between Record and End is captured
// into a CUDA graph instead of running; Launch replays the sequence -- same buffers, same sizes -- as one
// submission.  For launch-bound chains: the reference's own benchmarks run n = 1 (bench/mlp.go:36) and n = 10
// (bench/lintran.go:13-17).  Run the sequence once eagerly first.
type Graph struct{ h *C.af_graph }

func Record() *Graph {
	check(C.af_graph_begin(context()))
	return &Graph{}
}

func (g *Graph) End()          { check(C.af_graph_end(context(), &g.h)) }
func (g *Graph) Launch()       { check(C.af_graph_launch(g.h)) }
func (g *Graph) Launches() int { return int(C.af_graph_launches(g.h)) }
func (g *Graph) Close()        { C.af_graph_destroy(g.h); g.h = nil }
