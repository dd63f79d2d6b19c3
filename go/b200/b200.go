// Package b200 is the cgo shim that makes libautofunc_b200.so a drop-in for autofunc's dense
// differentiation path.  It implements autofunc.RFunc / autofunc.RBatcher (func.go:24-27,
// batcher.go:24-29) and autofunc.Result / autofunc.RResult (result.go:10-77) on top of the C ABI
// in include/autofunc_b200.h.
//
// NOTE: no Go toolchain exists in the build image, so this file is NOT compiled or tested there.
// It is kept 1:1 with the header and with autofunc_b200/api.py (the tested Python mirror of the
// same logic); every method cites the api.py routine it transliterates.
//
// Build (on a machine with Go, CUDA 12.9 and a B200):
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

// ctx is the process-wide context (device 0, own stream); api.py: context().
var ctx *C.af_ctx

func context() *C.af_ctx {
	if ctx == nil {
		runtime.LockOSThread() // cgo calls block this goroutine's OS thread; keep the CUDA context on it
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
		check(C.af_upload(context(), v.p, (*C.double)(unsafe.Pointer(&x[0])), C.int64_t(len(x))))
	}
	return v
}

func (v *devVec) download() linalg.Vector {
	out := make(linalg.Vector, v.n)
	if v.n > 0 {
		check(C.af_download(context(), (*C.double)(unsafe.Pointer(&out[0])), v.p, C.int64_t(v.n)))
	}
	return out
}

func ptr(v *devVec) *C.double {
	if v == nil {
		return nil // NULL R operand == identically zero (variable.go:85-91)
	}
	return v.p
}

// deviceResult is implemented by every Result this package returns, so that device nodes recognise
// each other (as slices.go:148 recognises *Variable) and hand device buffers down the recursion.
type deviceResult interface {
	dev() *devVec
	devR() *devVec
	bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient)
}

// session holds device accumulators for parameter gradients; flush adds them (+=) into the caller's
// maps before the outermost PropagateRGradient returns, which keeps the temp-variable protocol of
// pool.go:114-136 working (api.py: _Session).
type session struct {
	acc map[[2]uintptr]*sessionEntry
}
type sessionEntry struct {
	dst linalg.Vector
	buf *devVec
}

func (s *session) accumulator(m map[*autofunc.Variable]linalg.Vector, v *autofunc.Variable) *devVec {
	key := [2]uintptr{uintptr(unsafe.Pointer(&m)), uintptr(unsafe.Pointer(v))}
	if e, ok := s.acc[key]; ok {
		return e.buf
	}
	b := newDev(len(v.Vector))
	check(C.af_zero(context(), b.p, C.int64_t(b.n)))
	s.acc[key] = &sessionEntry{dst: m[v], buf: b}
	return b
}

func (s *session) flush() {
	for _, e := range s.acc {
		e.dst.Add(e.buf.download())
	}
	s.acc = map[[2]uintptr]*sessionEntry{}
}

func inputDev(in autofunc.RResult) (x, xR *devVec) {
	if d, ok := in.(deviceResult); ok {
		return d.dev(), d.devR()
	}
	if rv, ok := in.(*autofunc.RVariable); ok {
		// a Variable absent from the RVector has a fresh zero ROutput (variable.go:85-91); the
		// caller can pass ZeroR to skip that product (api.py: RVariable._dr()).
		return upload(rv.Output()), upload(rv.ROutput())
	}
	return upload(in.Output()), upload(in.ROutput())
}

func down(in autofunc.RResult, s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	if d, ok := in.(deviceResult); ok {
		d.bwdR(s, u, uR, rg, g)
		return
	}
	s.flush()
	in.PropagateRGradient(u.download(), uR.download(), rg, g)
}

// ---------------------------------------------------------------------------------------------
// LinTran: drop-in for autofunc.LinTran (lin_tran.go:17-425).  api.py: LinTran, _LinTranRResult.
// ---------------------------------------------------------------------------------------------
type LinTran struct {
	Data *autofunc.Variable
	Rows int
	Cols int
}

func (l *LinTran) Apply(in autofunc.Result) autofunc.Result {
	// The non-R interface has the same structure with af_lintran_fwd / _wgrad / _igrad; omitted
	// here for brevity, see api.py LinTran.batch / _LinTranResult.
	return (&autofunc.LinTran{Data: l.Data, Rows: l.Rows, Cols: l.Cols}).Apply(in)
}

func (l *LinTran) Batch(in autofunc.Result, n int) autofunc.Result {
	return (&autofunc.LinTran{Data: l.Data, Rows: l.Rows, Cols: l.Cols}).Batch(in, n)
}

func (l *LinTran) ApplyR(v autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	if len(in.Output()) != l.Cols {
		panic(fmt.Sprintf("input length should be %d but got %d", l.Cols, len(in.Output())))
	}
	return l.BatchR(v, in, 1)
}

func (l *LinTran) BatchR(v autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	if l.Cols*n != len(in.Output()) {
		panic(fmt.Sprintf("input length should be %d but got %d", l.Cols*n, len(in.Output())))
	}
	w := upload(l.Data.Vector)
	var rw *devVec
	if r, ok := v[l.Data]; ok {
		rw = upload(r)
	}
	x, xR := inputDev(in)
	y, yR := newDev(n*l.Rows), newDev(n*l.Rows)
	check(C.af_lintran_fwd_r(context(), w.p, ptr(rw), x.p, ptr(xR), y.p, yR.p,
		C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(n)))
	return &linTranRResult{m: l, in: in, rdata: autofunc.NewRVariable(l.Data, v), w: w, rw: rw, x: x, xR: xR, n: n, y: y, yR: yR}
}

type linTranRResult struct {
	m            *LinTran
	in           autofunc.RResult
	rdata        *autofunc.RVariable
	w, rw, x, xR *devVec
	n            int
	y, yR        *devVec
	out, outR    linalg.Vector
}

func (r *linTranRResult) dev() *devVec  { return r.y }
func (r *linTranRResult) devR() *devVec { return r.yR }

func (r *linTranRResult) Output() linalg.Vector {
	if r.out == nil {
		r.out = r.y.download()
	}
	return r.out
}

func (r *linTranRResult) ROutput() linalg.Vector {
	if r.outR == nil {
		r.outR = r.yR.download()
	}
	return r.outR
}

func (r *linTranRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g) && r.rdata.Constant(rg, g)
}

func (r *linTranRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	s := &session{acc: map[[2]uintptr]*sessionEntry{}}
	r.bwdR(s, upload(u), upload(uR), rg, g)
	s.flush()
}

// lin_tran.go:408-425
func (r *linTranRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	m := r.m
	var gW, rgW *devVec
	if g != nil && !m.Data.Constant(g) {
		gW = s.accumulator(g, m.Data)
	}
	if _, ok := rg[m.Data]; ok {
		rgW = s.accumulator(rg, m.Data)
	}
	if gW != nil || rgW != nil {
		check(C.af_lintran_wgrad_r(context(), u.p, uR.p, r.x.p, ptr(r.xR), ptr(gW), ptr(rgW),
			C.int64_t(m.Rows), C.int64_t(m.Cols), C.int64_t(r.n)))
	}
	if !r.in.Constant(rg, g) {
		d, dR := newDev(r.n*m.Cols), newDev(r.n*m.Cols)
		check(C.af_lintran_igrad_r(context(), u.p, uR.p, r.w.p, ptr(r.rw), d.p, dR.p,
			C.int64_t(m.Rows), C.int64_t(m.Cols), C.int64_t(r.n)))
		down(r.in, s, d, dR, rg, g)
	}
}

// ---------------------------------------------------------------------------------------------
// DenseSigmoid: one RBatcher for the triple {&LinTran{W}, &LinAdd{b}, Sigmoid{}} of
// bench/mlp.go:75-83; one fused kernel per direction.  api.py: DenseSigmoid, _DenseResult.
// ---------------------------------------------------------------------------------------------
type DenseSigmoid struct {
	Weights, Biases *autofunc.Variable
	Rows, Cols      int
}

func (l *DenseSigmoid) Apply(in autofunc.Result) autofunc.Result { return l.Batch(in, 1) }
func (l *DenseSigmoid) Batch(in autofunc.Result, n int) autofunc.Result {
	f := autofunc.ComposedFunc{&autofunc.LinTran{Data: l.Weights, Rows: l.Rows, Cols: l.Cols},
		&autofunc.LinAdd{Var: l.Biases}, autofunc.Sigmoid{}}
	b := autofunc.FuncBatcher{F: f}
	return b.Batch(in, n)
}
func (l *DenseSigmoid) ApplyR(v autofunc.RVector, in autofunc.RResult) autofunc.RResult {
	return l.BatchR(v, in, 1)
}

func (l *DenseSigmoid) BatchR(v autofunc.RVector, in autofunc.RResult, n int) autofunc.RResult {
	if l.Cols*n != len(in.Output()) {
		panic(fmt.Sprintf("input length should be %d but got %d", l.Cols*n, len(in.Output())))
	}
	up := func(x *autofunc.Variable) (*devVec, *devVec) {
		if r, ok := v[x]; ok {
			return upload(x.Vector), upload(r)
		}
		return upload(x.Vector), nil
	}
	w, rw := up(l.Weights)
	b, rb := up(l.Biases)
	x, xR := inputDev(in)
	s, sR := newDev(n*l.Rows), newDev(n*l.Rows)
	check(C.af_dense_fwd(context(), w.p, ptr(rw), b.p, ptr(rb), x.p, ptr(xR), s.p, sR.p,
		C.int64_t(l.Rows), C.int64_t(l.Cols), C.int64_t(n), 1))
	return &denseRResult{l: l, in: in, rv: v, w: w, rw: rw, x: x, xR: xR, n: n, s: s, sR: sR}
}

type denseRResult struct {
	l            *DenseSigmoid
	in           autofunc.RResult
	rv           autofunc.RVector
	w, rw, x, xR *devVec
	n            int
	s, sR        *devVec
	out, outR    linalg.Vector
}

func (r *denseRResult) dev() *devVec  { return r.s }
func (r *denseRResult) devR() *devVec { return r.sR }
func (r *denseRResult) Output() linalg.Vector {
	if r.out == nil {
		r.out = r.s.download()
	}
	return r.out
}
func (r *denseRResult) ROutput() linalg.Vector {
	if r.outR == nil {
		r.outR = r.sR.download()
	}
	return r.outR
}
func (r *denseRResult) Constant(rg autofunc.RGradient, g autofunc.Gradient) bool {
	return r.in.Constant(rg, g) && autofunc.NewRVariable(r.l.Weights, r.rv).Constant(rg, g) &&
		autofunc.NewRVariable(r.l.Biases, r.rv).Constant(rg, g)
}
func (r *denseRResult) PropagateRGradient(u, uR linalg.Vector, rg autofunc.RGradient, g autofunc.Gradient) {
	s := &session{acc: map[[2]uintptr]*sessionEntry{}}
	r.bwdR(s, upload(u), upload(uR), rg, g)
	s.flush()
}

func (r *denseRResult) bwdR(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient) {
	r.bwdRFrom(s, u, uR, rg, g, false)
}

// skipAct: the layer above already applied this layer's sigmoid backward in its igrad epilogue.
func (r *denseRResult) bwdRFrom(s *session, u, uR *devVec, rg autofunc.RGradient, g autofunc.Gradient, skipAct bool) {
	if r.Constant(rg, g) {
		return
	}
	l, c := r.l, context()
	if !skipAct { // math_funcs.go:357-374, in place
		check(C.af_sigmoid_bwd_r(c, r.s.p, r.sR.p, u.p, uR.p, C.int64_t(u.n)))
	}
	acc := func(m map[*autofunc.Variable]linalg.Vector, v *autofunc.Variable) *devVec {
		if m == nil {
			return nil
		}
		if _, ok := m[v]; !ok {
			return nil
		}
		return s.accumulator(m, v)
	}
	gb, rgb := acc(g, l.Biases), acc(rg, l.Biases)
	if gb != nil || rgb != nil { // lin_add.go:94-104, batched
		check(C.af_bias_grad_r(c, u.p, uR.p, ptr(gb), ptr(rgb), C.int64_t(l.Rows), C.int64_t(r.n)))
	}
	gw, rgw := acc(g, l.Weights), acc(rg, l.Weights)
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
			down(r.in, s, d, dR, rg, g)
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
		check(C.af_mlp_set_param(m.h, C.int(i), (*C.double)(unsafe.Pointer(&v.Vector[0])), C.int64_t(len(v.Vector))))
		if r, ok := rv[v]; ok {
			check(C.af_mlp_set_rvector(m.h, C.int(i), (*C.double)(unsafe.Pointer(&r[0])), C.int64_t(len(r))))
		} else {
			check(C.af_mlp_set_rvector(m.h, C.int(i), nil, 0))
		}
	}
}

// StepR is one iteration of bench/mlp.go:52-54 from zeroed accumulators; gradients are ADDED into
// grad / rgrad (variable.go:45) in the order of vars.
func (m *MLP) StepR(vars []*autofunc.Variable, input, outGrad, outRGrad linalg.Vector,
	rgrad autofunc.RGradient, grad autofunc.Gradient) (out, outR linalg.Vector) {
	nOut := m.Batch * m.Sizes[len(m.Sizes)-1]
	out, outR = make(linalg.Vector, nOut), make(linalg.Vector, nOut)
	tmpG := make([]linalg.Vector, len(vars))
	tmpRG := make([]linalg.Vector, len(vars))
	pg := make([]*C.double, len(vars))
	prg := make([]*C.double, len(vars))
	for i, v := range vars {
		tmpG[i] = make(linalg.Vector, len(v.Vector))
		tmpRG[i] = make(linalg.Vector, len(v.Vector))
		pg[i] = (*C.double)(unsafe.Pointer(&tmpG[i][0]))
		prg[i] = (*C.double)(unsafe.Pointer(&tmpRG[i][0]))
	}
	check(C.af_mlp_step_host(m.h, 1, (*C.double)(unsafe.Pointer(&input[0])),
		(*C.double)(unsafe.Pointer(&outGrad[0])), (*C.double)(unsafe.Pointer(&outRGrad[0])),
		(*C.double)(unsafe.Pointer(&out[0])), (*C.double)(unsafe.Pointer(&outR[0])), &pg[0], &prg[0]))
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
