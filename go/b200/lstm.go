// LSTM: bench/lstm.go's lstmNet (lstmBlock + output layer, bench/lstm.go:25-241) batched over B lanes, forward + BPTT
// with the R-operator carried through, as ONE resident plan (af_lstm_*).  lstm.py: LSTMPlan.
//
// Like b200.go this file is NOT compiled in the build image (no Go toolchain there); tests/test_abi_cpu.py lints every
// C.af_* call against the header, and the same entry points are driven by tests/test_gpu_lstm.py through lstm.py.
package b200

/*
#include "autofunc_b200.h"
*/
import "C"

import (
	"runtime"

	"github.com/unixpickle/autofunc"
	"github.com/unixpickle/num-analysis/linalg"
)

// LSTM holds the ten parameters of bench/lstm.go:75-99 in reference order (InWeights, InBiases, InGateWeights,
// InGateBiases, ForgetGateWeights, ForgetGateBiases, OutputGate, OutputGateBiases, OutputWeights, OutputBiases), their
// R-vectors, the activations of all T steps and the gradient / R-gradient accumulators in HBM.
type LSTM struct {
	h                                    *C.af_lstm
	Input, Hidden, Output, Steps, Lanes int
}

func NewLSTM(input, hidden, output, steps, lanes int) *LSTM {
	l := &LSTM{Input: input, Hidden: hidden, Output: output, Steps: steps, Lanes: lanes}
	check(C.af_lstm_create(context(), C.int64_t(input), C.int64_t(hidden), C.int64_t(output), C.int64_t(steps),
		C.int64_t(lanes), &l.h))
	runtime.SetFinalizer(l, func(l *LSTM) { C.af_lstm_destroy(l.h) })
	return l
}

// ParamLen is the length of parameter i: gate matrices are H x (H+I) over concat(state, x) (bench/lstm.go:140), the
// output matrix O x (H+I) over concat(masked, x) (:193), biases H or O.
func (l *LSTM) ParamLen(i int) int {
	rows := l.Hidden
	if i/2 == 4 {
		rows = l.Output
	}
	if i&1 == 1 {
		return rows
	}
	return rows * (l.Hidden + l.Input)
}

// SetParams uploads the ten variables and, for those the RVector holds, their R-vectors (absent: zero R, variable.go:85-91).
func (l *LSTM) SetParams(vars []*autofunc.Variable, rv autofunc.RVector) {
	for i, v := range vars {
		check(C.af_lstm_set_param(l.h, C.int(i), hostPtr(v.Vector), C.int64_t(len(v.Vector))))
		if r, ok := rv[v]; ok {
			check(C.af_lstm_set_rvector(l.h, C.int(i), hostPtr(r), C.int64_t(len(r))))
		} else {
			check(C.af_lstm_set_rvector(l.h, C.int(i), nil, 0))
		}
	}
}

// RunR is Reset + T x StepTime (bench/lstm.go:42-45) followed by PropagateGradient over all steps (:215-234), with the
// R-operator: inputs are T x B x I, upstreams / upstreamsR T x B x O, time-major and lane-major within a step (the
// packing seqfunc.MapBatcher builds, seqfunc/map.go:262-284).  Accumulators start from zero; the results are ADDED
// into grad / rgrad in the order of vars (variable.go:45).
func (l *LSTM) RunR(vars []*autofunc.Variable, inputs, upstreams, upstreamsR linalg.Vector,
	rgrad autofunc.RGradient, grad autofunc.Gradient) (out, outR linalg.Vector) {
	nOut := l.Steps * l.Lanes * l.Output
	out, outR = make(linalg.Vector, nOut), make(linalg.Vector, nOut)
	tmpG := make([]linalg.Vector, len(vars))
	tmpRG := make([]linalg.Vector, len(vars))
	if len(vars) != 10 {
		panic("b200.LSTM: expected the ten variables of bench/lstm.go:75-99")
	}
	pg, prg := newCPtrArray(10), newCPtrArray(10) // the library reads ten entries; NULL entries are skipped
	defer pg.free()
	defer prg.free()
	var pin runtime.Pinner
	defer pin.Unpin()
	for i := range vars {
		tmpG[i] = make(linalg.Vector, l.ParamLen(i))
		tmpRG[i] = make(linalg.Vector, l.ParamLen(i))
		pin.Pin(&tmpG[i][0])
		pin.Pin(&tmpRG[i][0])
		pg.set(i, hostPtr(tmpG[i]))
		prg.set(i, hostPtr(tmpRG[i]))
	}
	check(C.af_lstm_run_host(l.h, 1, hostPtr(inputs), hostPtr(upstreams), hostPtr(upstreamsR),
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

// AddToVars is Gradient.AddToVars (gradient.go:40-52) on the plan's own parameters, in HBM.
func (l *LSTM) AddToVars(scale float64) { check(C.af_lstm_add_to_vars(l.h, C.double(scale))) }

// ZeroGrads is Gradient.Zero / RGradient.Zero (gradient.go:22-24,69-71) on the resident accumulators.
func (l *LSTM) ZeroGrads() { check(C.af_lstm_zero_grads(l.h)) }
