// Data parallelism for the cgo shim: the multi-device form of Gradient.Add (gradient.go:28-30,108-120).
// The batch is sharded over the devices (rows of the packed input are independent: batcher.go:6-17), parameters and
// R-vectors are replicated, and the parameter gradients -- sums over samples in the reference -- are summed across
// devices inside the library (include/autofunc_b200.h "data parallelism": NCCL behind the C ABI, csrc/dp.cu).
//
// Two launches are supported, as in the header:
//   - one process per device (how bench.py runs): DPUniqueID on rank 0, ship the 128 bytes, DPInit on every rank;
//   - ONE Go process driving several devices: NewGroup(devices) -- one context per device, one goroutine (locked to
//     its OS thread) per context while a step is in flight.
//
// Like the rest of the package this file cannot be compiled in the build image (no Go toolchain); it is linted against
// the header by tests/test_abi_cpu.py and mirrors autofunc_b200/parallel.py + api.py (Context.dp_init, MLPPlan.step_dp),
// which tests/test_gpu_dp.py and tests/dp_worker.py drive on real GPUs.
package b200

/*
#include <stdlib.h>
#include "autofunc_b200.h"
*/
import "C"

import (
	"runtime"
	"sync"
	"unsafe"

	"github.com/unixpickle/autofunc"
	"github.com/unixpickle/num-analysis/linalg"
)

// DP scheduling modes of a plan (AF_DP_* in the header).
const (
	DPOff      = 0
	DPBlocking = 1 // one all-reduce after the backward pass
	DPOverlap  = 2 // per-layer all-reduces beside the remaining weight-gradient contractions, joined at the end of the step
	DPDeferred = 3 // as DPOverlap, joined when the accumulators are next used
)

// DPUniqueID is rank 0's half of the multi-process rendezvous (af_dp_unique_id).
func DPUniqueID() [C.AF_DP_ID_BYTES]byte {
	var id [C.AF_DP_ID_BYTES]byte
	check(C.af_dp_unique_id(unsafe.Pointer(&id[0])))
	return id
}

// DPInit joins the package context (this process's one device) to the communicator `id` names.
func DPInit(nranks, rank int, id [C.AF_DP_ID_BYTES]byte) {
	check(C.af_dp_init(context(), C.int(nranks), C.int(rank), unsafe.Pointer(&id[0])))
}

// AllReduceSum sums a device vector across the ranks in place (gradient.go:28-30 across devices), asynchronously on
// the context's stream.
func AllReduceSum(v *devVec) { check(C.af_allreduce_sum(context(), v.p, C.int64_t(v.n))) }

// SetDP chooses how the plan's host-buffer steps (StepR) schedule the sum over ranks; StepFreshDP always reduces.
func (m *MLP) SetDP(mode int) { check(C.af_mlp_set_dp(m.h, C.int(mode))) }

// StepFreshDP is one resident bench iteration (bench/mlp.go:43-57, fresh Gradient maps) followed by the sum of every
// rank's {grad | rgrad} block.
func (m *MLP) StepFreshDP(withR bool) {
	r := C.int(0)
	if withR {
		r = 1
	}
	check(C.af_mlp_step_fresh_dp(m.h, r))
}

// StepDP is af_mlp_step from accumulators the caller zeroed (ZeroGrads), then the sum over ranks.
func (m *MLP) StepDP(withR bool) {
	r := C.int(0)
	if withR {
		r = 1
	}
	check(C.af_mlp_step_dp(m.h, r))
}

// DPJoin makes the plan's stream wait for its outstanding reduces (DPDeferred leaves them running past the step).
func (m *MLP) DPJoin() { check(C.af_mlp_dp_join(m.h)) }

// SetHostShard: with several ranks, copy back only this rank's 1/nranks of the reduced gradient block per step.
func (m *MLP) SetHostShard(on bool) {
	v := C.int(0)
	if on {
		v = 1
	}
	check(C.af_mlp_set_host_shard(m.h, v))
}

// StepDP on the LSTM plan: fresh accumulators, forward, BPTT, one all-reduce of {grad | rgrad}.
func (l *LSTM) StepDP(withR bool) {
	r := C.int(0)
	if withR {
		r = 1
	}
	check(C.af_lstm_step_dp(l.h, r))
}

// Group is one Go process driving several devices: one context per device, sharing a communicator.
type Group struct {
	ctxs  []*C.af_ctx
	plans []*C.af_mlp
	sizes []int
	rows  []int // rows of the global batch owned by each rank
}

// NewGroup creates one context per device and joins them (af_dp_init_local).
func NewGroup(devices []int) *Group {
	g := &Group{ctxs: make([]*C.af_ctx, len(devices))}
	for i, d := range devices {
		check(C.af_ctx_create(C.int(d), &g.ctxs[i]))
	}
	arr := (**C.af_ctx)(C.calloc(C.size_t(len(devices)), C.size_t(unsafe.Sizeof((*C.af_ctx)(nil)))))
	defer C.free(unsafe.Pointer(arr))
	copy(unsafe.Slice(arr, len(devices)), g.ctxs)
	check(C.af_dp_init_local(arr, C.int(len(devices))))
	return g
}

// ShardRows splits n rows over the group the way parallel.shard_rows does: blocks differ by at most one row.
func (g *Group) ShardRows(n int) []int {
	w := len(g.ctxs)
	rows := make([]int, w)
	for r := range rows {
		rows[r] = n / w
		if r < n%w {
			rows[r]++
		}
	}
	return rows
}

// NewMLP builds the bench network on every device for a GLOBAL batch of n rows and uploads the replicated parameters.
func (g *Group) NewMLP(sizes []int, n int, vars []*autofunc.Variable, rv autofunc.RVector) {
	g.sizes, g.rows = sizes, g.ShardRows(n)
	cs := make([]C.int64_t, len(sizes))
	for i, s := range sizes {
		cs[i] = C.int64_t(s)
	}
	g.plans = make([]*C.af_mlp, len(g.ctxs))
	for r, c := range g.ctxs {
		check(C.af_mlp_create(c, &cs[0], C.int(len(sizes)-1), C.int64_t(g.rows[r]), &g.plans[r]))
		check(C.af_mlp_set_dp(g.plans[r], C.int(DPOverlap)))
		for i, v := range vars {
			check(C.af_mlp_set_param(g.plans[r], C.int(i), hostPtr(v.Vector), C.int64_t(len(v.Vector))))
			if rvec, ok := rv[v]; ok {
				check(C.af_mlp_set_rvector(g.plans[r], C.int(i), hostPtr(rvec), C.int64_t(len(rvec))))
			} else {
				check(C.af_mlp_set_rvector(g.plans[r], C.int(i), nil, 0))
			}
		}
	}
}

// StepR runs one data-parallel iteration of bench/mlp.go:52-54 on a global batch: every device takes its rows of
// `input`, the gradients are summed inside the library, and rank 0's copy is ADDED into grad / rgrad (variable.go:45).
// One goroutine per device, each locked to its OS thread for the duration of its calls.
func (g *Group) StepR(vars []*autofunc.Variable, input linalg.Vector, rgrad autofunc.RGradient,
	grad autofunc.Gradient) (out, outR linalg.Vector) {
	in, top := g.sizes[0], g.sizes[len(g.sizes)-1]
	total := 0
	for _, c := range g.rows {
		total += c
	}
	out, outR = make(linalg.Vector, total*top), make(linalg.Vector, total*top)
	nParams := len(vars)
	tmpG, tmpRG := make([]linalg.Vector, nParams), make([]linalg.Vector, nParams)
	for i, v := range vars {
		tmpG[i], tmpRG[i] = make(linalg.Vector, len(v.Vector)), make(linalg.Vector, len(v.Vector))
	}
	var wg sync.WaitGroup
	start := 0
	for r := range g.ctxs {
		wg.Add(1)
		go func(r, start int) {
			defer wg.Done()
			runtime.LockOSThread()
			defer runtime.UnlockOSThread()
			rows := g.rows[r]
			pg, prg := newCPtrArray(nParams), newCPtrArray(nParams)
			defer pg.free()
			defer prg.free()
			var pin runtime.Pinner
			defer pin.Unpin()
			if r == 0 { // every rank holds the same sums; one download is enough
				for i := range vars {
					if len(tmpG[i]) > 0 {
						pin.Pin(&tmpG[i][0])
						pin.Pin(&tmpRG[i][0])
					}
					pg.set(i, hostPtr(tmpG[i]))
					prg.set(i, hostPtr(tmpRG[i]))
				}
			}
			check(C.af_mlp_step_host(g.plans[r], 1, hostPtr(input[start*in:(start+rows)*in]), nil, nil,
				hostPtr(out[start*top:(start+rows)*top]), hostPtr(outR[start*top:(start+rows)*top]), pg.base, prg.base))
		}(r, start)
		start += g.rows[r]
	}
	wg.Wait()
	for i, v := range vars {
		if d, ok := grad[v]; ok {
			d.Add(tmpG[i])
		}
		if d, ok := rgrad[v]; ok {
			d.Add(tmpRG[i])
		}
	}
	return
}

// Close releases the plans, the communicator and the contexts.
func (g *Group) Close() {
	for _, p := range g.plans {
		C.af_mlp_destroy(p)
	}
	for _, c := range g.ctxs {
		C.af_dp_destroy(c)
		C.af_ctx_destroy(c)
	}
	g.plans, g.ctxs = nil, nil
}
