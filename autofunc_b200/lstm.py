"""Host-side wrapper of the LSTM fast path (af_lstm_*): bench/lstm.go's lstmNet batched over B lanes,
forward + BPTT (+ R-operator), everything resident in HBM."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import Context, DeviceVec, check, context

F64 = np.float64
PARAM_NAMES = ["InWeights", "InBiases", "InGateWeights", "InGateBiases", "ForgetGateWeights", "ForgetGateBiases",
               "OutputGate", "OutputGateBiases", "OutputWeights", "OutputBiases"]  # bench/lstm.go:75-99


class LSTMPlan:
    def __init__(self, input_size, hidden, output_size, time_steps, lanes, ctx: Context | None = None):
        self.ctx = ctx or context()
        self.I, self.H, self.O, self.T, self.B = map(int, (input_size, hidden, output_size, time_steps, lanes))
        h = C.c_void_p()
        check(self.ctx.lib.af_lstm_create(self.ctx.h, self.I, self.H, self.O, self.T, self.B, C.byref(h)))
        self.h = h

    def param_len(self, idx):
        rows = self.H if idx // 2 < 4 else self.O
        return rows if idx & 1 else rows * (self.H + self.I)

    def set_param(self, idx, value):
        a = np.ascontiguousarray(value, dtype=F64).reshape(-1)
        check(self.ctx.lib.af_lstm_set_param(self.h, idx, a.ctypes.data, a.size))

    def set_rvector(self, idx, value):
        if value is None:
            check(self.ctx.lib.af_lstm_set_rvector(self.h, idx, None, 0))
            return
        a = np.ascontiguousarray(value, dtype=F64).reshape(-1)
        check(self.ctx.lib.af_lstm_set_rvector(self.h, idx, a.ctypes.data, a.size))

    def _vec(self, ptr, n):
        return DeviceVec(self.ctx, n, ptr, owner=self)

    def grad(self, idx):
        return self._vec(self.ctx.lib.af_lstm_grad_ptr(self.h, idx), self.param_len(idx))

    def rgrad(self, idx):
        return self._vec(self.ctx.lib.af_lstm_rgrad_ptr(self.h, idx), self.param_len(idx))

    def output(self):
        return self._vec(self.ctx.lib.af_lstm_output_ptr(self.h), self.T * self.B * self.O)

    def routput(self):
        return self._vec(self.ctx.lib.af_lstm_routput_ptr(self.h), self.T * self.B * self.O)

    def upstream(self):
        return self._vec(self.ctx.lib.af_lstm_upstream_ptr(self.h), self.T * self.B * self.O)

    def upstream_r(self):
        return self._vec(self.ctx.lib.af_lstm_upstream_r_ptr(self.h), self.T * self.B * self.O)

    def grad_slab(self):
        p = C.c_void_p()
        n = self.ctx.lib.af_lstm_grad_slab(self.h, C.byref(p))
        return self._vec(p.value, int(n))

    def zero_grads(self):
        check(self.ctx.lib.af_lstm_zero_grads(self.h))

    def add_to_vars(self, scale: float):
        """Gradient.AddToVars (gradient.go:40-52) on the device: param += scale * grad."""
        check(self.ctx.lib.af_lstm_add_to_vars(self.h, float(scale)))

    def param(self, idx):
        return self._vec(self.ctx.lib.af_lstm_param_ptr(self.h, idx), self.param_len(idx))

    def set_inputs(self, d_x: DeviceVec):
        check(self.ctx.lib.af_lstm_set_inputs(self.h, d_x.ptr))

    def forward(self, with_r=True):
        check(self.ctx.lib.af_lstm_forward(self.h, int(with_r)))

    def backward(self, with_r=True):
        check(self.ctx.lib.af_lstm_backward(self.h, int(with_r)))

    def step_dp(self, with_r=True):
        """Fresh accumulators, forward, BPTT, then the sum of the ranks' {grad | rgrad} blocks (af_lstm_step_dp)."""
        check(self.ctx.lib.af_lstm_step_dp(self.h, int(with_r)))

    def run_host(self, x, upstreams, upstreams_r=None, with_r=True, out=None, rout=None, grads=None, rgrads=None):
        """x: T x B x I, upstreams: T x B x O (host arrays).  Returns (out, rout, grads, rgrads)."""
        n_out = self.T * self.B * self.O
        out = np.empty(n_out, dtype=F64) if out is None else out
        rout = np.empty(n_out, dtype=F64) if (rout is None and with_r) else rout
        grads = [np.empty(self.param_len(i), dtype=F64) for i in range(10)] if grads is None else grads
        rgrads = [np.empty(self.param_len(i), dtype=F64) for i in range(10)] if (rgrads is None and with_r) else rgrads
        HP = C.c_void_p * 10
        ptr = lambda a: None if a is None else a.ctypes.data
        x = np.ascontiguousarray(x, dtype=F64)
        upstreams = np.ascontiguousarray(upstreams, dtype=F64)
        upstreams_r = None if upstreams_r is None else np.ascontiguousarray(upstreams_r, dtype=F64)
        g_arr = HP(*[g.ctypes.data for g in grads])
        rg_arr = HP(*[g.ctypes.data for g in rgrads]) if rgrads is not None else None
        check(self.ctx.lib.af_lstm_run_host(self.h, int(with_r), ptr(x), ptr(upstreams), ptr(upstreams_r), ptr(out), ptr(rout),
                                            g_arr, rg_arr))
        return out, rout, grads, rgrads

    def close(self):
        if self.h is not None:
            self.ctx.lib.af_lstm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
