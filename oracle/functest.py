"""Restatement of the reference's acceptance procedure, functest.RFuncChecker.FullCheck
(functest/rfunc_checker.go:42-68, checker.go:33-113, util.go:13-38).  TEST INFRASTRUCTURE.

The checker is written against a *namespace* `ns` exposing the reference-shaped API
(Variable, RVariable, new_gradient, new_rgradient, add, add_r, mul, mul_r, ComposedRFunc),
so the same procedure runs against the oracle (ns = oracle.autofunc_ref) and against
the product (ns = autofunc_b200.api).  Failures are returned as a list of strings.
"""
from __future__ import annotations

import math

import numpy as np

DEFAULT_DELTA = 1e-5  # functest/func_checker.go:13-21
DEFAULT_PREC = 1e-5


class _AddTwice:  # functest/util.go:13-21
    def __init__(self, ns):
        self.ns = ns

    def apply(self, r):
        return self.ns.add(r, r)

    def apply_r(self, rv, r):
        return self.ns.add_r(r, r)


class _MulTwice:  # functest/util.go:30-38
    def __init__(self, ns):
        self.ns = ns

    def apply(self, r):
        return self.ns.mul(r, r)

    def apply_r(self, rv, r):
        return self.ns.mul_r(r, r)


def _np(x):
    return np.array(x, dtype=np.float64, copy=True)


class RFuncChecker:
    def __init__(self, ns, f, variables, inp, rv, delta=DEFAULT_DELTA, prec=DEFAULT_PREC):
        self.ns, self.f, self.vars, self.input, self.rv = ns, f, list(variables), inp, rv
        self.delta, self.prec = delta, prec
        self.errors: list[str] = []

    # -- helpers ------------------------------------------------------------------
    def _rinput(self):
        return self.ns.RVariable(self.input, self.rv)

    def _bad(self, a, e):
        return math.isnan(a) != math.isnan(e) or abs(a - e) > self.prec

    def _set(self, v, arr):
        """Write a variable's value (product Variables may cache a device copy)."""
        if hasattr(v, "set"):
            v.set(arr)
        else:
            v.vector[:] = arr

    def _get(self, v):
        return _np(v.vector)

    # -- exact quantities (rfunc_checker.go:98-113,150-166) ---------------------------
    def _jacobians(self):
        out = self.f.apply_r(self.rv, self._rinput())
        m = len(out.output())
        jac, jac_r = [], []
        for i in range(m):
            grad, rgrad = self.ns.new_gradient(self.vars), self.ns.new_rgradient(self.vars)
            up = np.zeros(m)
            up[i] = 1
            out.propagate_rgradient(up, np.zeros(m), rgrad, grad)
            jac.append({v: _np(grad[v]) for v in self.vars})
            jac_r.append({v: _np(rgrad[v]) for v in self.vars})
        return jac, jac_r

    def _jacobian_nonr(self):  # functest/func_checker.go Jacobian
        out = self.f.apply(self.input)
        m = len(out.output())
        jac = []
        for i in range(m):
            grad = self.ns.new_gradient(self.vars)
            up = np.zeros(m)
            up[i] = 1
            out.propagate_gradient(up, grad)
            jac.append({v: _np(grad[v]) for v in self.vars})
        return jac

    # -- finite differences (rfunc_checker.go:85-95,117-147,170-190) -------------------
    def _approx_partials(self, v, idx):
        old = self._get(v)
        p = old.copy()
        p[idx] = old[idx] + self.delta
        self._set(v, p)
        v1 = _np(self.f.apply_r(self.rv, self._rinput()).output())
        p[idx] = old[idx] - self.delta
        self._set(v, p)
        v2 = _np(self.f.apply_r(self.rv, self._rinput()).output())
        self._set(v, old)
        return (v1 - v2) * (0.5 / self.delta)

    def _shift_along_rv(self, fn):
        backups = {v: self._get(v) for v in self.rv}
        for v, r in self.rv.items():
            self._set(v, backups[v] - self.delta * _np(r))
        a = fn()
        for v, r in self.rv.items():
            self._set(v, backups[v] + self.delta * _np(r))
        b = fn()
        for v, bk in backups.items():
            self._set(v, bk)
        return (b - a) * (0.5 / self.delta)

    def _approx_partials_r(self, v, idx):
        def exact():
            jac, _ = self._jacobians()
            return np.array([g[v][idx] for g in jac])
        return self._shift_along_rv(exact)

    def _approx_output_r(self):
        return self._shift_along_rv(lambda: _np(self.f.apply_r(self.rv, self._rinput()).output()))

    # -- checks (checker.go:33-113) ---------------------------------------------------
    def check_r(self, tag):
        jac, jac_r = self._jacobians()
        for vi, v in enumerate(self.vars):
            for ei in range(len(v.vector)):
                approx = self._approx_partials(v, ei)
                for oi, g in enumerate(jac):
                    if self._bad(g[v][ei], approx[oi]):
                        self.errors.append(f"{tag} gradient: var {vi} output {oi} entry {ei}: expected {approx[oi]} got {g[v][ei]}")
        exp = self._approx_output_r()
        act = _np(self.f.apply_r(self.rv, self._rinput()).routput())
        for i, a in enumerate(act):
            if self._bad(a, exp[i]):
                self.errors.append(f"{tag} r-output {i}: expected {exp[i]} got {a}")
        for vi, v in enumerate(self.vars):
            for ei in range(len(v.vector)):
                approx = self._approx_partials_r(v, ei)
                for oi, g in enumerate(jac_r):
                    if self._bad(g[v][ei], approx[oi]):
                        self.errors.append(f"{tag} r-gradient: var {vi} output {oi} entry {ei}: expected {approx[oi]} got {g[v][ei]}")

    def check_consistency(self, tag):  # rfunc_checker.go:203-256
        a = _np(self.f.apply(self.input).output())
        b = _np(self.f.apply_r(self.rv, self._rinput()).output())
        if len(a) != len(b) or any(self._bad(x, y) for x, y in zip(a, b)):
            self.errors.append(f"{tag} output inconsistency: Apply {a} ApplyR {b}")
        jn = self._jacobian_nonr()
        jr, _ = self._jacobians()
        if len(jn) != len(jr):
            self.errors.append(f"{tag} jacobian counts differ")
            return
        for i, (x, y) in enumerate(zip(jr, jn)):
            for vi, v in enumerate(self.vars):
                if len(x[v]) != len(y[v]) or any(self._bad(p, q) for p, q in zip(x[v], y[v])):
                    self.errors.append(f"{tag} gradient {i} var {vi}: Apply {y[v]} ApplyR {x[v]}")

    def full_check(self):  # rfunc_checker.go:42-68
        def sub(f, variables, tag):
            c = RFuncChecker(self.ns, f, variables, self.input, self.rv, self.delta, self.prec)
            c.check_r(tag)
            c.check_consistency(tag)
            self.errors += c.errors
        sub(self.f, self.vars, "Standard")
        for i, v in enumerate(self.vars):
            sub(self.f, [v], f"Vars[{i}]")
        sub(self.ns.ComposedRFunc([self.f, _AddTwice(self.ns)]), self.vars, "Accumulation")
        sub(self.ns.ComposedRFunc([self.f, _MulTwice(self.ns)]), self.vars, "Square")
        return self.errors
