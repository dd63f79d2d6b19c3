#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: bench/mlp.go's MLP (1000-2000-512-10,
sigmoid) doing forward + Gradient + RGradient (RunR with backProp, bench/mlp.go:43-57), batched,
float64, samples/sec.

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference

One "step" = one pass of the hot path over one batch of synthetic inputs: fresh accumulators, fresh
outGrad = 1 / outRGrad = 0 (bench/mlp.go:118-126), ApplyR, PropagateRGradient with non-nil grad; for
N > 1 each rank owns `batch` samples (weak scaling) and the step includes the sum of {grad, rgrad}
over the ranks (Gradient.Add, gradient.go:28-30) -- issued INSIDE the library (af_mlp_step_fresh_dp:
NCCL behind the C ABI, csrc/dp.cu); torch.distributed only launches the ranks, ships the 128-byte
communicator id and takes the max of the ranks' timings.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

LAYER_SIZES = [1000, 2000, 512, 10]  # bench/mlp.go:17-20
C5_SIZES = [8192] * 5                # BASELINE configs[4]: 4 x (8192 -> 8192)
METRIC = "mlp_fwd_gradient_rgradient_samples_per_sec"
UNIT = "samples/s"
DP_MODES = {"blocking": 1, "overlap": 2, "deferred": 3}


def flops_per_sample(sizes, with_r=True):
    """per dense layer 18 K M (12 K M for layer 1, whose input is constant) with R,
    6 K M (4 K M) without."""
    f = 0
    for l in range(len(sizes) - 1):
        km = sizes[l] * sizes[l + 1]
        if with_r:
            f += (12 if l == 0 else 18) * km
        else:
            f += (4 if l == 0 else 6) * km
    return f


def workload_name(batch):
    return (f"bench/mlp.go MLP {'-'.join(map(str, LAYER_SIZES))} sigmoid, batched n={batch}/GPU, "
            "fwd+Gradient+RGradient (RunR backProp=true), float64")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy + OpenBLAS threads) -- the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _blas_limit(threads):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=threads, user_api="blas")
    except Exception:
        return None


def cpu_thread_sweep(n):
    """The stated CPU baseline is the oracle's BEST thread count, not its default: one step of the workload at every
    power-of-two fraction of the host's cores (OpenBLAS often peaks below the hyper-thread count)."""
    from oracle import autofunc_ref as ref
    net = ref.MLP(LAYER_SIZES)
    x = net.make_input(n)
    cores = host_cores()
    cands = sorted({max(1, cores >> s) for s in range(0, 4)}, reverse=True)
    sweep = {}
    for t in cands:
        lim = _blas_limit(t)
        try:
            net.step_r(x, n)  # warm (thread pool resize, page faults)
            t0 = time.perf_counter()
            net.step_r(x, n)
            sweep[t] = n / (time.perf_counter() - t0)
        finally:
            if lim is not None:
                lim.restore_original_limits()
    best = max(sweep, key=sweep.get)
    return best, {str(k): v for k, v in sweep.items()}


def run_cpu_oracle(n, steps, warmup, threads):
    """The CPU arm uses the host threads it is told to: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would
    throttle OpenBLAS to one thread, so the BLAS pool is sized explicitly."""
    from oracle import autofunc_ref as ref
    limiter = _blas_limit(threads)
    try:
        net = ref.MLP(LAYER_SIZES)
        x = net.make_input(n)
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            net.step_r(x, n)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
        return times
    finally:
        if limiter is not None:
            limiter.restore_original_limits()


def run_cpu_native(n, steps, warmup):
    """SURVEY 8(d)(i) "oracle-native": the same step with the reference's Gemm-per-call-site structure and a blocked OpenMP
    Gemm (oracle/native_ref.cpp) -- the stand-in for the reference's DEFAULT pure-Go BLAS backend (`go test -bench .
    ./bench`); the numpy/OpenBLAS oracle above stands in for ./bench_cblas and is the stronger baseline."""
    from oracle import autofunc_ref as ref, native
    native.set_threads(host_cores())
    net = ref.MLP(LAYER_SIZES)
    params = [v.vector for v in net.variables]
    rvecs = [net.rvec[v] for v in net.variables]
    x = net.make_input(n).vector
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        native.mlp_step_r(LAYER_SIZES, n, params, rvecs, x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return {"value": n * len(times) / sum(times), "unit": UNIT, "cores": native.threads(), "kind": "port",
            "sample": f"{len(times)} steps of {n} samples; oracle/native_ref.cpp: one blocked OpenMP Gemm per lin_tran.go call site "
                      "(9 per layer), separate bias / sigmoid passes, SSE2 codegen -- stand-in for the reference's default gonum "
                      "native backend (Go unavailable)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.cpu_batch or args.batch  # the SAME configuration as the GPU arm: n = 4096 per step
    threads, sweep = cpu_thread_sweep(n)
    times = run_cpu_oracle(n, args.steps, args.warmup, threads)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "layer_sizes": LAYER_SIZES, "batch_per_gpu": args.batch,
                   "global_batch": args.batch * args.gpus},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(times)} steps of {n} samples each (the GPU arm's own batch); "
                                   "CPU restatement of the reference (oracle/autofunc_ref.py, numpy + OpenBLAS): "
                                   "the reference is pure Go and no Go toolchain exists on this box",
                         "host_cores": host_cores(), "thread_sweep_samples_per_s": sweep},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    try:
        line["other_cpu_baselines"] = {"oracle_native": run_cpu_native(min(n, 256), 2, 1)}
    except Exception as e:  # the headline CPU number above must not depend on the second baseline
        line["other_cpu_baselines"] = {"oracle_native": {"unavailable": str(e)[:200]}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region, the way /opt/skills/guides/B200_PROFILING.md does it: an
    `nvidia-smi --query-gpu ... -lms` child process started before the region and stopped after it.  (Round 1 polled
    NVML from a thread of the benchmark process itself; with NCCL in the same process every poll stalled that rank's
    kernel launches -- measured at 2 GPUs: 5.6-5.9 ms per step with the in-process sampler, 4.68 ms without.)"""
    FIELDS = ["timestamp", "clocks.sm", "clocks.max.sm", "clocks_event_reasons.hw_slowdown", "clocks_event_reasons.hw_thermal_slowdown",
              "clocks_event_reasons.sw_thermal_slowdown", "clocks_event_reasons.sw_power_cap"]

    def __init__(self, gpu_id, period_ms=25):
        self.gpu_id, self.period_ms, self.proc = str(gpu_id), period_ms, None

    def start(self):
        import subprocess
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + ",".join(self.FIELDS), "--format=csv,noheader,nounits",
                                          "-i", self.gpu_id, "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            import atexit
            atexit.register(lambda p=self.proc: p.poll() is None and p.kill())  # never leave the looping child behind
        except Exception:
            self.proc = None

    def mark(self):
        """called at the start of the timed region: samples taken before this instant are dropped (the child is started
        early -- nvidia-smi needs a good part of a second to come up -- and samples every period from then on)"""
        self.t0 = time.time()

    def stop(self):
        import datetime
        t1 = time.time()
        rows, max_mhz = [], None
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
            for line in out.splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) != len(self.FIELDS):
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    mhz, max_mhz = int(float(f[1])), int(float(f[2]))
                except ValueError:
                    continue
                active = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:])
                          if v == "Active"]
                rows.append((ts, mhz, active))
        t0 = getattr(self, "t0", 0.0)
        inside = [r for r in rows if t0 <= r[0] <= t1 + 0.001 * self.period_ms]
        if not inside and rows:  # a region shorter than one period: the sample nearest to it
            inside = [min(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
        reasons = sorted({a for r in inside for a in r[2]})
        return {"sm_mhz": statistics.median([r[1] for r in inside]) if inside else None, "sm_max_mhz": max_mhz,
                "reasons": reasons, "samples": len(inside),
                "how": f"nvidia-smi --query-gpu -lms {self.period_ms} child process; samples inside the timed region"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measure_dgemm_peak(torch, size=8192, iters=6):
    """cuBLAS Dgemm 8192^3 -- the FP64 denominator (MEASURED_PEAKS.json holds no FP64 figure)."""
    a = torch.rand(size, size, dtype=torch.float64, device="cuda")
    b = torch.rand(size, size, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * size ** 3 / (best * 1e-3) / 1e12


def measure_dgemm_same_shapes(torch, n, sizes, iters=5):
    """Context for `roofline.frac`: cuBLAS Dgemm (torch.matmul f64) on the bench network's OWN contraction shapes -- the
    nine products of each layer's forward / input-gradient / weight-gradient launches, run one product at a time as the
    reference's bench_cblas path would (lin_tran.go:99-355).  Flop-weighted rate over the layers wider than 16 outputs."""
    tot_flops, tot_ms = 0.0, 0.0
    for l in range(len(sizes) - 1):
        K, M = sizes[l], sizes[l + 1]
        if M <= 16:
            continue
        x = torch.rand(n, K, dtype=torch.float64, device="cuda")
        w = torch.rand(M, K, dtype=torch.float64, device="cuda")
        u = torch.rand(n, M, dtype=torch.float64, device="cuda")
        shapes = [(lambda: torch.matmul(x, w.t()), 3 if l else 2),   # forward: X W^T, X RW^T (+ RX W^T)
                  (lambda: torch.matmul(u.t(), x), 3 if l else 2)]   # weight gradient: U^T X, UR^T X (+ U^T RX)
        if l:
            shapes.append((lambda: torch.matmul(u, w), 3))           # input gradient: U W, UR W, U RW
        for fn, products in shapes:
            fn(); fn()
            best = 1e30
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            tot_flops += products * 2.0 * n * K * M
            tot_ms += products * best
        del x, w, u
    torch.cuda.empty_cache()
    return tot_flops / (tot_ms * 1e-3) / 1e12 if tot_ms else None


def gpu_arm(args):
    import torch
    import torch.distributed as dist

    from autofunc_b200 import _lib, api, build as af_build, parallel, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- autofunc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.batch
    K, W = args.steps, args.warmup

    gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    sampler = ClockSampler(gpu_uuid if gpu_uuid.startswith("GPU-") else "GPU-" + gpu_uuid) if (rank == 0 and not args.no_clocks) else None
    if sampler:
        sampler.start()  # early: the child needs a while to come up; only samples inside the timed region are kept
    stream = torch.cuda.Stream()  # a real stream: kernels, copies and the timing events are ordered on it
    torch.cuda.set_stream(stream)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)
    api.set_context(ctx)
    ctx.set_gemm_path(2)  # TMA + DMMA or an error; never a silent fallback
    for name in ("big_cfg", "persist_tiles", "sk_bias", "sk_rotate", "dp_max_ctas", "dp_reserve_sms", "deterministic"):
        v = getattr(args, name)
        if v is not None:
            ctx.set_option(name, v)
    for kv in args.opt:  # A/B switches of the library by name (af_set_option), e.g. --opt few_rows_wgrad=2
        name, _, v = kv.partition("=")
        ctx.set_option(name, int(v))
    dp_info = None
    if world > 1:
        # the communicator lives inside libautofunc_b200 (af_dp_init); torch only carries its 128-byte id
        parallel.dp_init_torch(ctx, local_rank)
        nr, rk, ver = ctx.dp_info()
        assert (nr, rk) == (world, rank)
        dp_info = {"transport": "NCCL behind the C ABI (af_dp_init / af_mlp_step_fresh_dp, csrc/dp.cu)",
                   "nccl_version": ver, "mode": args.dp_mode, "max_ctas": args.dp_max_ctas if args.dp_max_ctas is not None else 8}
    plan = api.MLPPlan(LAYER_SIZES, n, ctx)
    params, rvecs = synth.mlp_params(LAYER_SIZES)
    for i, (p, r) in enumerate(zip(params, rvecs)):
        plan.set_param(i, p)
        plan.set_rvector(i, r)
    x_host = synth.mlp_input(LAYER_SIZES, n, seed=124 + 7919 * rank)
    ctx.lib.af_upload(ctx.h, plan.input().ptr, x_host.ctypes.data, x_host.size)
    n_out = n * LAYER_SIZES[-1]
    P = 2 * plan.L
    n_params = sum(plan.param_len(i) for i in range(P))
    plan.set_dp(DP_MODES[args.dp_mode] if world > 1 else 0)
    slab = plan.grad_slab()
    slab_t = torch.as_tensor(parallel.CudaArrayView(slab.ptr, len(slab)), device=f"cuda:{local_rank}")

    up, upr = plan.upstream(), plan.upstream_r()

    def legacy_step():  # round 1's call sequence for the same iteration (A/B of the fresh-step entry point; N = 1 only)
        plan.zero_grads()
        api.check(ctx.lib.af_fill(ctx.h, up.ptr, 1.0, n_out))
        api.check(ctx.lib.af_zero(ctx.h, upr.ptr, n_out))
        plan.step(True, True)

    def step():
        if args.legacy_step and world == 1:
            return legacy_step()
        # one bench iteration as ONE library call: fresh accumulators, outGrad = 1 / outRGrad = 0, ApplyR,
        # PropagateRGradient, and (N > 1) the sum over ranks scheduled per the plan's dp mode
        plan.step_fresh_dp(True)

    def drain():
        plan.dp_join()  # deferred mode: the last step's reduce completes inside the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, and (N > 1) the checksum identity of the data-parallel reduce:
    #      sum(reduced block) on every rank == sum over ranks of sum(local block)
    dp_check = None
    if world > 1:
        plan.step_fresh(True)  # local step, no reduce
        local = torch.stack([slab_t.sum(), slab_t.abs().sum()])
        dist.all_reduce(local)
        step()
        drain()
        got = slab_t.sum()
        rel = float((got - local[0]).abs() / local[1])
        lo, hi = got.clone(), got.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dp_check = {"identity": "sum(all-reduced {grad|rgrad}) == sum over ranks of sum(local {grad|rgrad})",
                    "rel_err_vs_sum_abs": rel, "ranks_identical": bool(float(lo) == float(hi)), "ok": rel < 1e-10}
        if not dp_check["ok"] or not dp_check["ranks_identical"]:
            raise SystemExit(f"bench.py: data-parallel checksum identity violated: {dp_check}")
    for _ in range(W):
        step()
    drain()
    barrier()
    # Per-launch device times for the roofline come from the LAST `traced` steps of the timed region (one CUDA event after
    # every launch costs ~1 % of the step, so the other steps run untraced); the clocks are sampled by rank 0 only --
    # eight processes polling NVML every 20 ms slowed every rank's launches by 0.2 ms per step at N = 8.
    traced = min(K, args.trace_steps)
    cap = 64 * traced + 8
    kinds, work, ms_arr = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_float * cap)()
    count = C.c_int(0)
    # (the trace's CUDA events are created here, outside the timed region: creating ~260 events between two timed steps
    # stalled the enqueueing thread for milliseconds once NCCL's threads were contending for the driver)
    api.check(ctx.lib.af_trace_begin(ctx.h, cap))
    api.check(ctx.lib.af_trace_end(ctx.h, cap, kinds, work, ms_arr, C.byref(count)))
    barrier()
    launches0 = ctx.launch_count()
    dp_calls0 = ctx.dp_calls()
    if sampler:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K):
        if traced and k == K - traced:
            api.check(ctx.lib.af_trace_begin(ctx.h, cap))
        step()
    drain()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    if traced:
        api.check(ctx.lib.af_trace_end(ctx.h, cap, kinds, work, ms_arr, C.byref(count)))
    else:
        count.value = 0
    launches = ctx.launch_count() - launches0
    dp_calls = ctx.dp_calls() - dp_calls0
    if args.trace and rank == 0:
        per_step = count.value // max(1, traced)
        names = {0: "gemm_dual", 1: "gemm_single", 2: "gemm_generic", 3: "elementwise", 4: "reduce", 5: "streaming"}
        for i in range(count.value - per_step, count.value):
            rate = work[i] / (ms_arr[i] * 1e-3) / 1e12 if ms_arr[i] > 0 else 0.0
            unit = "TFLOP/s" if kinds[i] < 3 else "TB/s"
            print(f"trace[{i - (count.value - per_step):2d}] {names.get(kinds[i], str(kinds[i])):12s} {ms_arr[i] * 1e3:9.1f} us  work {work[i]:.4g}  {rate:7.2f} {unit}",
                  file=sys.stderr)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world * n * K / (ms_total * 1e-3)
    if args.profile_step:
        # launch-list runs (ncu --metrics gpu__time_duration.sum): nothing but the W + K steps of the timed workload is
        # launched, so the listed kernels ARE the step and a kernel's share of the list is its share of the step
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_total / K, "steps": K,
                              "warmup": W, "profile_step": True, "gpu_launches": int(launches)}), flush=True)
        plan.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    def timed(fn, reps, after=None):
        """device time of `reps` calls of fn on the ctx stream, max over ranks, ms per call"""
        for _ in range(3):
            fn()
        if after:
            after()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            fn()
        if after:
            after()
        a1.record(stream)
        barrier()
        return max_over_ranks(a0.elapsed_time(a1)) / reps

    # ---- the reference benchmark's other two modes (bench/mlp.go:29-41): Run(b, false) and Run(b, true) ----
    reps = max(3, K // 2)
    other_modes = {"fwd_only_samples_per_s": world * n / (timed(lambda: plan.step(False, False), reps) * 1e-3),
                   "fwd_gradient_samples_per_s": world * n / (timed(lambda: plan.step_fresh_dp(False), reps, drain) * 1e-3)}

    # ---- (N > 1) what the reduce costs: the same step without it, the collective alone, and the other schedules ----
    dp_report = None
    if world > 1:
        local_ms = timed(lambda: plan.step_fresh(True), K)
        ar_ms = timed(lambda: ctx.allreduce_sum(slab), K)
        modes = {}
        for name, code in DP_MODES.items():
            plan.set_dp(code)
            modes[name] = world * n / (timed(step, K, drain) * 1e-3)
        plan.set_dp(DP_MODES[args.dp_mode])
        dp_report = dict(dp_info, allreduce_bytes=len(slab) * 8, allreduce_alone_ms=ar_ms,
                         step_without_reduce_ms=local_ms, exposed_ms_per_step=ms_total / K - local_ms,
                         collectives_per_step=dp_calls / K, samples_per_s_by_mode=modes, check=dp_check)

    # ---- roofline of the dominant kernel (the dual TMA+DMMA contraction), from the live trace ----
    gemm_ms = sum(ms_arr[i] for i in range(count.value) if kinds[i] == 0)
    gemm_flops = sum(work[i] for i in range(count.value) if kinds[i] == 0)
    gemm_n = sum(1 for i in range(count.value) if kinds[i] == 0)
    other_ms = sum(ms_arr[i] for i in range(count.value) if kinds[i] != 0)
    traced_ms = sum(ms_arr[i] for i in range(count.value))
    peak = measure_dgemm_peak(torch) if rank == 0 else None
    same_shape = measure_dgemm_same_shapes(torch, n, LAYER_SIZES) if rank == 0 else None

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) ----
    def pinned(count_):
        return torch.empty(count_, dtype=torch.float64).pin_memory()
    # two sets of pinned host buffers: two steps are in flight (af_mlp_submit_host)
    sets = []
    for _ in range(2):
        hx = pinned(x_host.size)
        hx.numpy()[:] = x_host
        hg = [pinned(plan.param_len(i)) for i in range(P)]
        hrg = [pinned(plan.param_len(i)) for i in range(P)]
        HP = C.c_void_p * P
        sets.append(dict(hx=hx, hout=pinned(n_out), hrout=pinned(n_out), hg=hg, hrg=hrg,
                         g_arr=HP(*[t.data_ptr() for t in hg]), rg_arr=HP(*[t.data_ptr() for t in hrg])))
    # N > 1: the reduced gradient is the same on every rank, so each rank copies back only its 1/N of the block
    # (af_mlp_set_host_shard): the host receives the full reduced gradient once per step, not N times
    if world > 1:
        plan.set_host_shard(True)
    h2d = x_host.size * 8
    d2h = 2 * n_out * 8 + (len(slab) * 8 // world if world > 1 else 2 * n_params * 8)
    e2e_k = [0]

    def e2e_step():
        b = sets[e2e_k[0] & 1]
        e2e_k[0] += 1
        # the reference-facing host-buffer call, pipelined: upload of step k+1 and download of step k-1 overlap the
        # compute (and, for N > 1, the reduce) of step k; every step moves its own inputs and results
        api.check(ctx.lib.af_mlp_submit_host(plan.h, 1, b["hx"].data_ptr(), None, None, b["hout"].data_ptr(),
                                             b["hrout"].data_ptr(), b["g_arr"], b["rg_arr"]))

    def e2e_drain():
        api.check(ctx.lib.af_mlp_wait(plan.h))
        api.check(ctx.lib.af_mlp_wait(plan.h))
    for _ in range(max(1, min(W, 3))):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    e2e_drain()
    torch.cuda.synchronize()
    e2e_value = world * n * K / max_over_ranks(time.perf_counter() - t0)
    plan.set_host_shard(False)
    api.check(ctx.lib.af_mlp_bind_slot(plan.h, 0))

    # ---- (N > 1) strong scaling: the same global problems split N ways ----
    strong = None
    if world > 1 and not args.no_strong:
        strong = {}
        try:
            strong["mlp_global_4096"] = strong_mlp(torch, dist, ctx, api, parallel, synth, world, rank, local_rank, stream,
                                                   LAYER_SIZES, 4096, K, args.dp_mode, max_over_ranks, barrier)
        except Exception as e:
            strong["mlp_global_4096"] = {"error": repr(e)[:300]}
        plan.close()
        plan = None
        try:
            strong["c5_global_65536"] = strong_mlp(torch, dist, ctx, api, parallel, synth, world, rank, local_rank, stream,
                                                   C5_SIZES, 65536, 2, args.dp_mode, max_over_ranks, barrier, device_init=True)
        except Exception as e:
            strong["c5_global_65536"] = {"error": repr(e)[:300]}

    if rank == 0:
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("kernel_source_sha") == af_build.kernel_source_sha():
                    traffic = tj.get("dram_bytes_per_launch")
                else:
                    traffic_note = ("null: profiles/roofline_traffic.json was captured on kernel sources "
                                    f"{tj.get('kernel_source_sha')}, this build is {af_build.kernel_source_sha()}")
            except Exception:
                traffic = None
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "layer_sizes": LAYER_SIZES, "batch_per_gpu": n,
                       "global_batch": n * world, "parallelism": f"dp{world}",
                       "allreduce": "none" if world == 1 else
                                    {"blocking": "one all-reduce of {grad|rgrad} after the backward pass, on the step's stream",
                                     "overlap": "per-layer all-reduces on the communicator's stream beside the remaining weight-gradient "
                                                "contractions (largest layer first), joined at the end of every step",
                                     "deferred": "as overlap, the tail of the reduce joined before the NEXT step's accumulators are "
                                                 "zeroed (it overlaps that step's forward pass); all reduces complete inside the "
                                                 "timed region"}[args.dp_mode],
                       "l2": "no flush: each step streams ~%.0f MB of activations/upstreams (> 126 MB L2)" %
                             (n * sum(LAYER_SIZES) * 8 * 4 / 1e6),
                       "flops_per_sample": flops_per_sample(LAYER_SIZES),
                       "deterministic": bool(args.deterministic)},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if (achieved and peak) else None, "traffic": traffic,
                         "kernel": "gemm_dmma_kernel<*,*,2> (TMA + DMMA.8x8x4 dual contraction)",
                         "launches_timed": gemm_n, "avg_launch_ms": gemm_ms / gemm_n if gemm_n else None,
                         "algorithmic_flops_per_launch": gemm_flops / gemm_n if gemm_n else None,
                         "kernel_share_of_step": gemm_ms / traced_ms if traced_ms else None,
                         "other_kernels_ms_per_step": other_ms / max(1, traced),
                         "traced_steps": traced,
                         "peak_source": "cuBLAS Dgemm 8192^3 (torch.matmul f64, best of 6) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "cublas_dgemm_on_the_same_shapes_tflops": same_shape,
                         "frac_of_cublas_on_the_same_shapes": (achieved / same_shape) if (achieved and same_shape) else None,
                         "whole_step_tflops": flops_per_sample(LAYER_SIZES) * n * K / (ms_total * 1e-3) / 1e12 * world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "other_modes": other_modes,  # Run(b,false) / Run(b,true) of bench/mlp.go:29-41, resident, same batch
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "how")} if clocks else None,
        }
        if traffic_note:
            line["roofline"]["traffic_note"] = traffic_note
        if world > 1:
            line["e2e"]["note"] = ("per rank: every rank uploads its own batch and downloads its outputs plus its 1/N of the "
                                   "all-reduced {grad|rgrad} block (af_mlp_set_host_shard)")
            line["data_parallel"] = dp_report
            if strong is not None:
                line["strong_scaling"] = strong
        if world == 1 and not args.no_cpu_baseline:
            nb = args.cpu_batch or n
            threads, sweep = cpu_thread_sweep(nb)
            times = run_cpu_oracle(nb, 3, 1, threads)
            line["cpu_baseline"] = {
                "value": nb * len(times) / sum(times), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"3 steps of {nb} samples (same network, same step, same batch); oracle/autofunc_ref.py "
                          "(numpy + OpenBLAS at its best thread count) -- CPU restatement, Go unavailable",
                "host_cores": host_cores(), "thread_sweep_samples_per_s": sweep}
            try:  # the weaker of the two CPU baselines SURVEY 8(d) names, for the record
                line["other_cpu_baselines"] = {"oracle_native": run_cpu_native(min(nb, 256), 2, 1)}
            except Exception as e:
                line["other_cpu_baselines"] = {"oracle_native": {"unavailable": str(e)[:200]}}
        print(json.dumps(line), flush=True)
    if plan is not None:
        plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def strong_mlp(torch, dist, ctx, api, parallel, synth, world, rank, local_rank, stream, sizes, n_global, steps, dp_mode,
               max_over_ranks, barrier, device_init=False):
    """Strong scaling: a FIXED global batch split over the ranks (parallel.shard_rows), the same data-parallel step."""
    start, count = parallel.shard_rows(n_global, world, rank)
    plan = api.MLPPlan(sizes, count, ctx)
    try:
        P = 2 * plan.L
        if device_init:
            # C5's parameters are 4 x 537 MB (+ R-vectors): filled on the device (timing only; values U[-1,1)/sqrt(K))
            for i in range(P):
                for ptr_of in (ctx.lib.af_mlp_param_ptr, ctx.lib.af_mlp_rvector_ptr):
                    t = torch.as_tensor(parallel.CudaArrayView(ptr_of(plan.h, i), plan.param_len(i)), device=f"cuda:{local_rank}")
                    with torch.cuda.stream(stream):
                        t.uniform_(-1.0, 1.0)
                        if i % 2 == 0 and ptr_of is ctx.lib.af_mlp_param_ptr:
                            t.mul_(1.0 / np.sqrt(sizes[i // 2]))
            xt = torch.as_tensor(parallel.CudaArrayView(plan.input().ptr, count * sizes[0]), device=f"cuda:{local_rank}")
            with torch.cuda.stream(stream):
                xt.uniform_(-1.0, 1.0)
        else:
            params, rvecs = synth.mlp_params(sizes)
            for i, (p, r) in enumerate(zip(params, rvecs)):
                plan.set_param(i, p)
                plan.set_rvector(i, r)
            x = synth.mlp_input(sizes, n_global).reshape(n_global, sizes[0])[start:start + count]
            x = np.ascontiguousarray(x).reshape(-1)
            ctx.lib.af_upload(ctx.h, plan.input().ptr, x.ctypes.data, x.size)
        plan.set_dp(DP_MODES[dp_mode])
        for _ in range(1 if device_init else 3):
            plan.step_fresh_dp(True)
        plan.dp_join()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(steps):
            plan.step_fresh_dp(True)
        plan.dp_join()
        a1.record(stream)
        barrier()
        ms = max_over_ranks(a0.elapsed_time(a1)) / steps
        return {"layer_sizes": sizes, "global_batch": n_global, "batch_per_gpu": count, "n_gpus": world, "steps": steps,
                "ms_per_step": ms, "samples_per_s": n_global / (ms * 1e-3),
                "tflops_reference_count": flops_per_sample(sizes) * n_global / (ms * 1e-3) / 1e12,
                "allreduce_bytes": len(plan.grad_slab()) * 8, "scaling": "strong"}
    finally:
        plan.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="samples per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=0, help="samples per CPU-oracle step (0 = the GPU arm's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--big-cfg", dest="big_cfg", type=int, default=None, help="DMMA tile configuration for wide outputs (0 or 2)")
    ap.add_argument("--sk-bias", dest="sk_bias", type=int, default=None, help="stream-K cost in the decomposition model, percent (default: 103 for short unit ranges, 96 for long ones)")
    ap.add_argument("--sk-rotate", dest="sk_rotate", type=int, default=None,
                    help="stream-K launches walk each tile segment's k-tiles in rotated order so CTAs sweep k in lockstep (L2 reuse): 0 / 1")
    ap.add_argument("--persist-tiles", dest="persist_tiles", type=int, default=None,
                    help="multi-wave contractions as persistent CTAs over whole tiles: 0 off, 1 contiguous, 2 wave order "
                         "(needs a library built with -DAF_PERSIST_TILES=1; the mode measured slower and is compiled out by default)")
    ap.add_argument("--deterministic", type=int, default=None, choices=[0, 1],
                    help="1: af_set_option deterministic (run-to-run bit-identical sums: whole tiles per CTA, ordered column sums); default 0")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="af_set_option(NAME, VALUE) before the run (A/B switches)")
    ap.add_argument("--dp-mode", dest="dp_mode", default="deferred", choices=sorted(DP_MODES),
                    help="N > 1: how the library schedules the sum over ranks (include/autofunc_b200.h AF_DP_*)")
    ap.add_argument("--dp-max-ctas", dest="dp_max_ctas", type=int, default=None, help="CTA cap of the library's NCCL communicator (default 8)")
    ap.add_argument("--dp-reserve-sms", dest="dp_reserve_sms", type=int, default=None,
                    help="SMs left to the overlapped all-reduce when sizing stream-K launches (default = dp_max_ctas)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling block (n = 4096 and C5 n = 65536 split N ways)")
    ap.add_argument("--profile-step", action="store_true",
                    help="run only the warm-up and timed steps (no other modes, peak measurement, end-to-end or CPU legs): "
                         "the command to put under ncu for a launch list of the step")
    ap.add_argument("--trace-steps", dest="trace_steps", type=int, default=4,
                    help="how many of the last timed steps carry the per-launch event trace the roofline is computed from (0: none)")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample NVML clocks during the timed region (A/B of the sampler's cost)")
    ap.add_argument("--legacy-step", action="store_true", help="N = 1: zero_grads + fill + zero + af_mlp_step (four calls) instead of af_mlp_step_fresh_dp")
    ap.add_argument("--trace", action="store_true", help="print the last step's per-launch device times to stderr")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
