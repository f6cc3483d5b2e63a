#!/usr/bin/env python3
"""bench.py — throughput of the EMPEROR hot path on B200 (one JSON line on rank 0).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Metric (BASELINE.json): logL evals/sec x datapoints = walkers x temps x datapoints / s, counting only
likelihoods that are really EVALUATED: a proposal outside the prior support is rejected without an
evaluation (emcee semantics), on the GPU and in the reference alike, so it is not counted
(`value_nominal` keeps the every-proposal figure; the CPU arm evaluates 100 % of what it is given).
Workload (config.workload): BASELINE configs[3] — synthetic 5-planet RV, 4 instruments +
global MA(1) noise, 10 000 points, 32 temperatures x 2048 walkers per GPU (weak scaling: N
GPUs hold 32*N temperatures, interleaved over the ranks, swap sweep every step).

A "step" is one parallel-tempering sweep with nsteps=1: red/blue stretch move of every walker
of every temperature (proposal + prior, batched likelihood + Metropolis accept, per half), the hot->cold
temperature-swap sweep (logL all-gather over NCCL when N > 1), the ladder adaptation and the chain store —
six kernel launches, no host synchronisation.

  value : all draws of the timed steps already resident in HBM when the clock starts.
  e2e   : the user-level loop (PTSampler.run_mcmc): draws generated on the host every step, staged in
          pinned memory, copied H2D, the step (replayed from a CUDA graph), the chain streamed to pinned
          host memory and a D2H read of logL[T,W].
  roofline : the likelihood kernel, timed per launch with CUDA events inside the timed region,
          algorithmic FP64 flops F(K) = 230 K + 30 (+25 MA) per walker-datapoint (SURVEY.md §8d
          row D4) against the FP64 FMA peak measured in the same run (emp_fp64_peak); plus the
          executed-instruction figures of the committed ncu capture (profiles/kernel_ncu.json).
  cpu_baseline : the oracle (NumPy port of the generated script + C Kepler solver), one call per
          walker through multiprocessing.Pool(all cores) like support/pools/01.pool, on a
          bounded sample of the same workload; plus the single-core rate.
  configs : short legs of the other BASELINE configs (c1, c2, c3; c5 with 8 GPUs), each next to its CPU rate.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # synthetic: (seed, N, nins, K, MA, parameterisation); golden: real data shipped as a test fixture
    "c1": dict(golden="c1_51peg_k1_p0", kplan=1, ma=None, T=2, W=100,
               desc="BASELINE configs[0]: 51Peg RV (256 points), 1 Keplerian, reddemcee setup [2,100,500,1]"),
    "c2": dict(seed=2, n=2000, nins=2, kplan=3, ma=None, param=1, T=10, W=512,
               desc="BASELINE configs[1]: synthetic 3-planet RV, 2 instruments with jitter, 2k points, "
                    "10 temps x 512 walkers"),
    "c3": dict(golden="c3_hip21850_am_k2", kplan=2, ma=None, T=11, W=256,
               desc="BASELINE configs[2]: HIP21850 joint RV (97 points) + Hipparcos-Gaia astrometry (148 epochs), "
                    "2 Keplerians, 11 temps x 256 walkers"),
    "c4": dict(seed=4, n=10000, nins=4, kplan=5, ma="global", param=0, T=32, W=2048,
               desc="BASELINE configs[3]: synthetic 5-planet RV, 4 instruments + MA(1) (global recurrence), "
                    "10k points, 32 temps x 2048 walkers per GPU"),
    "c4noop": dict(seed=4, n=10000, nins=4, kplan=5, ma="perins", param=0, T=32, W=2048,
                   desc="configs[3] with the reference's default per-instrument MA template (no-op on logL)"),
    "c5": dict(seed=5, n=50000, nins=4, kplan=5, ma=None, param=0, T=8, W=8192, ladder_T=64,
               desc="BASELINE configs[4]: 5 Keplerians, 50k points, 8 temps x 8192 walkers per GPU"),
    "tiny": dict(seed=1, n=400, nins=2, kplan=2, ma="global", param=0, T=2, W=64, desc="smoke-sized"),
}
UNIT = "walker*temp*datapoint/s"
METRIC = "logL evals/sec (walkers x temps x datapoints / s)"


def flops_per_point(kplan, ma):
    return 230.0 * kplan + 30.0 + (25.0 if ma == "global" else 0.0)


class Workload:
    """Data + model of one BASELINE config."""

    def __init__(self, name):
        w = dict(WORKLOADS[name])
        self.name, self.w, self.am = name, w, None
        if "golden" in w:
            from astroemperor_b200.modelspec import ModelSpec
            d = os.path.join(REPO, "tests", "golden")
            g = np.load(os.path.join(d, w["golden"] + ".npz"))
            self.spec = ModelSpec.from_json(open(os.path.join(d, w["golden"] + ".json")).read())
            self.t, self.y, self.yerr, self.flag = g["t"], g["y"], g["yerr"], g["flag"]
            if any(k.startswith("am_") for k in g.files):
                self.am = {k[3:]: g[k] for k in g.files if k.startswith("am_")}
            w["n"] = len(self.t)
            w["nins"] = int(self.flag.max())
        else:
            from astroemperor_b200.data import from_instrument_tables
            from astroemperor_b200.frontend import default_spec
            from astroemperor_b200.synth import make_synthetic_rv
            data = from_instrument_tables(make_synthetic_rv(seed=w["seed"], n=w["n"], nins=w["nins"], kplan=w["kplan"],
                                                            ma=w["ma"] is not None))
            moav = None if w["ma"] is None else {"order": 1, "global": w["ma"] == "global"}
            self.spec = default_spec(data, kplan=w["kplan"], parameterisation=w["param"], moav=moav)
            self.t, self.y, self.yerr, self.flag = data.t, data.y, data.yerr, data.flag
        self.n = int(w["n"])
        # data points one likelihood evaluation touches (the unit of the metric): RV points (+ IAD epochs)
        self.n_units = self.n + (len(self.am["time_hipp"]) + len(self.am["time_gost"]) if self.am else 0)

    def engine(self, device):
        from astroemperor_b200.engine import LikelihoodEngine
        return LikelihoodEngine(self.spec, self.t, self.y, self.yerr, self.flag, am=self.am, device=device)

    def oracle(self):
        """(theta) -> (logl, logp) the way the generated script computes them (oracle/, CPU)."""
        from oracle.rv_oracle import RVOracle
        cm = self.spec.compile()
        # c4noop: the reference's default per-instrument MA template executes a Python loop whose writes are lost
        ro = RVOracle(cm, self.t, self.y, self.yerr, self.flag, run_noop_ma_loop=(self.w.get("ma") == "perins"))
        ao = None
        if self.am is not None:
            from oracle.am_oracle import AMOracle
            ao = AMOracle(cm, self.am)

        def ev(theta):
            lp = ro.my_prior(theta)
            if lp == -np.inf:
                return -np.inf, lp
            ll = ro.my_likelihood(theta)
            if ao is not None:
                with np.errstate(all="ignore"):
                    ll = float(ll + ao.loglike_AM(theta))
            return ll, lp
        return ev


def build_workload(name):  # kept for scripts/ that import it
    wl = Workload(name)
    return wl.w, wl, wl.spec


# ------------------------------------------------------------------ clocks ----
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [(ts, line) for ts, line in self.lines if t0 <= ts <= t1 + 0.25]
        if not inside and self.lines:  # timed region shorter than the sampling period: the nearest sample
            inside = [min(self.lines, key=lambda tl: abs(tl[0] - 0.5 * (t0 + t1)))]
        for ts, line in inside:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# -------------------------------------------------------------- CPU baseline ----
_EV = None
_ORACLE_LIB = None


def _cpu_init(name):
    global _EV
    _EV = Workload(name).oracle()


def _cpu_eval(theta):
    return _EV(theta)


def _worker_solver_lib(_):
    """Which shared object the worker's Kepler solver comes from (proof of what the CPU arm ran)."""
    out = [ln.split()[-1] for ln in open("/proc/self/maps") if "libemp_oracle" in ln]
    return out[0] if out else None


def valid_thetas(spec, n, seed=0):
    """Walker positions with test_init semantics (emp.py:617-684): inside the prior support, i.e. the CPU arm
    evaluates 100 % of what it is given — like for like with the GPU arm's count of EVALUATED likelihoods."""
    from astroemperor_b200.draws import initial_positions
    rng = np.random.RandomState(seed)
    return initial_positions(rng, spec, 1, n)[0]


def cpu_baseline(name, n_eval, cores=None, repeats=1):
    """Pool.map of the oracle's my_prior + my_likelihood over n_eval walkers (support/pools/01.pool);
    cores = 1: the same calls in this process, no pool."""
    import multiprocessing as mp
    wl = Workload(name)
    cores = cores or os.cpu_count()
    th = valid_thetas(wl.spec, n_eval, seed=123)
    if cores == 1:
        ev = wl.oracle()
        ev(th[0])
        t0 = time.perf_counter()
        for x in th:
            ev(x)
        best = time.perf_counter() - t0
        return dict(value=n_eval * wl.n_units / best, seconds=best, cores=1, n_eval=n_eval, solver_lib=None)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(name,)) as pool:
        pool.map(_cpu_eval, list(th[: cores]))  # warm the workers
        libs = sorted({x for x in pool.map(_worker_solver_lib, range(cores)) if x})
        best = np.inf
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_eval, list(th))
            best = min(best, time.perf_counter() - t0)
    return dict(value=n_eval * wl.n_units / best, seconds=best, cores=cores, n_eval=n_eval,
                solver_lib=libs[0] if libs else None)


def cpu_block(name, budget_s, cores, single_core=True):
    """cpu_baseline entry of one config: all-core Pool rate (+ single-core rate) on a sample sized for
    ~budget_s seconds of wall time."""
    wl = Workload(name)
    ev = wl.oracle()
    th = valid_thetas(wl.spec, 8, seed=7)
    ev(th[0])
    t0 = time.perf_counter()
    for x in th[:4]:
        ev(x)
    per_call = (time.perf_counter() - t0) / 4
    n_all = int(max(cores * 4, min(budget_s / per_call * cores * 0.6, 200000)))
    cb = cpu_baseline(name, n_all, cores)
    out = {"value": cb["value"], "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{n_all} walkers x {wl.n_units} points of the same workload (all inside the prior support) "
                     f"through multiprocessing.Pool({cores}) ({cb['seconds']:.1f} s)",
           "ms_per_call_one_core": per_call * 1e3, "solver_lib": cb["solver_lib"]}
    if single_core:
        n_one = int(max(4, min(budget_s * 0.3 / per_call, 20000)))
        c1 = cpu_baseline(name, n_one, 1)
        out["single_core"] = {"value": c1["value"], "unit": UNIT, "cores": 1,
                              "sample": f"{n_one} walkers, one process, no pool ({c1['seconds']:.1f} s)"}
    return out


# ------------------------------------------------------------------- GPU legs ----
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as td
        self.torch, self.td = torch, td
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier(self):
        if self.world > 1:
            self.td.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t.item())

    def gather(self, row):
        if self.world == 1:
            return [list(map(float, row))]
        mine = self.torch.tensor(row, dtype=self.torch.float64, device="cuda")
        allr = self.torch.empty((self.world, len(row)), dtype=self.torch.float64, device="cuda")
        self.td.all_gather_into_tensor(allr, mine)
        return allr.cpu().tolist()


LADDER_MODE = "range"   # --ladder: "range" (default) or "extend" (the default spacing carried on to T x world rungs)


def ladder(wl, world):
    """The ladder of a run on `world` GPUs: T = T_per_gpu x world rungs, geometric, between beta = 1 and the hottest
    rung of the config's own ladder (`ladder_T` rungs, default T_per_gpu, with the default spacing).  More GPUs
    buy a DENSER ladder over the same temperature range, so every GPU holds the same mix of cold and hot rungs
    (strided layout) and does the same work at every N: that is what weak scaling compares.  (Extending the
    default spacing to 8 x 32 rungs would only add rungs at beta < 1e-5 that sample the prior, propose 35 % of
    their moves outside its box and are never evaluated.)"""
    from astroemperor_b200.draws import default_betas
    T = wl.w["T"] * world
    if LADDER_MODE == "extend":
        return default_betas(wl.spec.ndim, T)
    b1 = default_betas(wl.spec.ndim, wl.w.get("ladder_T", wl.w["T"]))
    if len(b1) == T or T < 2:
        return b1[:T]
    return b1[-1] ** (np.arange(T, dtype=np.float64) / (T - 1))


def make_sampler(cx, wl, total_sweeps=0, seed=2026, store="host", **kw):
    from astroemperor_b200.sampler import PTSampler
    eng = wl.engine(cx.local_rank)
    T = wl.w["T"] * cx.world
    samp = PTSampler(wl.w["W"], wl.spec.ndim, eng, ntemps=T, seed=seed, store=store, betas=ladder(wl, cx.world), **kw)
    samp.D_ = wl.spec.prior_widths()
    p0 = samp.initial_positions(wl.spec) if cx.rank == 0 else None
    if cx.world > 1:
        obj = [p0]
        cx.td.broadcast_object_list(obj, src=0)
        p0 = obj[0]
    samp._init_state(p0)
    if total_sweeps:  # chain / history storage allocated up front: nothing is (re)allocated inside a timed region
        samp._alloc_store(total_sweeps)
        samp._alloc_hist(total_sweeps)
    return eng, samp, T


def run_e2e(cx, samp, eng, steps, T, W):
    """The user's call: run_mcmc with host draws every sweep.  Every step's sample (positions, logL, logP of the
    whole local ladder) is streamed to pinned host memory by the chain store (store='host', a copy stream one sweep
    behind the compute stream) — that is the step's device -> host read; the host consumes logL[T_loc, W] of the
    PREVIOUS sweep there while the current one runs.  (A separate copy of logL on the compute stream queued behind
    the 20 MB chain drain in the copy engine and stalled the next sweep by 0.2 ms on one GPU, 0.9 ms on eight.)"""
    torch = cx.torch
    io = {"d2h": 0, "sum": 0.0, "prev": None}

    def read_back(s, k):
        if s.store != "host" or not s._copy_done:
            return
        cur = (s._stored - 1, s._copy_done[-1][1])   # this sweep's sample index and the event of its drain
        if io["prev"] is not None:
            idx, ev = io["prev"]
            ev.synchronize()
            io["sum"] += float(s._ll[idx][0, 0])
        io["prev"] = cur

    # small ensembles: run_mcmc replays k sweeps per graph launch; every step's sample (positions, logL, logP) still
    # lands in pinned host memory through the chain store, and logL of every step is read there after the run
    k = samp._chunk_len(1)
    chunked = k > 1 and samp.store == "host"
    cb = None if chunked else read_back
    samp.run_mcmc(None, nsweeps=2 * k, nsteps=1, on_sweep=cb)  # warm the staging buffers and the graphs
    io["d2h"] = 0
    draws_bytes = samp.draw(1).nbytes()
    c0 = eng.counters()
    samp.timings = {"draws": 0.0, "h2d": 0.0}
    first = samp._stored
    cx.barrier()
    te0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    samp.run_mcmc(None, nsweeps=steps, nsteps=1, on_sweep=cb)
    samp._sync_store()
    if chunked:
        io["sum"] += float(samp._ll[first:first + steps].sum())  # the host copy of every step's logL[T, W]
    ee1.record()
    cx.barrier()
    te1 = time.perf_counter()
    c1 = eng.counters()
    ms = cx.max_over_ranks(max(ee0.elapsed_time(ee1), (te1 - te0) * 1e3))  # host-bound loops are wall-clock bound
    chain_bytes = samp.shard.n_local * W * (samp.ndim + 2) * 8 if samp.store == "host" else 0
    host = {k: v * 1e3 / steps for k, v in samp.timings.items() if v}
    return dict(ms=ms, host_ms_per_step=host, chunk=k if chunked else 1,
                evaluated=cx.sum_over_ranks(c1["in_prior"] - c0["in_prior"]),
                proposals=cx.sum_over_ranks(c1["proposals"] - c0["proposals"]),
                h2d=draws_bytes, d2h=io["d2h"] // steps + chain_bytes)


def run_leg(cx, name, steps, warmup, burn, cpu_budget):
    """A short leg of another BASELINE config: device-side rate with pre-generated draws and the e2e rate."""
    wl = Workload(name)
    eng, samp, T = make_sampler(cx, wl)
    W, N = wl.w["W"], wl.n_units
    k = samp._chunk_len(1)   # sweeps per graph launch (1 for the large ensembles and for sharded ladders)
    if k > 1:
        steps = (max(steps, 4 * k) + k - 1) // k * k
    total = burn + warmup + 2 * steps + 6 * k + 8
    samp._alloc_store(total)   # chain / history storage up front: nothing is (re)allocated inside a timed region
    samp._alloc_hist(total)
    samp.run_mcmc(None, nsweeps=burn + warmup, nsteps=1)
    # value: the draws of the timed steps are resident in HBM before the clock starts; per step (per chunk of k
    # steps) a device-to-device copy into the block the captured graph reads, then the graph replay
    if k > 1:
        samp._prefetched = None
        for _ in range(2):   # both device blocks' graphs are captured before the clock starts
            samp.run_chunk_resident(samp.draw_chunk_resident(k))
        pre = [samp.draw_chunk_resident(k) for _ in range(steps // k)]
    else:
        pre = [samp.draw_resident(1) for _ in range(steps)]
    cx.barrier()
    c0 = eng.counters()
    l0 = eng.launch_count
    ev0, ev1 = cx.torch.cuda.Event(enable_timing=True), cx.torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    ev0.record()
    if k > 1:
        for d in pre:
            samp.run_chunk_resident(d)
    else:
        for d in pre:
            samp.sweep_begin(samp.stage_resident(d))
    ev1.record()
    cx.barrier()
    tw1 = time.perf_counter()
    ms = cx.max_over_ranks(max(ev0.elapsed_time(ev1), 0.0))
    c1 = eng.counters()
    launches = eng.launch_count - l0
    evaluated = cx.sum_over_ranks(c1["in_prior"] - c0["in_prior"])
    proposals = cx.sum_over_ranks(c1["proposals"] - c0["proposals"])
    samp._prefetched = None
    del pre
    e = run_e2e(cx, samp, eng, steps, T, W)
    out = {"workload": wl.w["desc"], "ntemps": T, "nwalkers": W, "n_points": N, "ndim": wl.spec.ndim, "steps": steps,
           "value": evaluated * N / (ms * 1e-3), "value_nominal": proposals * N / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms / steps, "evaluated_fraction": evaluated / max(proposals, 1),
           "e2e": {"value": e["evaluated"] * N / (e["ms"] * 1e-3), "unit": UNIT, "ms_per_step": e["ms"] / steps,
                   "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"],
                   "host_ms_per_step_rank0": e["host_ms_per_step"]},
           "e2e_over_value": (e["evaluated"] / e["ms"]) / (evaluated / ms),
           "gpu_launches_per_step": launches / steps, "graph_captures": eng.graph_captures,
           "sweeps_per_graph_launch": k,
           "host_wall_ms_per_step": (tw1 - tw0) * 1e3 / steps}
    del samp
    eng.close()
    if cx.rank == 0 and cpu_budget > 0:
        out["cpu_baseline"] = cpu_block(name, cpu_budget, os.cpu_count())
    return out


def dist_parity(cx):
    """N > 1: a fixed small ladder run sharded over the ranks and unsharded on every rank must give the same
    chains bit for bit (same seed -> same per-temperature draw streams)."""
    td = cx.td
    wl = Workload("tiny")
    from astroemperor_b200.sampler import PTSampler
    T, W, nsweeps, nsteps = 4 * cx.world, 64, 6, 2
    eng = wl.engine(cx.local_rank)
    samp = PTSampler(W, wl.spec.ndim, eng, ntemps=T, seed=77)
    samp.D_ = wl.spec.prior_widths()
    obj = [samp.initial_positions(wl.spec) if cx.rank == 0 else None]
    td.broadcast_object_list(obj, src=0)
    p0 = obj[0]
    samp.run_mcmc(p0, nsweeps=nsweeps, nsteps=nsteps)
    got = [samp.get_chain(), samp.get_log_like(), samp.get_betas(), samp.get_tsw(), samp.get_smd()]
    solo_group = None
    for r in range(cx.world):
        grp = td.new_group([r])
        if r == cx.rank:
            solo_group = grp
    solo = PTSampler(W, wl.spec.ndim, eng, ntemps=T, seed=77, group=solo_group)
    solo.D_ = wl.spec.prior_widths()
    solo.run_mcmc(p0, nsweeps=nsweeps, nsteps=nsteps)
    ref = [solo.get_chain(), solo.get_log_like(), solo.get_betas(), solo.get_tsw(), solo.get_smd()]
    same = all(np.array_equal(a, b) for a, b in zip(got[:4], ref[:4])) and np.allclose(got[4], ref[4], rtol=1e-12)
    res = cx.torch.tensor([1 if same else 0], device="cuda")
    td.all_reduce(res, op=td.ReduceOp.MIN)
    h = hashlib.sha1(np.ascontiguousarray(got[0]).tobytes()).hexdigest()[:16]
    del samp, solo
    eng.close()
    return {"dist_parity": bool(res.item() == 1), "chain_sha1": h, "ntemps": T, "nwalkers": W,
            "sweeps": nsweeps, "nsteps": nsteps, "exchange": "peer"}


# ------------------------------------------------------------------- main ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-evals", type=int, default=0, help="walkers in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="auto", help="comma list of extra configs to run after the main one "
                    "(auto: c1,c2,c3 and c5 with 8 GPUs; none)")
    ap.add_argument("--solver", default="grid", choices=["grid", "kepler.py"],
                    help="Kepler solver of the likelihood kernel (A/B; the default is the product path)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "allgather"],
                    help="sharded swap: rows read from the owners' HBM over NVLink (CUDA IPC), or NCCL all-gather")
    ap.add_argument("--burn", type=int, default=40,
                    help="untimed sweeps before the warm-up so that the ensemble has left its uniform "
                         "initial state (a young chain proposes ~40%% of its moves outside the prior box, "
                         "which are never evaluated)")
    ap.add_argument("--ladder", default="range", choices=["range", "extend"],
                    help="N-GPU ladder: 'range' = T x N rungs over the temperature range of the config's own ladder "
                         "(denser ladder, the same work per GPU at every N); 'extend' = the default spacing carried on "
                         "to T x N rungs (round 1: the added rungs sample the prior and are mostly never evaluated)")
    args = ap.parse_args()
    global LADDER_MODE
    LADDER_MODE = args.ladder
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as td
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from astroemperor_b200.engine import fp64_peak_tflops
    cx = Ctx()

    parity = dist_parity(cx) if world > 1 else None

    wl = Workload(args.workload)
    w, N = wl.w, wl.n_units
    eng, samp, T = make_sampler(cx, wl, total_sweeps=args.burn + args.warmup + 3 * args.steps + 16,
                                exchange=args.exchange)
    W, ndim = w["W"], wl.spec.ndim
    samp_betas0 = ladder(wl, world)
    eng.set_solver(args.solver)
    samp.run_mcmc(None, nsweeps=args.burn, nsteps=1)
    peak_tf = fp64_peak_tflops(local_rank)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---------------- value: draws resident in HBM, the product path (CUDA-graph replays) -----------------
    for _ in range(args.warmup):
        samp.sweep_begin(samp.draw_staged(1))
    staged = [samp.draw_resident(1) for _ in range(args.steps + 2)]
    for d in staged[:2]:   # the graphs of this (state block, staging slot) pairing exist before the clock starts
        samp.sweep_begin(samp.stage_resident(d))
    c0 = eng.counters()
    l0 = eng.launch_count
    cx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    ev0.record()
    for d in staged[2:]:
        samp.sweep_begin(samp.stage_resident(d))
    ev1.record()
    cx.barrier()
    tw1 = time.perf_counter()
    ms_local = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    c1 = eng.counters()
    # ---------------- the same K steps again with per-launch CUDA events around the likelihood kernel ---------
    # (events cannot bracket kernels inside a replayed graph: in this pass the sweeps are launched directly, which is
    # up to 5 % slower per step on 8 GPUs; the pass yields the kernel's launch duration and its share of ITS step)
    staged = [samp.draw_resident(1) for _ in range(args.steps)]
    eng.set_timing(True)
    samp.profile = True
    samp.phase_times()
    ct0 = eng.counters()
    cx.barrier()
    tv0, tv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tv0.record()
    for d in staged:
        samp.sweep_begin(samp.stage_resident(d))
    tv1.record()
    cx.barrier()
    ms_timing_pass = tv0.elapsed_time(tv1)
    kern_ms, kern_n = eng.timing_collect()
    eng.set_timing(False)
    phases = samp.phase_times()
    samp.profile = False
    ct1 = eng.counters()
    per_rank = cx.gather([ms_local, kern_ms, float(c1["in_prior"] - c0["in_prior"]),
                          float(c1["proposals"] - c0["proposals"])])
    ms = max(r[0] for r in per_rank)  # max over ranks
    evaluated = sum(r[2] for r in per_rank)
    proposals = sum(r[3] for r in per_rank)
    value = evaluated * N / (ms * 1e-3)
    value_nominal = proposals * N / (ms * 1e-3)
    del staged
    samp._prefetched = None

    # ---------------- e2e: the user's call (PTSampler.run_mcmc) with host buffers -----------------
    e = run_e2e(cx, samp, eng, args.steps, T, W)
    e2e_value = e["evaluated"] * N / (e["ms"] * 1e-3)

    # ---------------- the likelihood callable alone, host buffers (emp_logl_batch_host) -----------------
    # what an unmodified CPU sampler would call instead of Pool.map(my_likelihood): theta[n, ndim] in pageable
    # host memory -> H2D -> prior + likelihood kernels -> D2H of logL, logP; n = this rank's walkers (all inside
    # the prior support: every one is evaluated)
    th_host = samp.p.reshape(-1, ndim).cpu().numpy().copy()
    eng.logl_batch(th_host[:64])
    cx.barrier()
    tc0 = time.perf_counter()
    for _ in range(max(args.steps // 2, 2)):
        eng.logl_batch(th_host)
    call_s = (time.perf_counter() - tc0) / max(args.steps // 2, 2)
    callable_value = len(th_host) * N / call_s * world  # every rank does the same amount concurrently

    if rank == 0:
        clocks.stop()
    nan_total = eng.counters()["nan"]
    del samp
    eng.close()

    # ---------------- the other BASELINE configs, short legs -----------------------------------------------
    if args.legs == "auto":
        legs = [c for c in ("c1", "c2", "c3") if c != args.workload] + (["c5"] if world == 8 else [])
    elif args.legs in ("none", ""):
        legs = []
    else:
        legs = [c for c in args.legs.split(",") if c]
    leg_out = {}
    for name in legs:
        big = name in ("c4", "c5", "c4noop")
        leg_out[name] = run_leg(cx, name, steps=5 if big else 50, warmup=3, burn=10 if big else 60,
                                cpu_budget=0 if args.no_cpu_baseline else 6)

    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return

    # ---------------- roofline of the likelihood kernel -----------------------------------------
    n_eval_active = ct1["in_prior"] - ct0["in_prior"]  # this rank's evaluated likelihoods in the kernel-timing pass
    F = flops_per_point(w["kplan"], w["ma"])
    alg_flops = F * n_eval_active * wl.n  # this rank, whole timed region
    achieved_tf = alg_flops / (kern_ms * 1e-3) * 1e-12 if kern_ms > 0 else None
    traffic, traffic_src, executed = None, None, {}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
        with open(os.path.join(REPO, "profiles", "kernel_traffic.json")) as fh:
            tj = json.load(fh).get(args.workload)
        if tj:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except (OSError, ValueError, KeyError):
        pass
    try:  # executed-instruction figures of the committed ncu capture of the same kernel on the same workload
        with open(os.path.join(REPO, "profiles", "kernel_ncu.json")) as fh:
            executed = json.load(fh).get(args.workload) or {}
    except (OSError, ValueError):
        pass
    roofline = {"bound": "fp64", "kernel": "emp::logl_rv_kernel", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                "traffic_unit": "bytes per launch (HBM); the path is compute-bound, see DESIGN.md §4.1",
                "traffic_source": traffic_src,
                "note": "frac = ALGORITHMIC flops (the oracle's operation count, SURVEY.md §8d row D4: 230 per "
                        "planet-point with libm calls at 20) / launch time / FP64 peak; it can exceed 1 because the "
                        "kernel reaches the same root with ~35 FP64 + ~40 FP32 instructions per planet-point. "
                        "frac_executed and the *_pct keys are what the hardware really did (ncu capture of this "
                        "kernel on this workload, profiles/kernel_ncu.json)",
                "peak_source": "FP64 FMA microbenchmark measured in this run (emp_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "launches": kern_n, "avg_launch_ms": kern_ms / max(kern_n, 1),
                "alg_flops_per_walker_point": F, "kernel_share_of_step": kern_ms / ms_timing_pass,
                "timing_pass_ms_per_step": ms_timing_pass / args.steps,
                "timing_pass": "the K timed steps repeated with per-launch CUDA events around the likelihood kernel "
                               "(sweeps launched directly instead of replayed from the graph)",
                "evaluated_fraction": n_eval_active / max(ct1["proposals"] - ct0["proposals"], 1)}
    for k in ("frac_executed", "fp64_pipe_pct", "issue_active_pct", "xu_pct", "fma_pipe_pct", "l2_to_sm_gbs",
              "l1_hit_pct", "source"):
        if k in executed:
            roofline[k if k != "source" else "executed_source"] = executed[k]

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic (seeded, SURVEY.md §8d row D2)",
           "config": {"workload": w["desc"], "name": args.workload, "n_points": wl.n, "n_keplerians": w["kplan"],
                      "n_instruments": w["nins"], "ndim": ndim, "ntemps": T, "nwalkers": W,
                      "parallelism": f"temperature ladder sharded over {world} GPU(s)", "solver": args.solver,
                      "ladder": (f"{T} rungs, geometric between beta = 1 and {float(samp_betas0[-1]):.3g} (the range of "
                                 f"the config's {w.get('ladder_T', w['T'])}-rung ladder at every N: more GPUs = a denser "
                                 "ladder, the same work per GPU), adapted on the device every sweep"
                                 if args.ladder == "range" else
                                 f"{T} rungs with the default spacing (coldest 1, hottest {float(samp_betas0[-1]):.3g}), "
                                 "adapted on the device every sweep"),
                      "exchange": args.exchange if world > 1 else None,
                      "l2": "each step's inputs (draws 2.4 MB/step + state 18 MB) differ per step; the 280 KB "
                            "data set is L2-resident by design (re-read by every CTA)"},
           "counts": "value / e2e count EVALUATED likelihoods (proposals inside the prior support) x datapoints; "
                     "value_nominal counts every proposal",
           "value_nominal": value_nominal,
           "evaluated_fraction": evaluated / max(proposals, 1),
           "logl_evals_per_s": evaluated / (ms * 1e-3),
           "pt_sweeps_per_s": args.steps / (ms * 1e-3),   # full PT steps (stretch + swap + adaptation) per second
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e["h2d"],
                   "d2h_bytes_per_step": e["d2h"], "ms_per_step": e["ms"] / args.steps,
                   "value_nominal": e["proposals"] * N / (e["ms"] * 1e-3),
                   "host_ms_per_step_rank0": e["host_ms_per_step"],
                   "includes": "host RNG draws, pinned staging, H2D, the sweep (CUDA graph), every step's sample "
                               "[T,W,ndim+2] (positions, logL, logP) streamed to pinned host memory, where the host "
                               "reads that step's logL one sweep later"},
           "callable_host": {"value": callable_value, "unit": UNIT, "ms_per_call": call_s * 1e3,
                             "n_eval_per_call": int(len(th_host)),
                             "what": "emp_logl_batch_host on the current ensemble (all inside the prior): pageable "
                                     "host theta -> H2D -> prior + likelihood kernels -> D2H of logL, logP"},
           "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps, "burn_in_sweeps": args.burn,
           "roofline": roofline, "clocks": clocks.summary(tw0, tw1),
           "phase_ms_per_step_rank0": {k: v / args.steps for k, v in phases.items()},
           "per_rank": [{"ms_per_step": r[0] / args.steps, "kernel_ms_per_step": r[1] / args.steps,
                         "evaluated_per_step": r[2] / args.steps} for r in per_rank],
           "acceptance_fraction": (c1["accepted"] - c0["accepted"]) / max(c1["proposals"] - c0["proposals"], 1),
           "nan_likelihoods": nan_total}
    if parity is not None:
        out.update(parity)
    if leg_out:
        out["configs"] = leg_out

    if not args.no_cpu_baseline:
        cores = os.cpu_count()
        if args.cpu_evals:
            cb = cpu_baseline(args.workload, args.cpu_evals, cores)
            out["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{args.cpu_evals} walkers x {N} points through Pool({cores})",
                                   "solver_lib": cb["solver_lib"]}
        else:
            out["cpu_baseline"] = cpu_block(args.workload, 12, cores)
        if args.workload == "c4":  # the reference's DEFAULT per-instrument MA template
            nn = max(cores, 8)
            cn = cpu_baseline("c4noop", nn, cores)
            out["cpu_baseline"]["default_ma_template"] = {
                "value": cn["value"], "unit": UNIT, "cores": cores,
                "sample": f"{nn} walkers x {N} points, moav00.model (a no-op on logL that costs "
                          f"{cn['seconds'] * 1e3 * cores / nn:.0f} ms per call), Pool({cores}) ({cn['seconds']:.1f} s)"}
    print(json.dumps(out))


def run_reference(args, rank):
    """The reference's own CPU implementation of the path: generated-script-equivalent NumPy
    (oracle port; kepler.py / reddemcee are not installable) mapped over walkers
    with multiprocessing.Pool(all cores) exactly like support/pools/01.pool."""
    if rank != 0:
        return
    wl = Workload(args.workload)
    w = wl.w
    cores = os.cpu_count()
    n_cpu = args.cpu_evals or max(cores * 128, 256)  # ~2 s per step on all host cores at C4
    vals = []
    t_all0 = time.perf_counter()
    for _ in range(max(args.warmup, 0)):
        cpu_baseline(args.workload, cores, cores)
    for _ in range(max(args.steps, 1)):
        vals.append(cpu_baseline(args.workload, n_cpu, cores))
        if time.perf_counter() - t_all0 > 150:
            break
    sec = float(np.mean([v["seconds"] for v in vals]))
    value = n_cpu * wl.n_units / sec
    world = int(os.environ.get("WORLD_SIZE", "1"))
    T = w["T"] * world
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": len(vals), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic (seeded, SURVEY.md §8d row D2)",
           "config": {"workload": w["desc"], "name": args.workload, "n_points": wl.n, "n_keplerians": w["kplan"],
                      "n_instruments": w["nins"], "ntemps": T, "nwalkers": w["W"]},
           "counts": "every walker of the sample is inside the prior support: all are evaluated",
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"each step = {n_cpu} walkers x {wl.n_units} points (of {T * w['W']}) through "
                                      f"multiprocessing.Pool({cores}); likelihood+prior per walker",
                            "ms_per_call_per_core": sec * 1e3 * cores / n_cpu,
                            "solver_lib": vals[-1]["solver_lib"]},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
