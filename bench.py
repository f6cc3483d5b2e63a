#!/usr/bin/env python3
"""bench.py — throughput of the EMPEROR hot path on B200 (one JSON line on rank 0).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Metric (BASELINE.json): logL evals/sec x datapoints = walkers x temps x datapoints / s.
Workload (config.workload): BASELINE configs[3] — synthetic 5-planet RV, 4 instruments +
global MA(1) noise, 10 000 points, 32 temperatures x 2048 walkers per GPU (weak scaling: N
GPUs hold 32*N temperatures, interleaved over the ranks, swap sweep every step).

A "step" is one parallel-tempering sweep with nsteps=1: red/blue stretch move of every walker
of every temperature (propose -> batched likelihood+prior -> accept, per half), the hot->cold
temperature-swap sweep (logL all-gather over NCCL when N > 1) and the ladder adaptation.

  value : all draws of the timed steps already resident in HBM when the clock starts.
  e2e   : the user-level loop: draws generated on the host every step, staged in pinned
          memory, copied H2D, the step, and a D2H read of logL[T,W] (+ swap counts).
  roofline : the likelihood kernel, timed per launch with CUDA events inside the timed region,
          algorithmic FP64 flops F(K) = 230 K + 30 (+25 MA) per walker-datapoint (SURVEY.md §8d
          row D4) against the FP64 FMA peak measured in the same run (emp_fp64_peak).
  cpu_baseline : the oracle (NumPy port of the generated script + C Kepler solver), one call per
          walker through multiprocessing.Pool(all cores) like support/pools/01.pool, on a
          bounded sample of the same workload.
"""
import argparse
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
    # name: (seed, N, nins, K, ma_global, parameterisation, T per GPU, W)
    "c4": dict(seed=4, n=10000, nins=4, kplan=5, ma="global", param=0, T=32, W=2048,
               desc="BASELINE configs[3]: synthetic 5-planet RV, 4 instruments + MA(1) (global recurrence), "
                    "10k points, 32 temps x 2048 walkers per GPU"),
    "c4noop": dict(seed=4, n=10000, nins=4, kplan=5, ma="perins", param=0, T=32, W=2048,
                   desc="configs[3] with the reference's default per-instrument MA template (no-op on logL)"),
    "c2": dict(seed=2, n=2000, nins=2, kplan=3, ma=None, param=1, T=10, W=512,
               desc="BASELINE configs[1]: synthetic 3-planet RV, 2 instruments with jitter, 2k points, "
                    "10 temps x 512 walkers"),
    "c5": dict(seed=5, n=50000, nins=4, kplan=5, ma=None, param=0, T=8, W=8192,
               desc="BASELINE configs[4]: 5 Keplerians, 50k points, 8 temps x 8192 walkers per GPU"),
    "tiny": dict(seed=1, n=400, nins=2, kplan=2, ma="global", param=0, T=2, W=64, desc="smoke-sized"),
}


def flops_per_point(kplan, ma):
    return 230.0 * kplan + 30.0 + (25.0 if ma == "global" else 0.0)


def build_workload(name):
    from astroemperor_b200.data import from_instrument_tables
    from astroemperor_b200.frontend import default_spec
    from astroemperor_b200.synth import make_synthetic_rv
    w = WORKLOADS[name]
    data = from_instrument_tables(make_synthetic_rv(seed=w["seed"], n=w["n"], nins=w["nins"], kplan=w["kplan"],
                                                    ma=w["ma"] is not None))
    moav = None if w["ma"] is None else {"order": 1, "global": w["ma"] == "global"}
    spec = default_spec(data, kplan=w["kplan"], parameterisation=w["param"], moav=moav)
    return w, data, spec


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
_ORC = None


def _cpu_init(name):
    global _ORC
    from oracle.rv_oracle import RVOracle
    w, data, spec = build_workload(name)
    _ORC = RVOracle(spec.compile(), data.t, data.y, data.yerr, data.flag)


def _cpu_eval(theta):
    lp = _ORC.my_prior(theta)
    if lp == -np.inf:
        return -np.inf, lp
    return _ORC.my_likelihood(theta), lp


def valid_thetas(spec, n, seed=0):
    """Walker positions with test_init semantics (emp.py:617-684): inside the prior support."""
    from astroemperor_b200.draws import initial_positions
    rng = np.random.RandomState(seed)
    return initial_positions(rng, spec, 1, n)[0]


def cpu_baseline(name, n_eval, cores=None, repeats=1):
    """Pool.map of the oracle's my_prior + my_likelihood over n_eval walkers (support/pools/01.pool)."""
    import multiprocessing as mp
    w, data, spec = build_workload(name)
    cores = cores or os.cpu_count()
    th = valid_thetas(spec, n_eval, seed=123)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(name,)) as pool:
        pool.map(_cpu_eval, list(th[: cores]))  # warm the workers
        best = np.inf
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_eval, list(th))
            best = min(best, time.perf_counter() - t0)
    return dict(value=n_eval * w["n"] / best, seconds=best, cores=cores, n_eval=n_eval)


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
    ap.add_argument("--solver", default="grid", choices=["grid", "kepler.py"],
                    help="Kepler solver of the likelihood kernel (A/B; the default is the product path)")
    ap.add_argument("--burn", type=int, default=40,
                    help="untimed sweeps before the warm-up so that the ensemble has left its uniform "
                         "initial state (a young chain proposes ~40%% of its moves outside the prior box, "
                         "which are never evaluated and would flatter the step time)")
    args = ap.parse_args()
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
    from astroemperor_b200.engine import LikelihoodEngine, fp64_peak_tflops
    from astroemperor_b200.sampler import PTSampler

    w, data, spec = build_workload(args.workload)
    T, W, N, ndim = w["T"] * world, w["W"], w["n"], spec.ndim
    eng = LikelihoodEngine(spec, data.t, data.y, data.yerr, data.flag, device=local_rank)
    eng.set_solver(args.solver)
    samp = PTSampler(W, ndim, eng, ntemps=T, seed=2026, store="device")
    p0 = samp.initial_positions(spec) if rank == 0 else None
    if world > 1:
        obj = [p0]
        td.broadcast_object_list(obj, src=0)
        p0 = obj[0]
    samp._init_state(p0)
    samp._alloc_store(2 * args.steps + args.warmup + 4)
    for _ in range(args.burn):
        samp.sweep(samp.draw(1))
    peak_tf = fp64_peak_tflops(local_rank)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def store():
        j = samp._stored
        samp._chain[j].copy_(samp.p, non_blocking=True)
        samp._ll[j].copy_(samp.logl, non_blocking=True)
        samp._lp[j].copy_(samp.logp, non_blocking=True)
        samp._stored += 1

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---------------- value: draws resident in HBM -----------------------------------------
    for _ in range(args.warmup):
        samp.sweep(samp.draw(1)); store()
    staged = [samp.stage_draws(samp.draw(1)) for _ in range(args.steps)]
    eng.set_timing(True)
    samp.profile = True
    samp.phase_times()
    c0 = eng.counters()
    l0 = eng.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        samp.sweep(staged[k]); store()
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    kern_ms, kern_n = eng.timing_collect()
    eng.set_timing(False)
    phases = samp.phase_times()
    samp.profile = False
    launches = eng.launch_count - l0
    c1 = eng.counters()
    per_rank = [[ms, kern_ms, float(c1["in_prior"] - c0["in_prior"])]]
    if world > 1:
        mine = torch.tensor(per_rank[0], dtype=torch.float64, device="cuda")
        allr = torch.empty((world, 3), dtype=torch.float64, device="cuda")
        td.all_gather_into_tensor(allr, mine)
        per_rank = allr.cpu().tolist()
        ms = max(r[0] for r in per_rank)  # max over ranks
    value = T * W * N * args.steps / (ms * 1e-3)
    del staged

    # ---------------- e2e: the user's call (PTSampler.run_mcmc) with host buffers -----------------
    # every sweep: host RNG draws -> pinned staging -> H2D -> step -> D2H read of logL[T,W]
    ll_host = torch.empty((samp.shard.n_local, W), dtype=torch.float64).pin_memory()
    io = {"d2h": 0}

    def read_back(s, k):
        ll_host.copy_(s.logl, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        io["d2h"] += ll_host.numel() * 8 + 4 * (T - 1)

    samp.run_mcmc(None, nsweeps=2, nsteps=1, on_sweep=read_back)  # warm the staging buffers / thread
    io["d2h"] = 0
    h2d = samp.draw(1).nbytes() * args.steps  # same size every sweep
    barrier()
    te0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    samp.run_mcmc(None, nsweeps=args.steps, nsteps=1, on_sweep=read_back)
    ee1.record()
    barrier()
    te1 = time.perf_counter()
    d2h = io["d2h"]
    e2e_ms = max(ee0.elapsed_time(ee1), (te1 - te0) * 1e3)  # host-bound loops are wall-clock bound
    if world > 1:
        tms = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        td.all_reduce(tms, op=td.ReduceOp.MAX)
        e2e_ms = float(tms.item())
    e2e_value = T * W * N * args.steps / (e2e_ms * 1e-3)

    # ---------------- the likelihood callable alone, host buffers (emp_logl_batch_host) -----------------
    # what an unmodified CPU sampler would call instead of Pool.map(my_likelihood): theta[n, ndim] in pageable
    # host memory -> H2D -> prior + likelihood kernels -> D2H of logL, logP; n = this rank's walkers
    th_host = samp.p.reshape(-1, ndim).cpu().numpy().copy()
    eng.logl_batch(th_host[:64])
    barrier()
    tc0 = time.perf_counter()
    for _ in range(max(args.steps // 2, 2)):
        eng.logl_batch(th_host)
    call_s = (time.perf_counter() - tc0) / max(args.steps // 2, 2)
    callable_value = len(th_host) * N / call_s * world  # every rank does the same amount concurrently

    if rank == 0:
        clocks.stop()
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return

    # ---------------- roofline of the likelihood kernel -----------------------------------------
    n_eval_active = c1["in_prior"] - c0["in_prior"]  # proposals whose likelihood was really evaluated
    F = flops_per_point(w["kplan"], w["ma"])
    alg_flops = F * n_eval_active * N  # this rank, whole timed region
    achieved_tf = alg_flops / (kern_ms * 1e-3) * 1e-12 if kern_ms > 0 else None
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
        with open(os.path.join(REPO, "profiles", "kernel_traffic.json")) as fh:
            tj = json.load(fh).get(args.workload)
        if tj:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "fp64", "kernel": "emp::logl_rv_kernel", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                "traffic_unit": "bytes per launch (HBM); the path is compute-bound, see DESIGN.md §4.1",
                "traffic_source": traffic_src,
                "note": "achieved = ALGORITHMIC flops (the oracle's operation count, SURVEY.md §8d row D4: 230 per "
                        "planet-point with libm calls at 20) / launch time; the kernel reaches the same root with "
                        "~35 FP64 + ~40 FP32 instructions per planet-point, so frac can exceed 1 — the executed-"
                        "instruction pipe utilisation (ncu) is in profiles/r01_notes.md",
                "peak_source": "FP64 FMA microbenchmark measured in this run (emp_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "launches": kern_n, "avg_launch_ms": kern_ms / max(kern_n, 1),
                "alg_flops_per_walker_point": F, "kernel_share_of_step": kern_ms / ms,
                "evaluated_fraction": n_eval_active / max(c1["proposals"] - c0["proposals"], 1)}

    out = {"metric": "logL evals/sec (walkers x temps x datapoints / s)", "value": value,
           "unit": "walker*temp*datapoint/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (seeded, SURVEY.md §8d row D2)",
           "config": {"workload": w["desc"], "name": args.workload, "n_points": N, "n_keplerians": w["kplan"],
                      "n_instruments": w["nins"], "ndim": ndim, "ntemps": T, "nwalkers": W,
                      "parallelism": f"temperature ladder sharded over {world} GPU(s)", "solver": args.solver,
                      "l2": "each step's inputs (draws 2.4 MB/step + state 18 MB) differ per step; the 280 KB "
                            "data set is L2-resident by design (re-read by every CTA)"},
           "logl_evals_per_s": T * W * args.steps / (ms * 1e-3),
           "e2e": {"value": e2e_value, "unit": "walker*temp*datapoint/s", "h2d_bytes_per_step": h2d // args.steps,
                   "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps,
                   "includes": "host RNG draws, pinned staging, H2D, step, D2H of logL[T,W]"},
           "callable_host": {"value": callable_value, "unit": "walker*temp*datapoint/s", "ms_per_call": call_s * 1e3,
                             "n_eval_per_call": int(len(th_host)),
                             "what": "emp_logl_batch_host on the current ensemble (all inside the prior): pageable "
                                     "host theta -> H2D -> prior + likelihood kernels -> D2H of logL, logP"},
           "gpu_launches": launches, "burn_in_sweeps": args.burn, "roofline": roofline,
           "clocks": clocks.summary(tw0, tw1),
           "phase_ms_per_step_rank0": {k: v / args.steps for k, v in phases.items()},
           "per_rank": [{"ms_per_step": r[0] / args.steps, "kernel_ms_per_step": r[1] / args.steps,
                         "evaluated_per_step": r[2] / args.steps} for r in per_rank],
           "acceptance_fraction": (c1["accepted"] - c0["accepted"]) / max(c1["proposals"] - c0["proposals"], 1)}

    if not args.no_cpu_baseline:
        cores = os.cpu_count()
        n_cpu = args.cpu_evals or max(cores * 600, 640)  # ~10 s of work on all host cores at C4
        cb = cpu_baseline(args.workload, n_cpu, cores)
        out["cpu_baseline"] = {"value": cb["value"], "unit": "walker*temp*datapoint/s", "cores": cores, "kind": "port",
                               "sample": f"{n_cpu} walkers x {N} points of the same workload through "
                                         f"multiprocessing.Pool({cores}) ({cb['seconds']:.1f} s)"}
    print(json.dumps(out))


def run_reference(args, rank):
    """The reference's own CPU implementation of the path: generated-script-equivalent NumPy
    (oracle port; kepler.py / reddemcee are not installable) mapped over walkers
    with multiprocessing.Pool(all cores) exactly like support/pools/01.pool."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
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
    value = n_cpu * w["n"] / sec
    world = int(os.environ.get("WORLD_SIZE", "1"))
    T = w["T"] * world
    out = {"impl": "reference", "metric": "logL evals/sec (walkers x temps x datapoints / s)", "value": value,
           "unit": "walker*temp*datapoint/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic (seeded, SURVEY.md §8d row D2)",
           "config": {"workload": w["desc"], "name": args.workload, "n_points": w["n"], "n_keplerians": w["kplan"],
                      "n_instruments": w["nins"], "ntemps": T, "nwalkers": w["W"]},
           "cpu_baseline": {"value": value, "unit": "walker*temp*datapoint/s", "cores": cores, "kind": "port",
                            "sample": f"each step = {n_cpu} walkers x {w['n']} points (of {T * w['W']}) through "
                                      f"multiprocessing.Pool({cores}); likelihood+prior per walker"},
           "e2e": {"value": value, "unit": "walker*temp*datapoint/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
