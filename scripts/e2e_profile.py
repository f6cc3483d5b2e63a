"""Host-side phase breakdown of the e2e sweep loop (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from astroemperor_b200.engine import LikelihoodEngine
from astroemperor_b200.sampler import PTSampler
w, data, spec = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
eng = LikelihoodEngine(spec, data.t, data.y, data.yerr, data.flag)
samp = PTSampler(w["W"], spec.ndim, eng, ntemps=w["T"], seed=1, store=None)
samp._init_state(samp.initial_positions(spec))
acc = {}
def tk(name, t0):
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
ll_host = torch.empty((w["T"], w["W"]), dtype=torch.float64).pin_memory()
st = samp.stage_draws(samp.draw(1), pinned=True)
for it in range(9):
    if it == 3: acc.clear(); torch.cuda.synchronize(); T0 = time.perf_counter()
    t0 = time.perf_counter(); samp.sweep_begin(st); tk("enqueue", t0)
    t0 = time.perf_counter(); d = samp.draw(1); tk("draw", t0)
    t0 = time.perf_counter(); st = samp.stage_draws(d, pinned=True); tk("stage", t0)
    t0 = time.perf_counter(); samp.sweep_end(); tk("wait_end", t0)
    t0 = time.perf_counter(); ll_host.copy_(samp.logl, non_blocking=True); torch.cuda.current_stream().synchronize(); tk("readback", t0)
tot = time.perf_counter() - T0
print({k: round(v / 6 * 1e3, 3) for k, v in acc.items()}, "ms/sweep; total", round(tot / 6 * 1e3, 3), "cpu count", os.cpu_count())
# steady-state cost of the public call with and without chain storage
for store in (None, "device"):
    s2 = PTSampler(w["W"], spec.ndim, eng, ntemps=w["T"], seed=1, store=store)
    s2.run_mcmc(samp.initial_positions(spec), nsweeps=3, nsteps=1)
    for n in (5, 30):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s2.run_mcmc(None, nsweeps=n, nsteps=1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"run_mcmc store={store} nsweeps={n}: {dt / n * 1e3:.3f} ms/sweep")
