"""Quick device timing of the likelihood kernel on the BASELINE shapes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astroemperor_b200.synth import make_synthetic_rv
from astroemperor_b200.data import from_instrument_tables
from astroemperor_b200.frontend import default_spec
from astroemperor_b200.engine import LikelihoodEngine

def run(name, seed, n, nins, kplan, ma_global, param, n_eval):
    data = from_instrument_tables(make_synthetic_rv(seed=seed, n=n, nins=nins, kplan=kplan, ma=ma_global is not None))
    spec = default_spec(data, kplan=kplan, parameterisation=param, moav=None if ma_global is None else dict(order=1, **{'global': ma_global}))
    eng = LikelihoodEngine(spec, data.t, data.y, data.yerr, data.flag)
    rng = np.random.default_rng(0)
    fp = spec.free_params()
    lo = np.array([p.limits[0] for p in fp]); hi = np.array([p.limits[1] for p in fp])
    th = rng.uniform(lo, hi, size=(n_eval, len(fp)))
    thd = torch.as_tensor(th, device='cuda')
    ll, lp = eng.logl_batch_device(thd)
    torch.cuda.synchronize()
    frac = float(torch.isfinite(lp).double().mean())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): eng.logl_batch_device(thd, ll, lp)
    ev0.record()
    reps = 5
    for _ in range(reps): eng.logl_batch_device(thd, ll, lp)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    act = frac * n_eval
    print(f"{name}: n_eval={n_eval} N={n} K={kplan} finite_prior={frac:.3f} {ms:.3f} ms/launch "
          f"-> {n_eval*n/ms*1e-6:.3f} G pt-evals/s (all), {act*n*kplan/ms*1e-6:.3f} G planet-pt/s (active)", flush=True)

if __name__ == '__main__':
    run('C1', 1, 256, 1, 1, None, 0, 200)
    run('C2', 2, 2000, 2, 3, None, 1, 5120)
    run('C4-noMA', 4, 10000, 4, 5, False, 0, 65536)
    run('C4-globalMA', 4, 10000, 4, 5, True, 0, 65536)
    run('C4-globalMA-p1', 4, 10000, 4, 5, True, 1, 65536)
