"""Device time of the swap-plan kernel alone (the serial term of a sharded sweep): T pairs x W walkers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden
from astroemperor_b200.engine import LikelihoodEngine

g, spec = load_golden("c1_51peg_k0")
eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"])
rng = np.random.default_rng(0)
for T, W in ([] if os.environ.get("PLAN_BENCH_PRODUCT_ONLY") else
             [(32, 2048), (64, 2048), (256, 2048), (64, 8192), (256, 512), (256, 1024), (256, 4096)]):
    ll = torch.as_tensor(rng.normal(size=(T, W)) * 5 - 3e4, device="cuda")
    betas = torch.as_tensor(np.geomspace(1, 1e-3, T), device="cuda")
    perm = torch.as_tensor(np.stack([np.stack([rng.permutation(W), rng.permutation(W)]) for _ in range(T - 1)]).astype(np.int32), device="cuda")
    lnu = torch.as_tensor(np.log(rng.uniform(size=(T - 1, W))), device="cuda")
    src = torch.empty((T, W), dtype=torch.int32, device="cuda")
    nacc = torch.empty((T - 1,), dtype=torch.int32, device="cuda")
    for _ in range(3):
        eng.pt_swap_plan(ll, betas, perm, lnu, src, nacc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.pt_swap_plan(ll, betas, perm, lnu, src, nacc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"plan T={T} W={W}: {ms*1e3:.1f} us per sweep, {ms*1e3/(T-1):.2f} us per pair, accept {nacc.float().mean().item()/W:.2f}", flush=True)

# the product path: pairs listed by their slot in the warmer row (register-resident plan kernel) + adaptation +
# plan application, timed as whole swap phases of a sampler that does no stretch steps (nsteps = 0)
from astroemperor_b200.sampler import PTSampler
for T, W in [(32, 2048), (64, 2048), (256, 2048), (64, 8192), (256, 512), (256, 1024), (256, 4096)]:
    samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=3, store=None, betas=np.geomspace(1, 1e-3, T))
    samp.D_ = spec.prior_widths()
    samp._init_state(samp.initial_positions(spec))
    samp._alloc_hist(64)
    pre = [samp.draw_resident(0) for _ in range(24)]
    for d in pre[:4]:
        samp.sweep_begin(samp.stage_resident(d))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for d in pre[4:]:
        samp.sweep_begin(samp.stage_resident(d))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"swap phase (sorted plan + adapt + apply) T={T} W={W}: {ms*1e3:.1f} us per sweep, {ms*1e3/(T-1):.2f} us per pair, "
          f"swap rate {samp.get_tsw().mean():.2f}", flush=True)
    del samp
