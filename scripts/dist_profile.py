"""Per-phase wall-clock breakdown of one sweep on a sharded ladder (development aid; torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
import bench

def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.sampler import PTSampler
    w, data, spec = bench.build_workload("c4")
    T, W = w["T"] * world, w["W"]
    eng = LikelihoodEngine(spec, data.t, data.y, data.yerr, data.flag, device=lr)
    samp = PTSampler(W, spec.ndim, eng, ntemps=T, seed=1, store=None)
    obj = [samp.initial_positions(spec) if rank == 0 else None]
    if world > 1: td.broadcast_object_list(obj, src=0)
    samp._init_state(obj[0])
    sync = lambda: torch.cuda.synchronize()
    acc = {}
    def tick(name, t0):
        sync(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
    for it in range(8):
        d = samp.stage_draws(samp.draw(1)); sync()
        if it == 3: acc.clear()
        t0 = time.perf_counter()
        sl = samp.shard.local_slice
        eng.pt_stretch_step(samp.p, samp.logl, samp.logp, samp._betas_dev[sl], d["half_idx"][0], d["zz"][0], d["rint"][0], d["factors"][0], d["lnu"][0], samp.accepted)
        tick("stretch", t0); t0 = time.perf_counter()
        logl_all = samp.shard.all_gather_rows(samp.logl)
        tick("allgather", t0); t0 = time.perf_counter()
        eng.pt_swap_plan(logl_all, samp._betas_dev, d["perm"], d["lnu_swap"], samp._src, samp._n_acc)
        tick("plan", t0); t0 = time.perf_counter()
        samp._apply_plan()
        tick("apply", t0); t0 = time.perf_counter()
        n = samp._n_acc.cpu().numpy()
        tick("nacc_d2h", t0)
    if rank == 0:
        print({k: round(v / 5 * 1e3, 3) for k, v in acc.items()}, "ms per sweep (avg of 5)")
    if world > 1: td.destroy_process_group()
main()
