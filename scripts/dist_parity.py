"""Multi-GPU check (run under torchrun): a ladder sharded over N GPUs must produce the same
chains, bit for bit, as the same ladder on one GPU (same seed -> same per-temperature draws).
Rank 0 also runs the single-GPU reference and compares."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    td.init_process_group("nccl", device_id=torch.device("cuda", lr))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from conftest import load_golden
    from astroemperor_b200.engine import LikelihoodEngine
    from astroemperor_b200.sampler import PTSampler
    g, spec = load_golden("c2_synth3p_2ins_n400")
    T, W, nsweeps, nsteps = 4 * world, 64, 6, 2
    eng = LikelihoodEngine(spec, g["t"], g["y"], g["yerr"], g["flag"], device=lr)
    exchange = sys.argv[1] if len(sys.argv) > 1 else "peer"
    samp = PTSampler(W, eng.ndim, eng, ntemps=T, seed=77, exchange=exchange)
    samp.D_ = spec.prior_widths()
    obj = [samp.initial_positions(spec) if rank == 0 else None]
    td.broadcast_object_list(obj, src=0)
    p0 = obj[0]
    samp.run_mcmc(p0, nsweeps=nsweeps, nsteps=nsteps)
    chain, ll = samp.get_chain(), samp.get_log_like()
    betas, tsw = samp.get_betas(), samp.get_tsw()
    ok = True
    if rank == 0:
        # same ladder on ONE gpu: a sampler whose shard is the whole ladder (its own 1-rank group)
        pass
    td.barrier()
    # single-process run on every rank with a fresh 1-rank view: use group of size 1
    solo_group = None
    for r in range(world):
        grp = td.new_group([r])
        if r == rank:
            solo_group = grp
    solo = PTSampler(W, eng.ndim, eng, ntemps=T, seed=77, group=solo_group)
    solo.D_ = spec.prior_widths()
    solo.run_mcmc(p0, nsweeps=nsweeps, nsteps=nsteps)
    c1, l1 = solo.get_chain(), solo.get_log_like()
    same = np.array_equal(c1, chain) and np.array_equal(l1, ll) and np.array_equal(solo.get_betas(), betas) \
        and np.array_equal(solo.get_tsw(), tsw) and np.allclose(solo.get_smd(), samp.get_smd(), rtol=1e-12)
    res = torch.tensor([1 if same else 0], device="cuda")
    td.all_reduce(res, op=td.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_parity world={world} T={T} W={W} exchange={exchange}: sharded == single-GPU chains: {bool(res.item())}; "
              f"swap rates {tsw.mean(axis=0).round(3)}")
    td.destroy_process_group()
    sys.exit(0 if res.item() == 1 else 1)

if __name__ == "__main__":
    main()
