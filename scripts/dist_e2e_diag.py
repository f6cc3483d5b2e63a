"""Multi-GPU end-to-end diagnosis (run under torchrun): device-side phase times of the run_mcmc loop with
variations (chain store off / on, logL read-back off / on)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
import bench

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    td.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cx = bench.Ctx()
    wl = bench.Workload("c4")
    for store in (None, "host", "device"):
        for readback in (False, True):
            eng, samp, T = bench.make_sampler(cx, wl, total_sweeps=80, store=store)
            W = wl.w["W"]
            samp.run_mcmc(None, nsweeps=30, nsteps=1)
            ll_host = torch.empty((samp.shard.n_local, W), dtype=torch.float64).pin_memory()
            ev = [torch.cuda.Event(), torch.cuda.Event()]
            def rb(s, k):
                i = k & 1
                ll_host.copy_(s.logl, non_blocking=True); ev[i].record()
                if k >= 1: ev[1 - i].synchronize()
            samp.run_mcmc(None, nsweeps=3, nsteps=1, on_sweep=rb if readback else None)
            samp.profile = True; samp.phase_times()
            samp.timings = {"draws": 0.0, "h2d": 0.0}
            cx.barrier()
            t0 = time.perf_counter()
            samp.run_mcmc(None, nsweeps=10, nsteps=1, on_sweep=rb if readback else None)
            samp._sync_store(); cx.barrier()
            dt = (time.perf_counter() - t0) * 100
            evs = list(getattr(samp, "_phase_events", []))
            ph = samp.phase_times()
            gaps = [e0.elapsed_time(e1) for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]) if n1 == "start"]
            if rank == 0:
                print(f"store={store} readback={readback}: wall {dt:.2f} ms/sweep; device phases/sweep "
                      f"{ {k: round(v / 10, 3) for k, v in ph.items()} } gap-before-start {np.mean(gaps):.3f} ms; "
                      f"host { {k: round(v * 100, 3) for k, v in samp.timings.items()} }", flush=True)
            del samp; eng.close()
    td.destroy_process_group()

if __name__ == "__main__":
    main()
