"""Multi-GPU end-to-end diagnosis (run under torchrun): is the gap between the device-resident rate and the
run_mcmc rate a fixed cost per call or a cost per sweep?  Times bench.run_e2e for several step counts and store modes."""
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
    for store in ("host", "device"):
        eng, samp, T = bench.make_sampler(cx, wl, total_sweeps=200, store=store)
        W = wl.w["W"]
        samp.run_mcmc(None, nsweeps=30, nsteps=1)
        for steps in (10, 30, 10):
            e = bench.run_e2e(cx, samp, eng, steps, T, W)
            # the same number of sweeps without the per-sweep read-back, device events only
            cx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            samp.run_mcmc(None, nsweeps=steps, nsteps=1)
            e1.record()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            samp._sync_store()
            t2 = time.perf_counter()
            cx.barrier()
            t3 = time.perf_counter()
            if rank == 0:
                print(f"store={store} steps={steps}: e2e total {e['ms']:.2f} ms = {e['ms'] / steps:.3f} ms/sweep | plain "
                      f"run_mcmc: device {e0.elapsed_time(e1):.2f} ms, wall {1e3 * (t1 - t0):.2f}, +sync_store "
                      f"{1e3 * (t2 - t1):.2f}, +barrier {1e3 * (t3 - t2):.2f}; host {e['host_ms_per_step']}", flush=True)
        del samp
        eng.close()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
