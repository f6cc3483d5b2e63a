"""Device parallel-tempering sampler: the drop-in for `reddemcee.PTSampler` as EMPEROR
constructs and drives it (emp.py:576-603 `_set_sampler_reddemcee`,
support/endit_reddemcee.scr:3 `sampler.run_mcmc(p1, nsweeps=, nsteps=, progress=)`) and
as the parent later reads it back (get_chain / get_log_like /
get_log_prob / betas / get_betas / get_tsw / acceptance_fraction ...).

State lives on the GPU for the whole run: `p[T,W,ndim]`, `logl[T,W]`, `logp[T,W]`.
Per sweep the host supplies the random draws (draws.py), the device does
nsteps x (propose, likelihood, accept) per half-ensemble and one swap sweep; the
ladder adaptation needs only the T-1 swap counts and runs on the host exactly like
the oracle (bit-identical beta history).

With `torch.distributed` initialised (one process per GPU) the temperature ladder
is sharded over the ranks, T/G temperatures each, interleaved by default (dist.py): the
stretch steps need no communication, the swap sweep all-gathers logL (and the swap draws
each rank generated for its own pairs) over NCCL and every rank replays the same plan.
"""
from __future__ import annotations

import time as _time
from typing import Optional

import numpy as np

from . import dist as _dist
from .draws import DrawStreams, SweepDraws, default_betas, draw_sweep, initial_positions


class PTSampler:
    def __init__(self, nwalkers: int, ndim: int, log_like, log_prior=None, ntemps: int = 1, pool=None,
                 backend=None, betas=None, tsw_history: bool = True, smd_history: bool = True,
                 adapt_tau: float = 1000, adapt_nu: float = 1, adapt_mode: int = 0, a: float = 2.0,
                 seed: Optional[int] = None, store: str = "device", thin_by: int = 1, adapt: bool = True, group=None,
                 layout: str = "strided", exchange: str = "allgather"):
        """`log_like` is the LikelihoodEngine (it carries the prior as well; `log_prior`,
        `pool` and `backend` are accepted for signature compatibility and ignored — the
        walkers are evaluated on the GPU, not through a multiprocessing pool)."""
        import torch
        from .engine import LikelihoodEngine
        if not isinstance(log_like, LikelihoodEngine):
            raise TypeError("log_like must be an astroemperor_b200.engine.LikelihoodEngine; the device "
                            "sampler cannot call Python likelihoods (no CPU fallback)")
        self.engine = log_like
        if ndim != self.engine.ndim:
            raise ValueError(f"ndim={ndim} but the engine's model has {self.engine.ndim} free parameters")
        if nwalkers % 2 or nwalkers < 2:
            raise ValueError("nwalkers must be even")
        if adapt_mode != 0:
            raise NotImplementedError("only adapt_mode=0 (equalise swap rates) is implemented")
        self.nwalkers, self.ndim, self.ntemps = int(nwalkers), int(ndim), int(ntemps)
        self.a = float(a)
        self.adapt_tau, self.adapt_nu, self.adapt_mode, self.adapt = adapt_tau, adapt_nu, adapt_mode, adapt
        self.tsw_history_bool, self.smd_history_bool = bool(tsw_history), bool(smd_history)
        self.betas = (np.array(betas, dtype=np.float64) if betas is not None
                      else default_betas(ndim, ntemps))
        if len(self.betas) != self.ntemps:
            raise ValueError(f"betas should have {ntemps} items")
        self.streams = DrawStreams(seed, self.ntemps)
        self.D_ = None
        self.store, self.thin_by = store, int(thin_by)
        self.torch = torch
        self.dev = self.engine.torch_device
        self.shard = _dist.LadderShard(self.ntemps, group=group, layout=layout)
        if exchange not in ("allgather", "p2p"):
            raise ValueError("exchange must be 'allgather' or 'p2p'")
        self.exchange = exchange
        self.iteration = 0  # sweeps done
        self.time = 0
        self._chain = self._ll = self._lp = None
        self._beta_hist, self._tsw_hist, self._smd_hist = [], [], []
        self._n_accepted = None
        self._n_steps = 0
        self.timings = {"draws": 0.0, "h2d": 0.0}

    # ------------------------------------------------------------------------------
    def initial_positions(self, spec, max_repeats: int = 100) -> np.ndarray:
        """set_init() + test_init() of the generated script (emp.py:617-684): redraw walkers whose
        prior is -inf, at most `max_repeats` rounds."""
        p0 = initial_positions(self.streams.init, spec, self.ntemps, self.nwalkers)
        for _ in range(max_repeats):
            lp = self.engine.my_prior(p0.reshape(-1, self.ndim)).reshape(self.ntemps, self.nwalkers)
            bad = ~np.isfinite(np.atleast_2d(lp))
            if not bad.any():
                break
            fresh = initial_positions(self.streams.init, spec, self.ntemps, self.nwalkers)
            p0[bad] = fresh[bad]
        else:
            print("COULDNT FIND VALID INITIAL POSITION")
        return p0

    def _upload(self, arr, dtype=None):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.dev, non_blocking=True)

    def _init_state(self, p0):
        torch = self.torch
        p0 = np.asarray(p0, dtype=np.float64)
        if p0.shape != (self.ntemps, self.nwalkers, self.ndim):
            raise ValueError(f"p0 must have shape {(self.ntemps, self.nwalkers, self.ndim)}")
        sl = self.shard.local_slice
        self.p = self._upload(p0[sl]).contiguous()
        Tl = self.shard.n_local
        self.logl = torch.empty((Tl, self.nwalkers), dtype=torch.float64, device=self.dev)
        self.logp = torch.empty_like(self.logl)
        self.engine.logl_batch_device(self.p.view(-1, self.ndim), self.logl.view(-1), self.logp.view(-1))
        self.accepted = torch.zeros((Tl, self.nwalkers), dtype=torch.uint8, device=self.dev)
        self._n_accepted = torch.zeros((Tl, self.nwalkers), dtype=torch.int64, device=self.dev)
        self._p_alt = torch.empty_like(self.p)
        self._ll_alt = torch.empty_like(self.logl)
        self._lp_alt = torch.empty_like(self.logp)
        self._src = torch.empty((self.ntemps, self.nwalkers), dtype=torch.int32, device=self.dev)
        self._n_acc = torch.zeros((max(self.ntemps - 1, 1),), dtype=torch.int32, device=self.dev)
        self._betas_dev = self._upload(self.betas)
        self.shard.warm_up(self.dev)

    def _alloc_store(self, nsweeps):
        """Make room for `nsweeps` more stored samples (capacity grows geometrically so repeated
        run_mcmc calls do not re-allocate and copy the chain every time)."""
        torch = self.torch
        n = (nsweeps + self.thin_by - 1) // self.thin_by
        Tl, W, nd = self.shard.n_local, self.nwalkers, self.ndim
        if self.store == "device":
            dev, pin = self.dev, False
        elif self.store == "host":
            dev, pin = "cpu", True
        else:
            self._chain = None
            return
        if self._chain is None:
            self._stored = 0
        need = self._stored + n
        cap = 0 if self._chain is None else self._chain.shape[0]
        if need <= cap:
            return
        new_cap = max(need, 2 * cap)
        kw = dict(dtype=torch.float64, device=dev)
        if pin:
            kw["pin_memory"] = True
        new = [torch.empty((new_cap, Tl, W, nd), **kw), torch.empty((new_cap, Tl, W), **kw),
               torch.empty((new_cap, Tl, W), **kw)]
        if self._chain is not None and self._stored:
            for dst, src in zip(new, (self._chain, self._ll, self._lp)):
                dst[: self._stored].copy_(src[: self._stored])
        self._chain, self._ll, self._lp = new

    # ------------------------------------------------------------------------------
    def stage_draws(self, draws: SweepDraws, pinned: bool = False):
        """Copy one sweep's draws to the device (async on the current stream).  `draws` holds the
        stretch draws of THIS rank's temperatures and the swap draws of the whole ladder.
        pinned=True packs all seven arrays into one reusable pinned staging buffer and issues a
        single H2D copy (double-buffered: the previous sweep may still be reading its draws)."""
        torch = self.torch
        fields = [(f, getattr(draws, f)) for f in SweepDraws.FIELDS
                  if not (f in ("perm", "lnu_swap") and self.ntemps < 2)]
        out = {f: None for f in SweepDraws.FIELDS}
        if not pinned:
            for f, a in fields:
                out[f] = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, non_blocking=True)
            return out
        offs, total = [], 0
        for f, a in fields:
            offs.append(total)
            total += (a.nbytes + 255) // 256 * 256
        if getattr(self, "_stage_cap", 0) < total:
            self._stage_host = [torch.empty(total, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._stage_dev = [torch.empty(total, dtype=torch.uint8, device=self.dev) for _ in range(2)]
            self._stage_evt = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_cap, self._stage_i = total, 0
        i = self._stage_i = 1 - self._stage_i
        self._stage_evt[i].synchronize()  # the H2D that last used this pinned buffer has finished
        host, dev = self._stage_host[i], self._stage_dev[i]
        hv = host.numpy()
        for (f, a), o in zip(fields, offs):
            hv[o:o + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
        dev[:total].copy_(host[:total], non_blocking=True)
        self._stage_evt[i].record()
        for (f, a), o in zip(fields, offs):
            tdt = torch.int32 if a.dtype == np.int32 else torch.float64
            out[f] = dev[o:o + a.nbytes].view(tdt).view(a.shape)
        return out

    def sweep_begin(self, draws):
        """Enqueue nsteps stretch steps of every local temperature + one swap sweep (asynchronous).
        `draws`: SweepDraws (host) or the dict `stage_draws` returned (already on the device)."""
        eng = self.engine
        sl = self.shard.local_slice
        if isinstance(draws, SweepDraws):
            t0 = _time.perf_counter()
            draws = self.stage_draws(draws)
            self.timings["h2d"] += _time.perf_counter() - t0
        nsteps = draws["zz"].shape[0]
        betas_loc = self._betas_dev[sl].contiguous()
        self._mark("start")
        for s in range(nsteps):
            eng.pt_stretch_step(self.p, self.logl, self.logp, betas_loc, draws["half_idx"][s], draws["zz"][s],
                                draws["rint"][s], draws["factors"][s], draws["lnu"][s], self.accepted)
            self._n_accepted += self.accepted
            self._n_steps += 1
        self._swap_draws = None
        if self.ntemps > 1:
            perm, lnu_swap = draws["perm"], draws["lnu_swap"]
            if self.shard.world > 1 and perm.shape[0] == self.shard.n_local:
                # every rank replays the whole plan: gather the pair rows (NCCL, behind the stretch kernels)
                perm = self.shard.all_gather_rows(perm)[: self.ntemps - 1]
                lnu_swap = self.shard.all_gather_rows(lnu_swap)[: self.ntemps - 1]
            self._swap_draws = (perm, lnu_swap)
        self._mark("stretch")

    def _mark(self, name):
        """Optional device-side phase timing (self.profile = True): CUDA events on the stream, read
        back by phase_times(); costs nothing when disabled."""
        if not getattr(self, "profile", False):
            return
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record()
        self._phase_events = getattr(self, "_phase_events", [])
        self._phase_events.append((name, ev))

    def phase_times(self):
        """{phase: total ms} since the last call (synchronises)."""
        self.torch.cuda.synchronize(self.dev)
        out, evs = {}, getattr(self, "_phase_events", [])
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            if n1 != "start":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._phase_events = []
        return out

    def sweep_end(self):
        """Swap sweep of the sweep begun last (all-gather of logL when sharded, plan, row exchange),
        then the ladder adaptation on the host from the T-1 swap counts (like the oracle:
        bit-identical beta history).  Everything that synchronises with the device lives here so
        that the caller can do host work (next sweep's draws) between sweep_begin and sweep_end."""
        n_acc = None
        if self._swap_draws is not None:
            perm, lnu_swap = self._swap_draws
            self._mark("host_gap")  # device idle time while the host was busy between begin and end
            logl_all = self.shard.all_gather_rows(self.logl)  # [T, W]; NCCL all-gather when sharded
            self._mark("allgather")
            self.engine.pt_swap_plan(logl_all, self._betas_dev, perm, lnu_swap, self._src, self._n_acc)
            self._mark("plan")
            self._apply_plan()
            self._mark("apply")
            if self.smd_history_bool and self.D_ is not None:
                self._record_smd()
            if not hasattr(self, "_n_acc_host"):
                self._n_acc_host = self.torch.empty(self._n_acc.shape, dtype=self.torch.int32).pin_memory()
            self._n_acc_host.copy_(self._n_acc, non_blocking=True)  # 4*(T-1) bytes
            self.torch.cuda.current_stream(self.dev).synchronize()
            n_acc = self._n_acc_host.numpy()[: self.ntemps - 1].copy()
        self.time += 1
        self.iteration += 1
        if n_acc is not None:
            ratios = n_acc / self.nwalkers
            if self.tsw_history_bool:
                self._tsw_hist.append(ratios)
            if self.adapt and self.ntemps > 2:
                self.betas = _adapt_ladder(self.betas, ratios, self.time, self.adapt_tau, self.adapt_nu)
                self._betas_dev = self._upload(self.betas)
        self._beta_hist.append(self.betas.copy())
        return n_acc

    def sweep(self, draws):
        """One full sweep: stretch steps + swap sweep + ladder adaptation."""
        self.sweep_begin(draws)
        return self.sweep_end()

    def _apply_plan(self):
        eng, sh = self.engine, self.shard
        if sh.world == 1:
            eng.pt_gather_rows(self._src.view(-1), self.p.view(-1, self.ndim), self.logl.view(-1),
                               self.logp.view(-1), self._p_alt.view(-1, self.ndim), self._ll_alt.view(-1),
                               self._lp_alt.view(-1))
        elif self.exchange == "p2p":
            # point-to-point exchange of exactly the rows that change rank (dist.exchange_rows)
            rows = self.torch.cat([self.p.view(-1, self.ndim), self.logl.view(-1, 1), self.logp.view(-1, 1)], 1)
            staged, src_local = sh.exchange_rows(self._src, rows, self.nwalkers)
            pin = staged[:, : self.ndim].contiguous()
            llin = staged[:, self.ndim].contiguous()
            lpin = staged[:, self.ndim + 1].contiguous()
            eng.pt_gather_rows(src_local, pin, llin, lpin, self._p_alt.view(-1, self.ndim),
                               self._ll_alt.view(-1), self._lp_alt.view(-1))
        else:
            # all-gather the ensemble over NVLink and gather locally: no host synchronisation, no
            # index compaction; T*W*(ndim+2)*8 bytes per sweep (150 MB at 256 x 2048 x 35, ~0.3 ms
            # of NVSwitch all-gather) buys back ~1.5 ms of latency-bound list building
            W, nl = self.nwalkers, sh.n_local * self.nwalkers
            p_all, ll_all, lp_all = sh.all_gather_flat(self.p.view(-1, self.ndim), self.logl.view(-1),
                                                       self.logp.view(-1))
            sg = self._src[sh.local_slice].reshape(-1).to(self.torch.int64)
            st, sw = sg // W, sg % W
            pos = (sh.owner_of_temp(st) * nl + sh.local_of_temp(st) * W + sw).to(self.torch.int32)
            eng.pt_gather_rows(pos, p_all, ll_all, lp_all, self._p_alt.view(-1, self.ndim),
                               self._ll_alt.view(-1), self._lp_alt.view(-1))
        self.p, self._p_alt = self._p_alt, self.p
        self.logl, self._ll_alt = self._ll_alt, self.logl
        self.logp, self._lp_alt = self._lp_alt, self.logp

    def draw(self, nsteps: int) -> SweepDraws:
        """Host draws of one sweep for this rank (draws.py)."""
        # sharded ladder: this rank draws the swap pairs whose row index is one of its temperatures
        # (row T-1 is padding); sweep_begin all-gathers them into ladder order
        rows = None if self.shard.world == 1 else range(self.ntemps)[self.shard.local_slice]
        return draw_sweep(self.streams, self.nwalkers, self.ndim, nsteps, self.a,
                          temps=self.shard.local_slice, swap=self.ntemps > 1, swap_rows=rows)

    def run_mcmc(self, p0, nsweeps: int, nsteps: int = 1, progress: bool = False, on_sweep=None):
        """sampler.run_mcmc(p1, nsweeps=, nsteps=, progress=) (support/endit_reddemcee.scr:3).
        While the device runs sweep k the host generates the draws of sweep k+1 (same thread: a
        background thread only fights the main thread for the GIL and stalls the launches);
        they are staged through a reusable pinned buffer.  `on_sweep(sampler, k)` is called after
        every sweep (bench.py uses it to read logL back)."""
        if p0 is not None:
            self._init_state(p0)
        elif not hasattr(self, "p"):
            raise ValueError("first call needs initial positions")
        self._alloc_store(nsweeps)
        it = range(nsweeps)
        if progress:
            try:
                from tqdm import tqdm
                it = tqdm(it, total=nsweeps)
            except Exception:
                pass
        def draw():
            t0 = _time.perf_counter()
            d = self.draw(nsteps)
            self.timings["draws"] += _time.perf_counter() - t0
            return d

        # the draws of the first sweep may already be staged: every call ends by drawing and staging one sweep
        # ahead (below), so that back-to-back calls — adaptation then production (support/endit_freeze1.scr),
        # warm-up then measurement — do not pay the 5 ms pipeline fill again
        staged, pre = None, getattr(self, "_prefetched", None)
        self._prefetched = None
        if nsweeps > 0:
            staged = pre[1] if (pre is not None and pre[0] == nsteps) else self.stage_draws(draw(), pinned=True)
        for k in it:
            self.sweep_begin(staged)
            # while the device runs this sweep the host draws the next one, packs it into the other pinned
            # buffer and enqueues its H2D copy behind the stretch kernels (double-buffered on both sides), so
            # nothing but the 4(T-1)-byte swap-count read sits between two sweeps
            staged = self.stage_draws(draw(), pinned=True)
            self.sweep_end()
            if self._chain is not None and (k % self.thin_by == 0):
                j = self._stored
                self._chain[j].copy_(self.p, non_blocking=True)
                self._ll[j].copy_(self.logl, non_blocking=True)
                self._lp[j].copy_(self.logp, non_blocking=True)
                self._stored += 1
            if on_sweep is not None:
                on_sweep(self, k)
        if nsweeps > 0:
            self._prefetched = (nsteps, staged)
        self.torch.cuda.synchronize(self.dev)
        return self.p

    def select_adjustment(self, mode):
        """reddemcee's ladder-adjustment selector as EMPEROR drives it: `support/endit_freeze1.scr:10` calls
        `sampler.select_adjustment('00')` after the adaptation burn-in to FREEZE the ladder for the production
        sweeps.  Other selectors are reddemcee internals that are not recoverable offline."""
        if str(mode) != "00":
            raise NotImplementedError(f"select_adjustment('{mode}'): only '00' (freeze the ladder) is implemented")
        self.adapt = False

    # ---- read-back API the reference's parent process uses --------
    def _get(self, buf, discard, thin, flat):
        if buf is None:
            raise RuntimeError("chain storage is disabled (store=None)")
        x = buf[: self._stored][discard::thin]
        x = self.shard.gather_to_all(x, dim=1) if self.shard.world > 1 else x
        x = x.cpu().numpy()
        x = np.swapaxes(x, 0, 1)  # [T, n, W, ...]
        if flat:
            x = x.reshape((x.shape[0], x.shape[1] * x.shape[2]) + x.shape[3:])
        return x

    def get_chain(self, discard=0, thin=1, flat=False):
        return self._get(self._chain, discard, thin, flat)

    def get_log_like(self, discard=0, thin=1, flat=False):
        return self._get(self._ll, discard, thin, flat)

    def get_log_prob(self, discard=0, thin=1, flat=False):
        """Tempered posterior beta*logL + logP per stored sample."""
        ll = self._get(self._ll, discard, thin, False)
        lp = self._get(self._lp, discard, thin, False)
        bh = np.array(self._beta_hist)[:: self.thin_by][discard::thin]  # [n, T]
        out = bh.T[:, :, None] * ll + lp
        return out.reshape(out.shape[0], -1) if flat else out

    def get_log_prior(self, discard=0, thin=1, flat=False):
        return self._get(self._lp, discard, thin, flat)

    def get_betas(self, discard=0):
        return np.array(self._beta_hist)[discard:]

    def get_tsw(self, discard=0):
        return np.array(self._tsw_hist)[discard:]

    def _record_smd(self):
        """Swap mean distance of this sweep (consumers emp.py:961-965, 1985-1990): for every
        temperature the mean, over the slots that received a walker from a hotter rung, of the
        distance between the walker that left and the one that arrived, in units of the prior
        widths `sampler.D_` (emp.py:595-602).  Device-side torch arithmetic on [T_loc, W, ndim]."""
        torch, sh, W = self.torch, self.shard, self.nwalkers
        if getattr(self, "_D_dev", None) is None or self._D_dev.shape[0] != self.ndim:
            self._D_dev = self._upload(np.asarray(self.D_, dtype=np.float64))
        src_t = self._src[sh.local_slice].to(torch.int64) // W                      # [T_loc, W]
        dest_t = torch.arange(self.ntemps, device=self.dev)[sh.local_slice].unsqueeze(1)
        came_down = src_t > dest_t
        dist = (((self.p - self._p_alt) / self._D_dev) ** 2).sum(-1).sqrt()           # new vs old content
        num = (dist * came_down).sum(1)
        cnt = came_down.sum(1).clamp(min=1)
        self._smd_hist.append(num / cnt)

    def get_smd(self, discard=0):
        """[n_sweeps, T-1] swap mean distances (rung j <-> j+1); needs `sampler.D_`."""
        if not self._smd_hist:
            return np.zeros((0, max(self.ntemps - 1, 0)))
        x = self.torch.stack(self._smd_hist)                                         # [n, T_loc]
        x = self.shard.gather_to_all(x, dim=1) if self.shard.world > 1 else x
        return x.cpu().numpy()[discard:, : self.ntemps - 1]

    @property
    def acceptance_fraction(self):
        a = self._n_accepted.double() / max(self._n_steps, 1)
        if self.shard.world > 1:
            a = self.shard.gather_to_all(a, dim=0)
        return a.cpu().numpy()

    def state_numpy(self):
        """(p, logl, logp) of the whole ladder as NumPy arrays."""
        p, ll, lp = self.p, self.logl, self.logp
        if self.shard.world > 1:
            p, ll, lp = (self.shard.gather_to_all(x, dim=0) for x in (p, ll, lp))
        return p.cpu().numpy(), ll.cpu().numpy(), lp.cpu().numpy()

    # ---- post-run reductions (emp.py:1375-1385, 1432-1447; host NumPy, postproc.py) ----------
    def get_evidence_ti(self, discard=0, pchip=False):
        """Thermodynamic-integration log-evidence: integral of <logL>_beta over beta (the estimator
        emp.py:1432-1447 falls back to), trapezoid or, with `pchip=True`, monotone cubic interpolation."""
        from .postproc import evidence_ti
        return evidence_ti(self.get_log_like(discard=discard), self.betas, pchip=pchip)

    def get_evidence_ss(self, discard=0, pchip=False):
        """Stepping-stone log-evidence with a batch-means error."""
        from .postproc import evidence_ss
        return evidence_ss(self.get_log_like(discard=discard), self.betas)

    def get_evidence_hybrid(self, discard=0, pchip=False):
        """reddemcee's 'hybrid' estimator is not recoverable offline: this returns the
        stepping-stone value with the TI/SS discrepancy added in quadrature to its error."""
        z_ss, e_ss = self.get_evidence_ss(discard=discard)
        z_ti, e_ti = self.get_evidence_ti(discard=discard, pchip=pchip)
        err = float(np.sqrt(np.nan_to_num(e_ss) ** 2 + (z_ss - z_ti) ** 2))
        return z_ss, err

    def get_autocorr_time(self, discard=0, thin=1, quiet=False, tol=50, c=5):
        """[T, ndim] integrated autocorrelation times of the stored chains (emcee's estimator),
        in units of stored samples x thin."""
        from .postproc import integrated_time
        ch = self.get_chain(discard=discard, thin=thin)  # [T, n, W, ndim]
        return np.array([thin * integrated_time(ch[t], c=c, tol=tol, quiet=quiet) for t in range(ch.shape[0])])

    @property
    def backend(self):
        """What EMPEROR's generated save section reads after the run (emp.py:727-761): `sampler.backend.iteration`,
        `.tsw_history[_bool]`, `.smd_history[_bool]` and, per temperature, `sampler.backend[t].iteration`,
        `.get_chain()`, `.get_log_like()`, `.get_log_prob()`, `.get_betas()`, `.accepted`.  A read-only view: the
        arrays are pulled from the device (and gathered over the ranks) once per view."""
        return _BackendView(self)

    def save_backend(self, name, discard=0):
        """Chain sink in the layout EMPEROR writes after a run (emp.py:722-762)."""
        from .postproc import save_backend
        return save_backend(self, name, discard=discard)


class _TemperatureBackend:
    """`sampler.backend[t]` of reddemcee as EMPEROR reads it (emp.py:749-761)."""

    def __init__(self, view, t):
        self._v, self._t = view, t

    @property
    def iteration(self):
        return self._v.iteration

    def get_chain(self):
        return self._v._chain[self._t]          # [iteration, W, ndim]

    def get_log_like(self):
        return self._v._ll[self._t]             # [iteration, W]

    def get_log_prob(self):
        return self._v._lpost[self._t]          # tempered posterior beta*logL + logP, [iteration, W]

    def get_betas(self):
        return self._v._betas[:, self._t]       # [iteration]

    @property
    def accepted(self):
        return self._v._accepted[self._t]       # accepted moves per walker, [W]


class _BackendView:
    def __init__(self, s: "PTSampler"):
        self._chain = s.get_chain()
        self._ll = s.get_log_like()
        self._lpost = s.get_log_prob()
        self._betas = s.get_betas()[:: s.thin_by]
        self._accepted = np.rint(s.acceptance_fraction * max(s._n_steps, 1)).astype(np.int64)
        self.iteration = self._chain.shape[1]
        self.ntemps = s.ntemps
        self.tsw_history_bool, self.smd_history_bool = s.tsw_history_bool, s.smd_history_bool
        self.tsw_history = s.get_tsw()
        self.smd_history = s.get_smd() if s.smd_history_bool else np.zeros((0, max(s.ntemps - 1, 0)))

    def __getitem__(self, t):
        if not -self.ntemps <= t < self.ntemps:
            raise IndexError(t)
        return _TemperatureBackend(self, t % self.ntemps)

    def __len__(self):
        return self.ntemps


def _adapt_ladder(betas, ratios, time, adapt_tau, adapt_nu):
    """Vousden, Farr & Mandel (2016) ladder dynamics in reddemcee's (adapt_tau, adapt_nu)
    parameterisation; same arithmetic as oracle/pt_oracle.py::adapt_ladder."""
    betas = betas.copy()
    decay = adapt_tau / (time + adapt_tau)
    kappa = decay / adapt_nu
    dSs = kappa * (ratios[:-1] - ratios[1:])
    deltaTs = np.diff(1 / betas[:-1])
    deltaTs = deltaTs * np.exp(dSs)
    betas[1:-1] = 1 / (np.cumsum(deltaTs) + 1 / betas[0])
    return betas
