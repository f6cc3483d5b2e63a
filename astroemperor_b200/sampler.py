"""Device parallel-tempering sampler: the drop-in for `reddemcee.PTSampler` as EMPEROR
constructs and drives it (emp.py:576-603 `_set_sampler_reddemcee`,
support/endit_reddemcee.scr:3 `sampler.run_mcmc(p1, nsweeps=, nsteps=, progress=)`) and
as the parent later reads it back (get_chain / get_log_like /
get_log_prob / betas / get_betas / get_tsw / acceptance_fraction ...).

Everything of a sweep runs on the GPU (include/emperor_b200.h `emp_pt_sweep`): nsteps x
(proposal + prior, likelihood + Metropolis accept) per half-ensemble, the hot->cold swap
plan, the ladder adaptation (reddemcee adapt_tau / adapt_nu, adapt_mode 0), the tsw / smd /
beta histories and the chain store.  The host only supplies the random draws (draws.py), one
sweep ahead, through a double-buffered pinned staging area; there is NO host synchronisation
inside a run: with nsteps = 1 a sweep is six kernel launches, replayed from a CUDA graph.
Small ensembles (a sweep of tens of microseconds) are replayed k sweeps per graph launch
(`emp_pt_sweep_chunk`: the graph uploads the k sweeps' draws with one copy node).

With `torch.distributed` initialised (one process per GPU) the temperature ladder is sharded
over the ranks, T/G temperatures each, interleaved by default (dist.py): the stretch steps need
no communication; per sweep every rank WRITES its rows of logL[T, W] and of the swap draws it
generated for its own pairs into every peer's gathered block over NVLink (CUDA IPC, one small
kernel + release/acquire flags), every rank replays the same plan + adaptation, and the swap is
applied by reading the source rows straight from the owners' HBM: only the rows a rank receives
cross the links, and a sweep contains no NCCL call (7 launches, graph-replayed).
`exchange='allgather'` keeps the NCCL formulation (all-gathers of logL, the draws and the blocks).
"""
from __future__ import annotations

import ctypes
import time as _time
import warnings
from typing import Optional

import numpy as np

from . import _lib
from . import dist as _dist
from .draws import DrawStreams, SweepDraws, default_betas, draw_sweep, initial_positions, sweep_shapes


class State:
    """What `run_mcmc` returns (emcee's State as EMPEROR uses it: support/endit_freeze1.scr passes it
    back into the next `run_mcmc`).  Handing the same object back continues the run without re-evaluating
    anything.  `coords` / `log_like` / `log_prior` (the whole ladder, [T, W, ndim] / [T, W], on the host) are
    pulled from the device on first access — a run that only chains `run_mcmc` calls never pays the copy — and
    must be read before the sampler moves on."""

    def __init__(self, sampler):
        self._sampler, self._sampler_id, self._iteration = sampler, id(sampler), sampler.iteration
        self._host = None

    def _pull(self):
        if self._host is None:
            if self._sampler.iteration != self._iteration:
                raise RuntimeError("this State was not read before the sampler continued; the ensemble has moved on")
            self._host = self._sampler.state_numpy()
        return self._host

    coords = property(lambda self: self._pull()[0])
    log_like = property(lambda self: self._pull()[1])
    log_prior = property(lambda self: self._pull()[2])

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.coords, dtype=dtype)

    def __iter__(self):
        return iter(self._pull())


class PTSampler:
    def __init__(self, nwalkers: int, ndim: int, log_like, log_prior=None, ntemps: int = 1, pool=None,
                 backend=None, betas=None, tsw_history: bool = True, smd_history: bool = True,
                 adapt_tau: float = 1000, adapt_nu: float = 1, adapt_mode: int = 0, a: float = 2.0,
                 seed: Optional[int] = None, store: str = "device", thin_by: int = 1, adapt: bool = True, group=None,
                 layout: str = "strided", exchange: str = "peer", graph: bool = True, native_draws: bool = True,
                 chunk: Optional[int] = None):
        """`log_like` is the LikelihoodEngine (it carries the prior as well; `log_prior` and `pool` are
        accepted for signature compatibility and ignored — the walkers are evaluated on the GPU, not
        through a multiprocessing pool).  `backend`: None, or a file name: the run is written there in
        the reference's HDF5 layout when `run_mcmc` returns (postproc.save_backend)."""
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
        b0 = (np.array(betas, dtype=np.float64) if betas is not None else default_betas(ndim, ntemps))
        if len(b0) != self.ntemps:
            raise ValueError(f"betas should have {ntemps} items")
        self._betas_host = b0
        self._betas_initial = b0.copy()
        self._betas_stale = False
        # native_draws: the C generator of the C-ABI library (bit-identical to numpy.random.RandomState, threaded)
        self.streams = DrawStreams(seed, self.ntemps, native=native_draws)
        self.D_ = None
        if store not in ("device", "host", None):
            raise ValueError("store must be 'device', 'host' or None")
        self.store, self.thin_by = store, max(int(thin_by), 1)
        self.backend_file = backend if isinstance(backend, str) else None
        self.torch = torch
        self.dev = self.engine.torch_device
        self.shard = _dist.LadderShard(self.ntemps, group=group, layout=layout)
        if exchange not in ("peer", "allgather"):
            raise ValueError("exchange must be 'peer' (CUDA IPC over NVLink) or 'allgather' (NCCL)")
        self.exchange = exchange
        self.graph = bool(graph)
        # sweeps per graph launch in run_mcmc: None = as many as fit ~1 MB of draws (at most 64; 1 for the large
        # ensembles, whose sweep takes milliseconds), 1 = sweep by sweep
        self.chunk = None if chunk is None else max(int(chunk), 1)
        self.iteration = 0   # sweeps done
        self.time = 0        # reddemcee's ladder clock (= sweeps done)
        self._n_steps = 0    # stretch steps done
        self._stored = 0     # samples stored
        self._sample_sweep = []  # sweep index of every stored sample
        self._chain = self._ll = self._lp = None
        self._hist_cap = 0
        self._beta_hist = self._nacc_hist = self._smd_hist = None
        self._D_dev = None
        self._ring = None
        self._nan_seen = 0
        self.timings = {"draws": 0.0, "h2d": 0.0}

    # ---- the ladder ------------------------------------------------------------------------------
    @property
    def betas(self):
        """Current ladder (adapted on the device after every sweep; reading it synchronises)."""
        if self._betas_stale:
            self._betas_host = self._betas_dev.cpu().numpy().copy()
            self._betas_stale = False
        return self._betas_host

    @betas.setter
    def betas(self, value):
        b = np.array(value, dtype=np.float64)
        if len(b) != self.ntemps:
            raise ValueError(f"betas should have {self.ntemps} items")
        self._betas_host, self._betas_stale = b, False
        if hasattr(self, "_betas_dev"):
            self._betas_dev.copy_(self.torch.from_numpy(b))

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

    def _alloc_state(self):
        """Device state, allocated once.  Both copies of (p | logl | logp) live in one block each; when the
        ladder is sharded the blocks are cudaMalloc'ed through the C-ABI and mapped into every peer (CUDA IPC)
        so that the swap can read its source rows from the owners' HBM."""
        torch = self.torch
        Tl, W, nd = self.shard.n_local, self.nwalkers, self.ndim
        n_row = Tl * W
        nbytes = n_row * (nd + 2) * 8
        self._blocks, self._peer_blocks = [], [[], []]
        self._state = []
        peer = self.shard.world > 1 and self.exchange == "peer"
        for par in range(2):
            if peer:
                from .engine import SharedDeviceBuffer
                blk = SharedDeviceBuffer(self.engine.device, nbytes)
                self._blocks.append(blk)
                mk = lambda shape, off, blk=blk: blk.tensor(shape, off)
            else:
                raw = torch.empty(n_row * (nd + 2), dtype=torch.float64, device=self.dev)
                self._blocks.append(raw)
                mk = lambda shape, off, raw=raw: raw[off // 8: off // 8 + int(np.prod(shape))].view(shape)
            self._state.append((mk((Tl, W, nd), 0), mk((Tl, W), n_row * nd * 8), mk((Tl, W), n_row * (nd + 1) * 8)))
        self._par = 0
        self.p, self.logl, self.logp = self._state[0]
        self._gath, self._peer_gath = [], [[], []]
        if peer:
            # gathered blocks of the peer-push exchange (one per sweep parity): every rank writes its rows of logL
            # and of the swap draws straight into them over NVLink (pt_publish_kernel) — no NCCL call in a sweep
            nb = ctypes.c_int64()
            _lib.check(_lib.lib().emp_gather_block_bytes(self.ntemps, W, ctypes.byref(nb)))
            from .engine import SharedDeviceBuffer
            for q in range(2):
                g = SharedDeviceBuffer(self.engine.device, nb.value)
                g.tensor((nb.value,), 0, typestr="|u1").zero_()
                self._gath.append(g)
            torch.cuda.synchronize(self.dev)
            self._exchange_ipc_handles(nbytes, nb.value)
        self.accepted = torch.zeros((Tl, W), dtype=torch.uint8, device=self.dev)
        self._n_accepted = torch.zeros((Tl, W), dtype=torch.int32, device=self.dev)
        self._src = torch.empty((self.ntemps, W), dtype=torch.int32, device=self.dev)
        self._n_acc = torch.zeros((max(self.ntemps - 1, 1),), dtype=torch.int32, device=self.dev)
        self._betas_dev = self._upload(self._betas_host)
        self._counters = torch.zeros(2, dtype=torch.int64, device=self.dev)  # [0] sweeps, [1] stretch steps
        self._counters[0] = self.iteration
        self._counters[1] = self._n_steps
        self.shard.warm_up(self.dev)

    def _exchange_ipc_handles(self, nbytes, gbytes):
        """All ranks publish the CUDA IPC handles of their two state blocks and their two gathered blocks once;
        every rank maps its peers'."""
        from .engine import SharedDeviceBuffer
        td, sh = self.shard.td, self.shard
        mine = [b.export() for b in self._blocks] + [g.export() for g in self._gath]
        allh = [None] * sh.world
        td.all_gather_object(allh, mine, group=sh.group)
        for par in range(2):
            row, grow = [], []
            for r in range(sh.world):
                row.append(self._blocks[par] if r == sh.rank
                           else SharedDeviceBuffer.open(self.engine.device, allh[r][par], nbytes))
                grow.append(self._gath[par] if r == sh.rank
                            else SharedDeviceBuffer.open(self.engine.device, allh[r][2 + par], gbytes))
            self._peer_blocks[par] = row
            self._peer_gath[par] = grow

    def _init_state(self, p0):
        torch = self.torch
        p0 = np.asarray(p0, dtype=np.float64)
        if p0.shape != (self.ntemps, self.nwalkers, self.ndim):
            raise ValueError(f"p0 must have shape {(self.ntemps, self.nwalkers, self.ndim)}")
        if not hasattr(self, "_state"):
            self._alloc_state()
        self.p.copy_(self._upload(p0[self.shard.local_slice]))
        self.engine.logl_batch_device(self.p.view(-1, self.ndim), self.logl.view(-1), self.logp.view(-1))
        self._prefetched = None
        if self.shard.world > 1:
            torch.cuda.synchronize(self.dev)
            self.shard.td.barrier(group=self.shard.group)  # peers may read this state from the first swap on

    def _alloc_hist(self, nsweeps):
        """Room for `nsweeps` more rows of the per-sweep histories (device; capacity grows geometrically)."""
        torch = self.torch
        need = self.iteration + nsweeps
        if need <= self._hist_cap:
            return
        cap = max(need, 2 * self._hist_cap, 16)
        T, Tl = self.ntemps, self.shard.n_local
        new = [torch.zeros((cap, T), dtype=torch.float64, device=self.dev),
               torch.zeros((cap, max(T - 1, 1)), dtype=torch.int32, device=self.dev),
               torch.zeros((cap, Tl), dtype=torch.float64, device=self.dev)]
        if self._beta_hist is not None and self.iteration:
            for dst, srcb in zip(new, (self._beta_hist, self._nacc_hist, self._smd_hist)):
                dst[: self.iteration].copy_(srcb[: self.iteration])
        self._beta_hist, self._nacc_hist, self._smd_hist = new
        self._hist_cap = cap

    def _alloc_store(self, nsamples):
        """Make room for `nsamples` more stored samples (capacity grows geometrically so repeated
        run_mcmc calls do not re-allocate and copy the chain every time)."""
        torch = self.torch
        Tl, W, nd = self.shard.n_local, self.nwalkers, self.ndim
        if self.store == "device":
            dev, pin = self.dev, False
        elif self.store == "host":
            dev, pin = "cpu", True
        else:
            self._chain = None
            return
        need = self._stored + nsamples
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
            torch.cuda.synchronize(self.dev)
            for dst, srcb in zip(new, (self._chain, self._ll, self._lp)):
                dst[: self._stored].copy_(srcb[: self._stored])
        self._chain, self._ll, self._lp = new

    def _alloc_ring(self, nsteps, k=1):
        """store='host': the device writes its samples into a small ring that a copy stream drains into the
        pinned host chain one sweep (one chunk of k sweeps) behind the compute stream (C5 produces 147 MB per
        step: nothing of the chain accumulates in HBM)."""
        torch = self.torch
        m = (nsteps + self.thin_by - 1) // self.thin_by + 1
        slots = 2 * m * k
        if self._ring is not None and self._ring[0].shape[0] >= slots:
            return
        torch.cuda.synchronize(self.dev)
        Tl, W, nd = self.shard.n_local, self.nwalkers, self.ndim
        kw = dict(dtype=torch.float64, device=self.dev)
        self._ring = [torch.empty((slots, Tl, W, nd), **kw), torch.empty((slots, Tl, W), **kw),
                      torch.empty((slots, Tl, W), **kw)]
        self._copy_stream = torch.cuda.Stream(device=self.dev)
        self._copy_done = []   # (sweep index, event) of the ring drains still worth waiting for

    # ------------------------------------------------------------------------------
    def stage_draws(self, draws: SweepDraws, pinned: bool = False):
        """Copy one sweep's draws to the device (async on the current stream).  `draws` holds the
        stretch draws of THIS rank's temperatures and the swap draws of the whole ladder (or, sharded, of
        this rank's pairs).  pinned=True packs all seven arrays into one reusable pinned staging buffer and
        issues a single H2D copy (double-buffered: the previous sweep may still be reading its draws)."""
        torch = self.torch
        fields = [(f, getattr(draws, f)) for f in SweepDraws.FIELDS
                  if not (f in ("perm", "lnu_swap") and self.ntemps < 2)]
        out = {f: None for f in SweepDraws.FIELDS}
        out["sharded_swap"] = draws.sharded_swap
        if not pinned:
            for f, a in fields:
                out[f] = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, non_blocking=True)
            return out
        offs, total = [], 0
        for f, a in fields:
            offs.append(total)
            total += (a.nbytes + 255) // 256 * 256
        if getattr(self, "_stage_cap", 0) != total:
            torch.cuda.synchronize(self.dev)
            self._stage_host = [torch.empty(total, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._stage_dev = [torch.empty(total, dtype=torch.uint8, device=self.dev) for _ in range(2)]
            self._stage_evt = [torch.cuda.Event(), torch.cuda.Event()]   # H2D of buffer i finished
            self._stage_cap, self._stage_i = total, 0
        i = self._stage_i = 1 - self._stage_i
        # the H2D that last used this pinned buffer has finished (this also keeps the host at most two sweeps
        # ahead of the device: that copy is stream-ordered behind the sweep before it)
        self._stage_evt[i].synchronize()
        host, dev = self._stage_host[i], self._stage_dev[i]
        hv = host.numpy()
        for (f, a), o in zip(fields, offs):
            hv[o:o + a.nbytes] = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
        dev[:total].copy_(host[:total], non_blocking=True)
        self._stage_evt[i].record()
        for (f, a), o in zip(fields, offs):
            tdt = torch.int32 if a.dtype == np.int32 else torch.float64
            out[f] = dev[o:o + a.nbytes].view(tdt).view(a.shape)
        out["_stage_index"] = i
        return out

    def draw_staged(self, nsteps: int):
        """Draw the next sweep straight into the pinned staging buffer (no intermediate arrays, no packing) and
        enqueue its H2D copy; returns the dict of device views `sweep_begin` takes.  The NumPy views of the two
        pinned buffers and the torch views of their device copies are built once per (nsteps) layout."""
        torch, sh = self.torch, self.shard
        lay = getattr(self, "_lay", None)
        if lay is None or lay["nsteps"] != nsteps:
            n_rows = 0 if self.ntemps < 2 else (self.ntemps - 1 if sh.world == 1 else sh.n_local)
            shapes = sweep_shapes(sh.n_local, self.nwalkers, nsteps, max(n_rows, 1))
            offs, total = [], 0
            for f, shp, dt in shapes:
                offs.append(total)
                total += (int(np.prod(shp)) * np.dtype(dt).itemsize + 255) // 256 * 256
            torch.cuda.synchronize(self.dev)
            lay = {"nsteps": nsteps, "total": total, "host": [], "dev": [], "np": [], "out": [], "i": 0,
                   "evt": [torch.cuda.Event(), torch.cuda.Event()],      # H2D of slot i landed
                   "used": [None, None],                                 # the sweep that read slot i has finished
                   "stream": torch.cuda.Stream(device=self.dev)}         # uploads overlap the running sweep
            for i in range(2):
                host = torch.empty(total, dtype=torch.uint8).pin_memory()
                dev = torch.empty(total, dtype=torch.uint8, device=self.dev)
                hv = host.numpy()
                views, out = {}, {"sharded_swap": sh.world > 1, "_stage_index": i}
                for (f, shp, dt), o in zip(shapes, offs):
                    nb = int(np.prod(shp)) * np.dtype(dt).itemsize
                    views[f] = hv[o:o + nb].view(dt).reshape(shp)
                    out[f] = dev[o:o + nb].view(torch.int32 if dt == np.int32 else torch.float64).view(shp)
                lay["host"].append(host), lay["dev"].append(dev), lay["np"].append(views), lay["out"].append(out)
            self._lay = lay
        i = lay["i"] = 1 - lay["i"]
        # the H2D that last used this pinned buffer has finished (this also keeps the host at most two sweeps
        # ahead of the device: that copy is stream-ordered behind the sweep before it)
        t_w = _time.perf_counter()
        lay["evt"][i].synchronize()
        self.timings["wait_device"] = self.timings.get("wait_device", 0.0) + _time.perf_counter() - t_w
        t0 = _time.perf_counter()
        rows = None if sh.world == 1 else range(self.ntemps)[sh.local_slice]
        draw_sweep(self.streams, self.nwalkers, self.ndim, nsteps, self.a, temps=sh.local_slice,
                   swap=self.ntemps > 1, swap_rows=rows, out=lay["np"][i])
        self.timings["draws"] += _time.perf_counter() - t0
        # upload on a side stream, behind the sweep that last read this slot: the copy overlaps the sweep in flight
        st = lay["stream"]
        if lay["used"][i] is not None:
            st.wait_event(lay["used"][i])
        with torch.cuda.stream(st):
            lay["dev"][i].copy_(lay["host"][i], non_blocking=True)
            lay["evt"][i].record(st)
        return lay["out"][i]

    def draw_resident(self, nsteps: int):
        """Draw one sweep and keep it packed in device memory (bench.py: the draws of the timed steps are
        resident in HBM before the clock starts); feed it back with `stage_resident`."""
        out = self.draw_staged(nsteps)
        i = out["_stage_index"]
        self.torch.cuda.current_stream(self.dev).wait_event(self._lay["evt"][i])
        return self._lay["dev"][i].clone()

    def stage_resident(self, packed):
        """Device-to-device copy of a `draw_resident` buffer into the next staging slot (the slots are what the
        captured graphs point at); returns the dict of device views `sweep_begin` takes."""
        lay = self._lay
        i = lay["i"] = 1 - lay["i"]
        lay["dev"][i].copy_(packed, non_blocking=True)
        lay["evt"][i].record()
        return lay["out"][i]

    def _sweep_args(self, draws, nsteps, par=None):
        """The EmpPtSweep argument block of one sweep (include/emperor_b200.h).  The blocks of the steady state
        (double-buffered state x double-buffered staging) are built once and reused.  `par`: which of the two
        state blocks the sweep starts from (default: the current one)."""
        par = self._par if par is None else par
        key = (par, draws["zz"].data_ptr() if draws.get("_stage_index") is not None else None, nsteps,
               self.adapt, self._hist_cap,
               None if self._chain is None else self._chain.data_ptr(),
               None if self._ring is None else self._ring[0].data_ptr(), self.D_ is not None)
        cache = getattr(self, "_args_cache", None)
        if cache is None:
            cache = self._args_cache = {}
        if key[1] is not None and key in cache:
            return cache[key]
        sh = self.shard
        A = _lib.EmpPtSweepC()
        A.T_loc, A.W, A.nsteps, A.T_all = sh.n_local, self.nwalkers, nsteps, self.ntemps
        A.n_ranks, A.rank, A.strided = sh.world, sh.rank, 1 if sh.layout == "strided" else 0
        # graphs are keyed on the argument block: only the double-buffered pinned staging repeats its pointers
        push = sh.world > 1 and self.exchange == "peer"
        A.use_graph = 1 if (self.graph and (sh.world == 1 or push) and draws.get("_stage_index") is not None) else 0
        cur, alt = self._state[par], self._state[1 - par]
        A.p, A.logl, A.logp = cur[0].data_ptr(), cur[1].data_ptr(), cur[2].data_ptr()
        A.p_alt, A.logl_alt, A.logp_alt = alt[0].data_ptr(), alt[1].data_ptr(), alt[2].data_ptr()
        A.betas = self._betas_dev.data_ptr()
        for f in ("half_idx", "zz", "rint", "factors", "lnu"):
            setattr(A, f, draws[f].data_ptr())
        A.accepted, A.n_accepted = self.accepted.data_ptr(), self._n_accepted.data_ptr()
        A.src, A.n_acc = self._src.data_ptr(), self._n_acc.data_ptr()
        A.adapt = 1 if (self.adapt and self.ntemps > 2) else 0
        A.thin = self.thin_by
        A.adapt_tau, A.adapt_nu = float(self.adapt_tau), float(self.adapt_nu)
        A.sweep_counter = self._counters[0:].data_ptr()
        A.step_counter = self._counters[1:].data_ptr()
        A.beta_hist = self._beta_hist.data_ptr()
        A.nacc_hist = self._nacc_hist.data_ptr() if self.ntemps > 1 else None
        A.hist_cap = self._hist_cap
        if self.smd_history_bool and self.D_ is not None and self.ntemps > 1:
            if self._D_dev is None or self._D_dev.shape[0] != self.ndim:
                self._D_dev = self._upload(np.asarray(self.D_, dtype=np.float64))
            A.D, A.smd_hist = self._D_dev.data_ptr(), self._smd_hist.data_ptr()
        if self._chain is not None:
            tgt = self._ring if self.store == "host" else (self._chain, self._ll, self._lp)
            A.chain, A.chain_ll, A.chain_lp = tgt[0].data_ptr(), tgt[1].data_ptr(), tgt[2].data_ptr()
            A.store_cap = tgt[0].shape[0]
            A.store_ring = 1 if self.store == "host" else 0
        A.perm_hot_sorted = 1  # draws.py lists every pair by its slot in the warmer row
        if (sh.world == 1 or push) and self.ntemps > 1:
            # the whole ladder's pairs, or (sharded, peer-push exchange) the pairs this rank drew
            A.perm, A.lnu_swap = draws["perm"].data_ptr(), draws["lnu_swap"].data_ptr()
        if push:
            if not draws.get("sharded_swap"):
                raise ValueError("the peer-push exchange takes the swap draws sharded (this rank's pair rows)")
            self._peer_pointers(A, par)
            for q in range(2):
                for r in range(sh.world):
                    A.peer_gath[q][r] = self._peer_gath[q][r].ptr
        if key[1] is not None:
            if len(cache) > 16:
                cache.clear()
            cache[key] = A
        return A

    def sweep_begin(self, draws):
        """Enqueue one whole sweep: nsteps stretch steps of every local temperature, the swap sweep, the ladder
        adaptation, histories and chain store (asynchronous; nothing here waits for the device).
        `draws`: SweepDraws (host) or the dict `stage_draws` returned (already on the device)."""
        torch, eng, sh = self.torch, self.engine, self.shard
        if isinstance(draws, SweepDraws):
            t0 = _time.perf_counter()
            draws = self.stage_draws(draws)
            self.timings["h2d"] += _time.perf_counter() - t0
        nsteps = int(draws["zz"].shape[0])
        self._alloc_hist(1)
        # storage bookkeeping (the device counts the same way: sample n is stored when n % thin == 0)
        n0 = self._n_steps
        stored_now = [n for n in range(n0, n0 + nsteps) if n % self.thin_by == 0]
        if self._chain is not None:
            if self.store == "host":
                self._alloc_ring(nsteps)
                # the ring slots this sweep overwrites were drained two sweeps ago at the latest
                while self._copy_done and self._copy_done[0][0] <= self.iteration - 2:
                    torch.cuda.current_stream(self.dev).wait_event(self._copy_done.pop(0)[1])
                while getattr(self, "_chunk_drains", None):   # drains of a chunked run still in flight
                    torch.cuda.current_stream(self.dev).wait_event(self._chunk_drains.pop(0))
            if self._stored + len(stored_now) > self._chain.shape[0]:
                self._alloc_store(max(len(stored_now), 1))
        slot = draws.get("_stage_index") if getattr(self, "_lay", None) is not None and \
            draws is self._lay["out"][draws.get("_stage_index") or 0] else None
        if slot is not None:
            torch.cuda.current_stream(self.dev).wait_event(self._lay["evt"][slot])
        self._mark("start")
        if self.ntemps > 1:
            perm, lnu_swap = draws["perm"], draws["lnu_swap"]
        A = self._sweep_args(draws, nsteps)
        if sh.world == 1:
            eng.pt_sweep(A)
            self._mark("sweep")
        elif self.exchange == "peer":
            eng.pt_sweep(A)   # stretch, publish to the peers' gathered blocks, plan, peer-read application
            self._mark("sweep")
        else:
            eng.pt_sweep_stretch(A)
            self._mark("stretch")
            if draws.get("sharded_swap"):
                # every rank replays the whole plan: gather the pair rows (NCCL, behind the stretch kernels: on a
                # side stream under them the NCCL kernels spin on SMs the likelihood kernel needs — measured
                # +0.35 ms of stretch phase for 0.09 ms of hidden gathers at 4 GPUs)
                perm = sh.all_gather_rows(perm)[: self.ntemps - 1].contiguous()
                lnu_swap = sh.all_gather_rows(lnu_swap)[: self.ntemps - 1].contiguous()
            logl_all = sh.all_gather_rows(self.logl).contiguous()  # [T, W] in ladder order
            self._mark("allgather")
            A.perm, A.lnu_swap, A.logl_all = perm.data_ptr(), lnu_swap.data_ptr(), logl_all.data_ptr()
            self._peer_pointers(A)
            eng.pt_sweep_swap(A)
            self._keep = (perm, lnu_swap, logl_all)
            self._mark("swap")
        if self.ntemps > 1:
            self._par = 1 - self._par
            self.p, self.logl, self.logp = self._state[self._par]
        if slot is not None:
            if self._lay["used"][slot] is None:
                self._lay["used"][slot] = torch.cuda.Event()
            self._lay["used"][slot].record()
        # host mirror of the device counters
        self._n_steps += nsteps
        first = self._stored
        self._stored += len(stored_now) if self._chain is not None else 0
        self._sample_sweep += [self.iteration] * (len(stored_now) if self._chain is not None else 0)
        if self._chain is not None and self.store == "host" and stored_now:
            ev = torch.cuda.Event()
            ev.record()
            slots = self._ring[0].shape[0]
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ev)
                for k, n in enumerate(stored_now):
                    s = (n // self.thin_by) % slots
                    for dst, ring in zip((self._chain, self._ll, self._lp), self._ring):
                        dst[first + k].copy_(ring[s], non_blocking=True)
                done = torch.cuda.Event()
                done.record()
            self._copy_done.append((self.iteration, done))
        self.time += 1
        self.iteration += 1
        self._betas_stale = self._betas_stale or bool(A.adapt)

    # ---- chunked runs: k sweeps per graph launch (ensembles whose sweep takes tens of microseconds) -----------
    def _chunk_len(self, nsteps):
        """Sweeps per graph launch of `run_mcmc`.  A sweep of a small ensemble (BASELINE configs 1-3) is 0.1-0.2 ms
        of device time: enqueueing sweep by sweep from Python (draw, stage, upload, launch: ~0.1 ms) would make
        the host the bottleneck.  Large ensembles (config 4: 3 MB of draws and 11 ms per sweep) stay at 1."""
        if not self.graph or self.shard.world > 1 or self.engine.timing_enabled:
            return 1
        if self.chunk is not None:
            return self.chunk
        n_rows = max(self.ntemps - 1, 1)
        per = sum(int(np.prod(shp)) * np.dtype(dt).itemsize
                  for _, shp, dt in sweep_shapes(self.shard.n_local, self.nwalkers, nsteps, n_rows))
        k = int(min(64, (1 << 20) // max(per, 1)))
        return max(k - (k & 1), 1)  # even: the state parity at the start of a chunk does not alternate

    def _chunk_layout(self, k, nsteps):
        """Pinned and device blocks for the draws of k sweeps (double-buffered), the NumPy views the generator
        and the vectorised log pass write through, the argument tuples of the native generator per sweep."""
        torch, sh = self.torch, self.shard
        lay = getattr(self, "_clay", None)
        if lay is not None and lay["k"] == k and lay["nsteps"] == nsteps:
            return lay
        T, W = self.ntemps, self.nwalkers
        n_rows = max(T - 1, 1)
        shapes = sweep_shapes(sh.n_local, W, nsteps, n_rows)
        offs, total = [], 0
        for f, shp, dt in shapes:
            offs.append(total)
            total += (int(np.prod(shp)) * np.dtype(dt).itemsize + 255) // 256 * 256
        torch.cuda.synchronize(self.dev)
        lay = {"k": k, "nsteps": nsteps, "total": total, "host": [], "dev": [], "np": [], "out": [], "gen": [],
               "used": [None, None], "i": 0, "blocks": {}}
        dh = self.streams.handle() if self.streams.native else None
        ts = np.arange(T, dtype=np.int32)
        rs = np.array([T + j for j in range(T - 1)], dtype=np.int32)
        lay["keep"] = (ts, rs)
        for i in range(2):
            host = torch.empty(k * total, dtype=torch.uint8).pin_memory()
            dev = torch.empty(k * total, dtype=torch.uint8, device=self.dev)
            hv = host.numpy()
            views, outs, gens = {}, [], []
            for (f, shp, dt), o in zip(shapes, offs):  # [k, ...] views, one row per sweep of the chunk
                st = np.empty(shp, dtype=dt).strides
                views[f] = np.ndarray((k,) + tuple(shp), dtype=dt, buffer=hv, offset=o, strides=(total,) + st)
            for j in range(k):
                out, ptr = {"sharded_swap": False, "_stage_index": None}, {}
                for (f, shp, dt), o in zip(shapes, offs):
                    nb = int(np.prod(shp)) * np.dtype(dt).itemsize
                    a, b = j * total + o, j * total + o + nb
                    out[f] = dev[a:b].view(torch.int32 if dt == np.int32 else torch.float64).view(shp)
                    ptr[f] = hv[a:b].ctypes.data
                outs.append(out)
                gens.append((dh, ts.ctypes.data, T, W, nsteps, ptr["half_idx"], ptr["zz"], ptr["rint"], ptr["lnu"],
                             rs.ctypes.data, len(rs), ptr["perm"], ptr["lnu_swap"]))
            lay["host"].append(host), lay["dev"].append(dev), lay["np"].append(views)
            lay["out"].append(outs), lay["gen"].append(gens)
            if T < 2:
                views["perm"][...] = 0
                views["lnu_swap"][...] = 1.0
        self._clay = lay
        return lay

    def _chunk_draw(self, lay, i, n):
        """Draw the next n sweeps into pinned block i: ONE call of the native generator (every stream draws its n
        sweeps in order, so the streams advance exactly as in the sweep-by-sweep path), then ONE vectorised pass for
        the stretch factors and the logs (NumPy's on every path: thresholds are the same bits for the device and
        the oracle)."""
        t0 = _time.perf_counter()
        v = lay["np"][i]
        if self.streams.native:
            g0 = lay["gen"][i][0]   # sweep q of the chunk sits q * total bytes behind sweep 0: one call draws them all
            _lib.check(_lib.lib().emp_draws_sweeps(g0[0], n, lay["total"], *g0[1:]))
        else:
            for j in range(n):
                draw_sweep(self.streams, self.nwalkers, self.ndim, lay["nsteps"], self.a, swap=self.ntemps > 1,
                           out={f: v[f][j] for f in SweepDraws.FIELDS})
            self.timings["draws"] += _time.perf_counter() - t0
            return
        zz, fac, lnu, lsw = v["zz"][:n], v["factors"][:n], v["lnu"][:n], v["lnu_swap"][:n]
        zz[...] = ((self.a - 1.0) * zz + 1) ** 2.0 / self.a  # draws._zz_from_u
        np.multiply(np.log(zz), self.ndim - 1.0, out=fac)
        with np.errstate(divide="ignore"):
            np.log(lnu, out=lnu)
            if self.ntemps > 1:
                np.log(lsw, out=lsw)
        self.timings["draws"] += _time.perf_counter() - t0

    def _chunk_blocks(self, lay, i, n):
        """The n argument blocks of a chunk read from device block i, starting at the current state parity."""
        key = (i, n, self._par, self.adapt, self._hist_cap, None if self._chain is None else self._chain.data_ptr(),
               None if self._ring is None else self._ring[0].data_ptr(), self.D_ is not None)
        arr = lay["blocks"].get(key)
        if arr is None:
            if len(lay["blocks"]) > 8:
                lay["blocks"].clear()
            arr = (_lib.EmpPtSweepC * n)()
            flip = 1 if self.ntemps > 1 else 0
            for j in range(n):
                arr[j] = self._sweep_args(lay["out"][i][j], lay["nsteps"], par=(self._par + j * flip) & 1)
            lay["blocks"][key] = arr
        return arr

    def _chunk_launch(self, lay, i, n, upload=True):
        """Enqueue the n sweeps whose draws are in block i (ONE graph launch; with `upload` the graph first copies
        the pinned block to the device), mirror the device counters on the host and, with store='host', queue the
        drain of the samples the chunk stores."""
        torch, eng = self.torch, self.engine
        nsteps = lay["nsteps"]
        main = torch.cuda.current_stream(self.dev)
        host_store = self._chain is not None and self.store == "host"
        if host_store:
            self._alloc_ring(nsteps, lay["k"])
            self._chunk_drains = getattr(self, "_chunk_drains", [])
            while len(self._chunk_drains) > 1:   # the ring holds two chunks: chunk c-2 must have been drained
                main.wait_event(self._chunk_drains.pop(0))
            while self._copy_done:               # drains of the sweep-by-sweep path still in flight
                main.wait_event(self._copy_done.pop(0)[1])
        arr = self._chunk_blocks(lay, i, n)
        if upload:
            _lib.check(eng._L.emp_pt_sweep_chunk(eng._h, arr, n, lay["dev"][i].data_ptr(), lay["host"][i].data_ptr(),
                                                 n * lay["total"]))
        else:
            _lib.check(eng._L.emp_pt_sweep_chunk(eng._h, arr, n, None, None, 0))
        if lay["used"][i] is None:
            lay["used"][i] = torch.cuda.Event()
        lay["used"][i].record()
        n0, first = self._n_steps, self._stored
        if self._chain is not None:
            stored = [n_ for n_ in range(n0, n0 + n * nsteps) if n_ % self.thin_by == 0]
            self._sample_sweep += [self.iteration + (n_ - n0) // nsteps for n_ in stored]
            self._stored += len(stored)
            if host_store and stored:
                slots = self._ring[0].shape[0]
                ev = torch.cuda.Event()
                ev.record()
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(ev)
                    s0, cnt, dst0 = (stored[0] // self.thin_by) % slots, len(stored), first
                    while cnt > 0:   # at most two contiguous runs (the ring wraps once)
                        run = min(cnt, slots - s0)
                        for dst, ring in zip((self._chain, self._ll, self._lp), self._ring):
                            dst[dst0:dst0 + run].copy_(ring[s0:s0 + run], non_blocking=True)
                        s0, cnt, dst0 = (s0 + run) % slots, cnt - run, dst0 + run
                    drained = torch.cuda.Event()
                    drained.record()
                self._chunk_drains.append(drained)
        self._n_steps += n * nsteps
        self.time += n
        self.iteration += n
        if self.ntemps > 1 and (n & 1):
            self._par = 1 - self._par
        self.p, self.logl, self.logp = self._state[self._par]
        self._betas_stale = self._betas_stale or bool(self.adapt and self.ntemps > 2)

    def _run_chunks(self, nsweeps, nsteps, k, bar=None):
        """`nsweeps` sweeps, k per graph launch.  While the device runs a chunk the host draws the next one into the
        other pinned block; the graph itself uploads its block (one copy node) before its first sweep."""
        lay = self._chunk_layout(k, nsteps)
        done = 0
        i = lay["i"]
        n_next = min(k, nsweeps)
        if lay["used"][i] is not None:
            lay["used"][i].synchronize()
        self._chunk_draw(lay, i, n_next)
        while done < nsweeps:
            n = n_next
            t0 = _time.perf_counter()
            self._chunk_launch(lay, i, n)
            done += n
            self.timings["enqueue"] = self.timings.get("enqueue", 0.0) + _time.perf_counter() - t0
            if bar is not None:
                bar.update(n)
            # the next chunk's draws, while the device works
            i = lay["i"] = 1 - i
            n_next = min(k, nsweeps - done)
            if n_next > 0:
                if lay["used"][i] is not None:
                    t_w = _time.perf_counter()
                    lay["used"][i].synchronize()
                    self.timings["wait_device"] = self.timings.get("wait_device", 0.0) + _time.perf_counter() - t_w
                self._chunk_draw(lay, i, n_next)

    def draw_chunk_resident(self, k, nsteps=1):
        """Draw k sweeps and keep them packed in device memory (bench.py: the draws of the timed steps are resident
        in HBM before the clock starts); feed the block back with `run_chunk_resident`."""
        lay = self._chunk_layout(k, nsteps)
        i = lay["i"]
        if lay["used"][i] is not None:
            lay["used"][i].synchronize()
        self._chunk_draw(lay, i, k)
        return lay["host"][i].to(self.dev, non_blocking=False)

    def run_chunk_resident(self, packed):
        """Device-to-device copy of a `draw_chunk_resident` block into the next device block, then its k sweeps
        from one graph launch (no host-to-device traffic)."""
        lay = self._clay
        i = lay["i"] = 1 - lay["i"]
        lay["dev"][i].copy_(packed, non_blocking=True)
        self._chunk_launch(lay, i, lay["k"], upload=False)

    def _peer_pointers(self, A, par=None):
        """Where the swap finds the CURRENT (p | logl | logp) block of every rank: peer HBM mapped with CUDA IPC
        (exchange='peer': only the rows a rank receives cross NVLink), or, with exchange='allgather' (the fallback
        for platforms without CUDA IPC), one NCCL all-gather of the blocks into a scratch buffer, read by the same
        kernel through the same pointer table."""
        sh = self.shard
        par = self._par if par is None else par
        nd, n_row = self.ndim, sh.n_local * self.nwalkers
        if self.exchange == "peer":
            bases = [self._peer_blocks[par][r].ptr for r in range(sh.world)]
        else:
            blk = self._blocks[par]
            self._gathered = sh.all_gather_flat(blk)[0]
            bases = [self._gathered.data_ptr() + r * blk.numel() * 8 for r in range(sh.world)]
        for r, base in enumerate(bases):
            A.peer_p[r], A.peer_logl[r], A.peer_logp[r] = base, base + n_row * nd * 8, base + n_row * (nd + 1) * 8

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
        """Swap counts of the sweep begun last (synchronises: tests and diagnostics only — run_mcmc never
        calls this; the histories stay on the device until somebody asks for them)."""
        if self.ntemps < 2:
            return None
        self.torch.cuda.current_stream(self.dev).synchronize()
        return self._n_acc.cpu().numpy()[: self.ntemps - 1].copy()

    def sweep(self, draws):
        """One full sweep: stretch steps + swap sweep + ladder adaptation; returns the swap counts."""
        self.sweep_begin(draws)
        return self.sweep_end()

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
        they are staged through a reusable pinned buffer.  Nothing synchronises with the device until
        the run is over.  Every stretch step is a stored sample (thin_by permitting): the chain of a run
        holds nsweeps*nsteps samples like reddemcee's, the tsw / smd / beta histories one row per sweep.
        `on_sweep(sampler, k)` is called after every sweep was enqueued (bench.py reads logL back there).
        Returns a `State`; pass it back (or None) to continue, or new positions to restart from them."""
        if isinstance(p0, State) and p0._sampler_id == id(self) and p0._iteration == self.iteration:
            p0 = None
        if p0 is not None:
            self._init_state(np.asarray(p0.coords if isinstance(p0, State) else p0))
        elif not hasattr(self, "_state"):
            raise ValueError("first call needs initial positions")
        n_new = sum(1 for n in range(self._n_steps, self._n_steps + nsweeps * nsteps) if n % self.thin_by == 0)
        self._alloc_store(n_new)
        self._alloc_hist(nsweeps)
        it = range(nsweeps)
        if progress:
            try:
                from tqdm import tqdm
                it = tqdm(it, total=nsweeps)
            except Exception:
                pass

        # the draws of the first sweep may already be staged: every call ends by drawing and staging one sweep
        # ahead (below), so that back-to-back calls — adaptation then production (support/endit_freeze1.scr),
        # warm-up then measurement — do not pay the pipeline fill again
        staged, pre = None, getattr(self, "_prefetched", None)
        self._prefetched = None
        k = self._chunk_len(nsteps) if on_sweep is None else 1
        if k > 1 and nsweeps > 1:
            # small ensembles: k sweeps per graph launch (a sweep staged by an earlier call goes first)
            if pre is not None and pre[0] == nsteps:
                self.sweep_begin(pre[1])
                nsweeps -= 1
            bar = it if hasattr(it, "update") else None
            self._run_chunks(nsweeps, nsteps, k, bar)
            if bar is not None:
                bar.close()
            nsweeps, it = 0, range(0)
        if nsweeps > 0:
            staged = pre[1] if (pre is not None and pre[0] == nsteps) else self.draw_staged(nsteps)
        for k in it:
            t0 = _time.perf_counter()
            self.sweep_begin(staged)
            self.timings["enqueue"] = self.timings.get("enqueue", 0.0) + _time.perf_counter() - t0
            # while the device runs this sweep the host draws the next one, packs it into the other pinned
            # buffer and enqueues its H2D copy (double-buffered on both sides)
            staged = self.draw_staged(nsteps)
            if on_sweep is not None:
                on_sweep(self, k)
        if nsweeps > 0:
            self._prefetched = (nsteps, staged)
        self.torch.cuda.synchronize(self.dev)
        nan = self.engine.nan_count()
        if nan > self._nan_seen:
            warnings.warn(f"{nan - self._nan_seen} proposal(s) had a NaN log-likelihood and were rejected "
                          "(emcee raises 'Probability function returned NaN' here)", RuntimeWarning)
            self._nan_seen = nan
        if self.backend_file:
            self.save_backend(self.backend_file)
        return State(self)

    def select_adjustment(self, mode):
        """reddemcee's ladder-adjustment selector as EMPEROR drives it: `support/endit_freeze1.scr:10` calls
        `sampler.select_adjustment('00')` after the adaptation burn-in to FREEZE the ladder for the production
        sweeps.  Other selectors are reddemcee internals that are not recoverable offline."""
        if str(mode) != "00":
            raise NotImplementedError(f"select_adjustment('{mode}'): only '00' (freeze the ladder) is implemented")
        self.adapt = False

    # ---- read-back API the reference's parent process uses --------
    def _sync_store(self):
        if self._ring is not None:
            self._copy_stream.synchronize()
        self.torch.cuda.synchronize(self.dev)

    def _get(self, buf, discard, thin, flat):
        if buf is None:
            raise RuntimeError("chain storage is disabled (store=None)")
        self._sync_store()
        x = buf[: self._stored][discard::thin]
        if self.shard.world > 1:
            x = self.shard.gather_to_all(x.to(self.dev), dim=1)  # NCCL gathers device tensors only
        x = x.cpu().numpy()
        x = np.swapaxes(x, 0, 1)  # [T, n, W, ...]
        if flat:
            x = x.reshape((x.shape[0], x.shape[1] * x.shape[2]) + x.shape[3:])
        return x

    def get_chain(self, discard=0, thin=1, flat=False):
        return self._get(self._chain, discard, thin, flat)

    def get_log_like(self, discard=0, thin=1, flat=False):
        return self._get(self._ll, discard, thin, flat)

    def get_betas_sweeps(self):
        """[n_sweeps, T] ladder after the adaptation of every sweep (device history)."""
        if self._beta_hist is None:
            return np.zeros((0, self.ntemps))
        return self._beta_hist[: self.iteration].cpu().numpy()

    def _betas_used(self):
        """[n_sweeps, T] ladder in force DURING each sweep (the initial one, then the adapted ones)."""
        bh = self.get_betas_sweeps()
        return np.concatenate([self._betas_initial[None, :], bh[:-1]], 0) if len(bh) else bh

    def get_log_prob(self, discard=0, thin=1, flat=False):
        """Tempered posterior beta*logL + logP per stored sample (beta: the ladder the sample was drawn under)."""
        ll = self._get(self._ll, discard, thin, False)
        lp = self._get(self._lp, discard, thin, False)
        bu = self._betas_used()[np.asarray(self._sample_sweep, dtype=np.int64)][discard::thin]  # [n, T]
        out = bu.T[:, :, None] * ll + lp
        return out.reshape(out.shape[0], -1) if flat else out

    def get_log_prior(self, discard=0, thin=1, flat=False):
        return self._get(self._lp, discard, thin, flat)

    def get_betas(self, discard=0):
        """[n_samples, T]: per stored sample (EMPEROR discards in steps, emp.py:962), the ladder after the
        adaptation of the sample's sweep — the ptemcee / reddemcee convention: the last row is `sampler.betas`."""
        bh = self.get_betas_sweeps()
        if not len(bh):
            return np.zeros((0, self.ntemps))
        return bh[np.asarray(self._sample_sweep, dtype=np.int64)][discard:]

    def get_tsw(self, discard=0):
        """[n_sweeps, T-1] swap acceptance ratio of every adjacent pair, per sweep."""
        if self._nacc_hist is None or self.ntemps < 2 or not self.tsw_history_bool:
            return np.zeros((0, max(self.ntemps - 1, 0)))
        return (self._nacc_hist[: self.iteration].cpu().numpy() / self.nwalkers)[discard:]

    def get_smd(self, discard=0):
        """[n_sweeps, T-1] swap mean distances (rung j <-> j+1); needs `sampler.D_`."""
        if self._smd_hist is None or self.D_ is None or not self.smd_history_bool or self.ntemps < 2:
            return np.zeros((0, max(self.ntemps - 1, 0)))
        x = self._smd_hist[: self.iteration]                                       # [n, T_loc]
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
            p, ll, lp = (self.shard.gather_to_all(x.contiguous(), dim=0) for x in (p, ll, lp))
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
        """reddemcee's 'hybrid' estimator is NOT recoverable offline (the package is not vendored): this raises
        so that EMPEROR's own fallback chain takes over — emp.py:1432-1447 wraps the call in try/except and
        falls back to `get_evidence_ti(..., pchip=False)`.  Ask for 'ss' or 'ti' explicitly instead."""
        raise NotImplementedError("reddemcee's hybrid evidence estimator is not available; use "
                                  "evidence_method 'ss' or 'ti' (EMPEROR falls back to TI by itself)")

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
        return self._v.nsamples

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
        self._betas = s.get_betas()
        self._accepted = np.rint(s.acceptance_fraction * max(s._n_steps, 1)).astype(np.int64)
        self.iteration = s.iteration            # sweeps: rows of tsw_history / smd_history (emp.py:733-745)
        self.nsamples = self._chain.shape[1]    # stored samples = stretch steps: backend[t].iteration (emp.py:748)
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
    """Host restatement of the device ladder adaptation (emp_pt.cuh::plan_tail): Vousden, Farr & Mandel (2016)
    dynamics in reddemcee's (adapt_tau, adapt_nu) parameterisation.  Kept for StoredRun / diagnostics; the
    sampler itself adapts on the device."""
    betas = np.array(betas, dtype=np.float64)
    decay = adapt_tau / (time + adapt_tau)
    kappa = decay / adapt_nu
    dSs = kappa * (ratios[:-1] - ratios[1:])
    deltaTs = np.diff(1 / betas[:-1])
    deltaTs = deltaTs * np.exp(dSs)
    betas[1:-1] = 1 / (np.cumsum(deltaTs) + 1 / betas[0])
    return betas
