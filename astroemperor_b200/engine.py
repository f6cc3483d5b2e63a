"""Device likelihood engine: the drop-in for the generated `my_likelihood` /
`my_prior` / `my_model` callables (support/likelihoods/00.like:3-5,
emp.py:190-254, emp_model.py:706-781) and the object the PT sampler drives.

`LikelihoodEngine` owns one `EmpHandle` (include/emperor_b200.h).  torch is used
for device memory and stream plumbing only: tensors are passed to the C-ABI as
raw device pointers and all kernels run on torch's current stream.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .modelspec import CompiledModel, ModelSpec


def _np64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class LikelihoodEngine:
    def __init__(self, model, t, y, yerr, flag, am=None, device: int = 0, sai=None):
        import torch  # device memory / streams only

        self.cm: CompiledModel = model.compile() if isinstance(model, ModelSpec) else model
        self.ndim = self.cm.ndim_free
        self.t, self.y, self.yerr = _np64(t), _np64(y), _np64(yerr)
        self.flag = np.ascontiguousarray(flag, dtype=np.int32)
        self.ndat = len(self.t)
        if not (len(self.y) == len(self.yerr) == len(self.flag) == self.ndat):
            raise ValueError("t, y, yerr, flag must have the same length")
        if not torch.cuda.is_available():
            raise _lib.EmperorB200Error(-3, "no CUDA device visible (there is no CPU fallback)")
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        L = _lib.lib()
        self._desc = self.cm.to_c()
        self._am_keepalive = None
        am_ptr = None
        if self.cm.am_enabled:
            if am is None:
                raise ValueError("model has an astrometric block but no `am` data was given")
            from .amdata import am_to_c
            am_c, self._am_keepalive = am_to_c(am)
            am_ptr = ctypes.cast(ctypes.pointer(am_c), ctypes.c_void_p)
        h = ctypes.c_void_p()
        _lib.check(L.emp_create(ctypes.cast(ctypes.pointer(self._desc), ctypes.c_void_p),
                                self.t.ctypes.data, self.y.ctypes.data, self.yerr.ctypes.data,
                                self.flag.ctypes.data, self.ndat, am_ptr, self.device, ctypes.byref(h)))
        self._h = h
        self._L = L
        n_sai = int(sum(getattr(self.cm, "sai_count", [])))
        if n_sai:
            # the SAI{j}_ columns of the generated script (emp_model.py:425-433), [ndat, n_sai]
            if sai is None:
                raise ValueError("model has a StellarActivityBlock but no `sai` columns were given")
            sai = _np64(sai).reshape(self.ndat, -1)
            if sai.shape[1] != n_sai:
                raise ValueError(f"sai has {sai.shape[1]} columns, the model expects {n_sai}")
            cols = np.ascontiguousarray(sai.T)  # column-major for the C-ABI: [n_sai][ndat]
            _lib.check(L.emp_attach_sai(self._h, cols.ctypes.data, n_sai))
        self.use_torch_stream()

    # ---- lifetime -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.emp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_torch_stream(self):
        """Run the handle's kernels on torch's current stream of this device."""
        import torch
        s = torch.cuda.current_stream(self.torch_device).cuda_stream
        _lib.check(self._L.emp_set_stream(self._h, ctypes.c_void_p(s)))

    def synchronize(self):
        _lib.check(self._L.emp_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        c = ctypes.c_int64()
        _lib.check(self._L.emp_launch_count(self._h, ctypes.byref(c)))
        return int(c.value)

    @property
    def graph_captures(self) -> int:
        """CUDA graphs captured by emp_pt_sweep so far (a steady run replays two)."""
        c = ctypes.c_int64()
        _lib.check(self._L.emp_graph_captures(self._h, ctypes.byref(c)))
        return int(c.value)

    SOLVERS = {"grid": 0, "kepler.py": 1}

    def set_solver(self, name: str):
        """'grid' (default): the grid-anchored refinement; 'kepler.py': the single high-order refinement of
        kepler.py 0.0.7 for every planet (include/emperor_b200.h emp_set_solver)."""
        _lib.check(self._L.emp_set_solver(self._h, self.SOLVERS[name]))

    timing_enabled = False

    def set_timing(self, on: bool):
        _lib.check(self._L.emp_set_timing(self._h, 1 if on else 0))
        self.timing_enabled = bool(on)

    def timing_collect(self):
        """(summed likelihood-kernel ms, launches) since the last collect; synchronises."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(self._L.emp_timing_collect(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def counters(self):
        """dict(proposals, in_prior, accepted, nan) of the PT steps run on this engine."""
        out = (ctypes.c_uint64 * 4)()
        _lib.check(self._L.emp_counters(self._h, out))
        return dict(proposals=int(out[0]), in_prior=int(out[1]), accepted=int(out[2]), nan=int(out[3]))

    def nan_count(self) -> int:
        c = ctypes.c_uint32()
        _lib.check(self._L.emp_nan_count(self._h, ctypes.byref(c)))
        return int(c.value)

    # ---- batched likelihood + prior ---------------------------------------------
    def logl_batch_device(self, theta, logl=None, logp=None):
        """theta: CUDA float64 tensor [n, ndim] -> (logl[n], logp[n]) CUDA tensors (async)."""
        import torch
        if theta.dtype != torch.float64 or not theta.is_cuda or not theta.is_contiguous():
            raise ValueError("theta must be a contiguous CUDA float64 tensor")
        if theta.shape[-1] != self.ndim:
            raise ValueError(f"theta has {theta.shape[-1]} columns, model has ndim={self.ndim}")
        n = theta.numel() // self.ndim
        if logl is None:
            logl = torch.empty(n, dtype=torch.float64, device=theta.device)
        if logp is None:
            logp = torch.empty(n, dtype=torch.float64, device=theta.device)
        _lib.check(self._L.emp_logl_batch(self._h, theta.data_ptr(), n, logl.data_ptr(), logp.data_ptr()))
        return logl, logp

    def logl_batch(self, thetas) -> Tuple[np.ndarray, np.ndarray]:
        """Host entry (the plugin call): thetas[n, ndim] -> (logl[n], logp[n]) NumPy arrays.
        Includes the H2D / D2H copies; synchronous."""
        th = _np64(thetas).reshape(-1, self.ndim)
        n = len(th)
        ll = np.empty(n, dtype=np.float64)
        lp = np.empty(n, dtype=np.float64)
        _lib.check(self._L.emp_logl_batch_host(self._h, th.ctypes.data, n, ll.ctypes.data, lp.ctypes.data))
        return ll, lp

    # ---- scalar-compatible callables (B1 of SURVEY.md §8b) ---------------------------
    def my_likelihood(self, theta) -> float:
        """my_likelihood(theta) (00.like:3-5) for theta INSIDE the prior support.  Unlike the
        reference's scalar callable, the device path never evaluates the model where the prior is
        -inf (the sampler never asks for it: emcee skips the likelihood there); such a call raises
        ValueError instead of returning a number nobody uses.  `logl_batch` returns -inf for those rows."""
        th = _np64(theta).reshape(-1, self.ndim)
        ll, lp = self.logl_batch(th)
        if np.any(~np.isfinite(lp)):
            raise ValueError("my_likelihood called outside the prior support; the device path "
                             "does not evaluate the model there (emcee never does either)")
        return float(ll[0]) if len(ll) == 1 else ll

    def my_prior(self, theta) -> float:
        th = _np64(theta).reshape(-1, self.ndim)
        _, lp = self.logl_batch(th)
        return float(lp[0]) if len(lp) == 1 else lp

    def my_model(self, theta):
        """my_model(theta) -> (model0[ndat], err20[ndat]) (emp_model.py:706-781)."""
        th = _np64(theta).reshape(self.ndim)
        model = np.empty(self.ndat)
        err2 = np.empty(self.ndat)
        _lib.check(self._L.emp_model_host(self._h, th.ctypes.data, model.ctypes.data, err2.ctypes.data))
        return model, err2

    # ---- PT step primitives (used by sampler.PTSampler) -------------------------------
    def pt_stretch_step(self, p, logl, logp, betas, half_idx, zz, rint, factors, lnu, accepted):
        T, W, nd = p.shape
        _lib.check(self._L.emp_pt_stretch_step(
            self._h, T, W, p.data_ptr(), logl.data_ptr(), logp.data_ptr(), betas.data_ptr(),
            half_idx.data_ptr(), zz.data_ptr(), rint.data_ptr(), factors.data_ptr(), lnu.data_ptr(),
            accepted.data_ptr()))

    def pt_swap_plan(self, logl_all, betas, perm, lnu, src, n_acc):
        T, W = logl_all.shape
        _lib.check(self._L.emp_pt_swap_plan(
            self._h, T, W, logl_all.data_ptr(), betas.data_ptr(),
            perm.data_ptr() if perm is not None else None, lnu.data_ptr() if lnu is not None else None,
            src.data_ptr(), n_acc.data_ptr()))

    def pt_sweep(self, args):
        """One whole single-GPU sweep (emp_pt_sweep); `args` is an `_lib.EmpPtSweepC`."""
        _lib.check(self._L.emp_pt_sweep(self._h, ctypes.byref(args)))

    def pt_sweep_stretch(self, args):
        _lib.check(self._L.emp_pt_sweep_stretch(self._h, ctypes.byref(args)))

    def pt_sweep_swap(self, args):
        _lib.check(self._L.emp_pt_sweep_swap(self._h, ctypes.byref(args)))

    def pt_gather_rows(self, src, p_in, ll_in, lp_in, p_out, ll_out, lp_out):
        n_rows = src.numel()
        _lib.check(self._L.emp_pt_gather_rows(
            self._h, n_rows, p_in.shape[-1], src.data_ptr(), p_in.data_ptr(), ll_in.data_ptr(),
            lp_in.data_ptr(), p_out.data_ptr(), ll_out.data_ptr(), lp_out.data_ptr()))


class SharedDeviceBuffer:
    """A cudaMalloc'ed block that the other ranks of the node can map (CUDA IPC, emp_dev_alloc /
    emp_ipc_export / emp_ipc_open): the sharded swap reads peer ensembles through it over NVLink.
    `tensor(shape, offset)` gives zero-copy torch views (torch consumes __cuda_array_interface__)."""

    def __init__(self, device: int, nbytes: int, ptr: Optional[int] = None, owner: bool = True):
        self.device, self.nbytes, self.owner = int(device), int(nbytes), owner
        if ptr is None:
            p = ctypes.c_void_p()
            _lib.check(_lib.lib().emp_dev_alloc(self.device, self.nbytes, ctypes.byref(p)))
            ptr = p.value
        self.ptr = int(ptr)

    def export(self) -> bytes:
        self.exported = True
        buf = ctypes.create_string_buffer(64)
        _lib.check(_lib.lib().emp_ipc_export(self.device, ctypes.c_void_p(self.ptr), buf))
        return buf.raw

    @classmethod
    def open(cls, device: int, handle: bytes, nbytes: int) -> "SharedDeviceBuffer":
        p = ctypes.c_void_p()
        _lib.check(_lib.lib().emp_ipc_open(int(device), handle, ctypes.byref(p)))
        return cls(device, nbytes, ptr=p.value, owner=False)

    def tensor(self, shape, offset_bytes: int = 0, typestr: str = "<f8"):
        import torch

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                      "data": (self.ptr + int(offset_bytes), False), "version": 2, "strides": None}
        v._keepalive = self
        return torch.as_tensor(v, device=torch.device("cuda", self.device))

    def close(self):
        if self.ptr:
            L = _lib.lib()
            (L.emp_dev_free if self.owner else L.emp_ipc_close)(self.device, ctypes.c_void_p(self.ptr))
            self.ptr = 0

    exported = False

    def __del__(self):
        # A sampler that goes away unmaps its peers' blocks and frees its own — except own blocks that were exported:
        # freeing memory a peer may still have mapped is undefined behaviour, and a destructor cannot run the
        # barrier that would order the two.  Those stay allocated until `close()` is called after a barrier, or
        # the process exits (two state blocks + two gathered blocks per sharded sampler).
        try:
            if not (self.owner and self.exported):
                self.close()
        except Exception:
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fp64_peak_tflops(device: int = 0) -> float:
    """Measured FP64 FMA peak of the device (roofline denominator, bench.py)."""
    v = ctypes.c_double()
    _lib.check(_lib.lib().emp_fp64_peak(int(device), ctypes.byref(v)))
    return float(v.value)


def kepler_solve(M, ecc, device: int = 0):
    """Device drop-in for `kepler.solve(M, ecc)` (kepler.py; kep00.model:6): E, elementwise."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    ecc = np.asarray(ecc, dtype=np.float64)
    scalar = ecc.ndim == 0 or ecc.size == 1
    ecc = np.ascontiguousarray(ecc.reshape(-1) if scalar else np.broadcast_to(ecc, M.shape))
    E = np.empty_like(M)
    _lib.check(_lib.lib().emp_kepler_solve_host(M.ctypes.data, ecc.ctypes.data, M.size, 1 if scalar else 0,
                                                E.ctypes.data, int(device)))
    return E


def kepler_solve_grid(M, ecc, device: int = 0):
    """The likelihood kernel's own solver (grid-anchored core) for arrays: returns (E, sin E, cos E).
    Same call shape as `kepler.solve(M, ecc)`; exists so that tests can pin the production core
    element by element (the kernel never materialises E)."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    ecc = np.asarray(ecc, dtype=np.float64)
    scalar = ecc.ndim == 0 or ecc.size == 1
    ecc = np.ascontiguousarray(ecc.reshape(-1) if scalar else np.broadcast_to(ecc, M.shape))
    E, s, c = np.empty_like(M), np.empty_like(M), np.empty_like(M)
    _lib.check(_lib.lib().emp_kepler_grid_host(M.ctypes.data, ecc.ctypes.data, M.size, 1 if scalar else 0,
                                               E.ctypes.data, s.ctypes.data, c.ctypes.data, int(device)))
    return E, s, c


def kepler_solve_grid_table(M, ecc: float, device: int = 0):
    """The grid core started from the per-walker starter table (what table-served planets run inside the likelihood
    kernel): (E, sin E, cos E) for one eccentricity in [0, 0.8]."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    E, s, c = np.empty_like(M), np.empty_like(M), np.empty_like(M)
    _lib.check(_lib.lib().emp_kepler_grid_table_host(M.ctypes.data, float(ecc), M.size, E.ctypes.data, s.ctypes.data,
                                                     c.ctypes.data, int(device)))
    return E, s, c
