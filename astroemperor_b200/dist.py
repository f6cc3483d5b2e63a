"""Sharding of the temperature ladder over GPUs.

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Rank g
holds the contiguous temperatures [g*T/G, (g+1)*T/G): positions, logL, logP and
a full replica of the data set.  The within-temperature stretch steps need no
communication.  Once per sweep:

  1. all-gather of logL[T, W] (FP64; config 5: 4 MiB in total);
  2. every rank replays the SAME sequential hot -> cold swap sweep on the gathered
     logL with the same host draws (kernel pt_swap_plan) and so knows the whole
     permutation `src[T, W]`;
  3. rows (position, logL, logP) whose source lives on another rank are exchanged
     point-to-point: because the plan is replicated, sender and receiver derive the
     same row lists without any handshake (`exchange_rows`).

The functions here are backend-agnostic torch code (NCCL on the GPUs, gloo in the
CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Optional, Tuple


class LadderShard:
    def __init__(self, ntemps: int, group=None):
        import torch
        import torch.distributed as td
        self.torch, self.td = torch, td
        self.group = group
        if td.is_available() and td.is_initialized():
            self.world = td.get_world_size(group)
            self.rank = td.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        if ntemps % self.world:
            raise ValueError(f"ntemps={ntemps} must be a multiple of the number of ranks ({self.world})")
        self.ntemps = ntemps
        self.n_local = ntemps // self.world
        self.t0 = self.rank * self.n_local
        self.local_slice = slice(self.t0, self.t0 + self.n_local)
        self._warm = False

    def warm_up(self, device):
        """Open the NCCL point-to-point channels to every peer once (lazy connection setup would
        otherwise land inside the first swap sweep that moves a walker across a shard edge)."""
        if self.world == 1 or self._warm:
            return
        torch, td = self.torch, self.td
        ops, bufs = [], []
        for r in range(self.world):
            if r == self.rank:
                continue
            snd = torch.zeros(8, dtype=torch.float64, device=device)
            rcv = torch.empty(8, dtype=torch.float64, device=device)
            bufs += [snd, rcv]
            ops.append(td.P2POp(td.isend, snd, self._global_rank(r), group=self.group))
            ops.append(td.P2POp(td.irecv, rcv, self._global_rank(r), group=self.group))
        for req in td.batch_isend_irecv(ops):
            req.wait()
        self._warm = True

    def owner(self, t):
        return t // self.n_local

    # -- collectives ----------------------------------------------------------------
    def all_gather_rows(self, x):
        """x [T_loc, ...] on every rank -> [T, ...] (rank order == temperature order)."""
        if self.world == 1:
            return x
        out = self.torch.empty((self.ntemps,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self.td.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return out

    def gather_to_all(self, x, dim=0):
        if self.world == 1:
            return x
        xs = x.movedim(dim, 0).contiguous()
        out = self.torch.empty((xs.shape[0] * self.world,) + tuple(xs.shape[1:]), dtype=x.dtype, device=x.device)
        self.td.all_gather_into_tensor(out, xs, group=self.group)
        return out.movedim(0, dim)

    # -- swap-plan row exchange ---------------------------------------------------------
    def exchange_rows(self, src_plan, rows, W: int) -> Tuple["object", "object"]:
        """src_plan [T, W] int32 (global flat source of every destination slot, identical on all
        ranks); rows [T_loc*W, C] local rows before the swap.
        Returns (staged [T_loc*W + n_remote, C], src_local [T_loc*W] int32) such that
        new_rows = staged[src_local]."""
        torch, td = self.torch, self.td
        nl = self.n_local * W
        lo = self.t0 * W
        src = src_plan.reshape(-1).to(torch.int64)
        owner_of_src = src // nl  # rank that holds each source row
        dest_rank = torch.arange(src.numel(), device=src.device) // nl
        mine = src[lo:lo + nl]
        mine_owner = owner_of_src[lo:lo + nl]
        src_local = torch.empty(nl, dtype=torch.int64, device=src.device)
        local_mask = mine_owner == self.rank
        src_local[local_mask] = mine[local_mask] - lo
        recv_bufs, ops = [], []
        offset = nl
        for r in range(self.world):
            if r == self.rank:
                continue
            # rows I need from r, in ascending destination order
            need = (mine_owner == r).nonzero(as_tuple=True)[0]
            # rows r needs from me, in ascending destination order of r
            their = src[r * nl:(r + 1) * nl]
            give = their[owner_of_src[r * nl:(r + 1) * nl] == self.rank] - lo
            if give.numel():
                send = rows.index_select(0, give).contiguous()
                ops.append(td.P2POp(td.isend, send, self._global_rank(r), group=self.group))
            if need.numel():
                buf = torch.empty((need.numel(), rows.shape[1]), dtype=rows.dtype, device=rows.device)
                ops.append(td.P2POp(td.irecv, buf, self._global_rank(r), group=self.group))
                src_local[need] = offset + torch.arange(need.numel(), device=src.device)
                offset += need.numel()
                recv_bufs.append(buf)
        if ops:
            for req in td.batch_isend_irecv(ops):
                req.wait()
        staged = torch.cat([rows] + recv_bufs, 0) if recv_bufs else rows
        return staged, src_local.to(torch.int32)

    def _global_rank(self, r):
        if self.group is None:
            return r
        return self.td.get_global_rank(self.group, r)
