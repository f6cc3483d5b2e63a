"""Sharding of the temperature ladder over GPUs.

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  Every rank holds T/G
temperatures (positions, logL, logP) and a full replica of the data set.  The
within-temperature stretch steps need no communication.  Once per sweep:

  1. logL[T, W] (FP64; config 5: 4 MiB in total) and the swap draws reach every rank: by peer writes
     over NVLink into CUDA-IPC-mapped gathered blocks (sampler.py, exchange='peer': no NCCL call in a
     sweep), or by NCCL all-gathers (`all_gather_rows`, exchange='allgather');
  2. every rank replays the SAME hot -> cold swap sweep on the gathered logL with the same host draws
     (kernel pt_swap_plan_chain) and so knows the whole permutation `src[T, W]`;
  3. rows (position, logL, logP) whose source lives on another rank are read from the owner's HBM by
     the plan-application kernel (exchange='peer'), gathered from an all-gathered copy of the blocks
     (exchange='allgather'), or — `exchange_rows`, the NCCL point-to-point formulation kept for
     platforms without peer access, not used by the sampler — sent pairwise: because the plan is
     replicated, sender and receiver derive the same row lists without any handshake.

Layouts:
  * "strided" (default): rank r holds temperatures r, r+G, r+2G, ...  Cold chains converge and
    then propose almost every move inside the prior box while hot chains keep ~40 % of their
    proposals outside (never evaluated), so contiguous blocks leave the rank with the coldest
    block ~25 % more likelihood work than the others (measured at 4 GPUs).  Interleaving gives
    every rank the same mix.  The price is that every accepted swap crosses a rank boundary:
    ~30 MB per rank and sweep at the 8-GPU bench shape — 0.05 ms of NVLink time.
  * "contiguous": rank r holds [r*T/G, (r+1)*T/G); only the G-1 block edges exchange rows.

The functions here are backend-agnostic torch code (NCCL on the GPUs, gloo in the CPU tests
of this host logic).
"""
from __future__ import annotations

from typing import Tuple


class LadderShard:
    def __init__(self, ntemps: int, group=None, layout: str = "strided"):
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
        if layout not in ("strided", "contiguous"):
            raise ValueError("layout must be 'strided' or 'contiguous'")
        self.layout = layout if self.world > 1 else "contiguous"
        self.ntemps = ntemps
        self.n_local = ntemps // self.world
        if self.layout == "contiguous":
            self.t0 = self.rank * self.n_local
            self.local_slice = slice(self.t0, self.t0 + self.n_local)
        else:
            self.local_slice = slice(self.rank, ntemps, self.world)
        self._warm = False

    # -- temperature <-> (rank, local row) maps (tensor or int arguments) ----------------
    def owner_of_temp(self, t):
        return t // self.n_local if self.layout == "contiguous" else t % self.world

    def local_of_temp(self, t):
        return t % self.n_local if self.layout == "contiguous" else t // self.world

    def temp_of(self, rank, j):
        return rank * self.n_local + j if self.layout == "contiguous" else j * self.world + rank

    def warm_up(self, device):
        """Open the NCCL point-to-point channels to every peer once (lazy connection setup would
        otherwise land inside the first swap sweep that moves a walker across ranks)."""
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

    # -- collectives ----------------------------------------------------------------
    def _to_temperature_order(self, out):
        """[G*T_loc, ...] in rank-major order -> temperature order."""
        if self.layout == "contiguous":
            return out
        G, n = self.world, self.n_local
        return out.view((G, n) + tuple(out.shape[1:])).transpose(0, 1).reshape(out.shape)

    def all_gather_rows(self, x):
        """x [T_loc, ...] on every rank -> [T, ...] in temperature order."""
        if self.world == 1:
            return x
        out = self.torch.empty((self.ntemps,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self.td.all_gather_into_tensor(out, x.contiguous(), group=self.group)
        return self._to_temperature_order(out)

    def all_gather_flat(self, *xs):
        """All-gather each [n_loc, ...] tensor into [G*n_loc, ...] in RANK-major order (the layout
        the swap gather indexes with owner*n_loc + local_row)."""
        outs = []
        for x in xs:
            out = self.torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            self.td.all_gather_into_tensor(out, x.contiguous(), group=self.group)
            outs.append(out)
        return outs

    def gather_to_all(self, x, dim=0):
        """All-gather along `dim` (the local-temperature axis) into temperature order."""
        if self.world == 1:
            return x
        return self.all_gather_rows(x.movedim(dim, 0)).movedim(0, dim)

    # -- swap-plan row exchange ---------------------------------------------------------
    def exchange_rows(self, src_plan, rows, W: int) -> Tuple["object", "object"]:
        """src_plan [T, W] int32: global flat source (t*W + w) of every destination slot,
        identical on all ranks; rows [T_loc*W, C]: this rank's rows before the swap, local
        temperature order.  Returns (staged [T_loc*W + n_remote, C], src_local [T_loc*W] int32)
        such that new_rows = staged[src_local].

        Every rank derives both its receive lists and its send lists from the replicated plan
        (no handshake); the only host synchronisation is one read of the 2*G row counts."""
        torch, td = self.torch, self.td
        G, nl, me = self.world, self.n_local * W, self.rank
        dev = src_plan.device
        plan = src_plan.reshape(self.ntemps, W).to(torch.int64)
        src_t, src_w = plan // W, plan % W
        src_owner = self.owner_of_temp(src_t)                      # [T, W] rank that holds the source row
        src_lrow = self.local_of_temp(src_t) * W + src_w           # its row index on that rank
        t_idx = torch.arange(self.ntemps, device=dev).unsqueeze(1).expand(self.ntemps, W)
        dst_owner = self.owner_of_temp(t_idx)                      # [T, W] rank that holds the destination
        dst_lrow = self.local_of_temp(t_idx) * W + torch.arange(W, device=dev).unsqueeze(0)
        # --- what I receive: my destinations whose source is remote, grouped by source rank,
        #     ascending destination order inside a group
        mine_owner = src_owner[self.local_slice].reshape(-1)
        mine_lrow = src_lrow[self.local_slice].reshape(-1)
        need_order = torch.argsort(mine_owner, stable=True)
        need_cnt = torch.bincount(mine_owner, minlength=G)
        # --- what I send: remote destinations whose source is mine, grouped by destination rank,
        #     ascending order of the receiver's destination slots inside a group
        give_mask = (src_owner == me) & (dst_owner != me)
        give_key = (dst_owner * nl + dst_lrow)[give_mask]
        give_rows = src_lrow[give_mask][torch.argsort(give_key)]
        give_cnt = torch.bincount(dst_owner[give_mask], minlength=G)
        cnt = torch.stack([need_cnt, give_cnt]).cpu()              # the one sync
        need_n, give_n = cnt[0].tolist(), cnt[1].tolist()
        n_remote = sum(need_n) - need_n[me]
        src_local = torch.empty(nl, dtype=torch.int64, device=dev)
        # local sources first, remote ones in (source rank, destination) order behind the local rows
        pos = 0
        offset = nl
        recv_all = torch.empty((n_remote, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        send_all = rows.index_select(0, give_rows) if give_rows.numel() else None
        ops, roff, soff = [], 0, 0
        for r in range(G):
            seg = need_order[pos:pos + need_n[r]]
            pos += need_n[r]
            if r == me:
                src_local[seg] = mine_lrow[seg]
                continue
            if need_n[r]:
                src_local[seg] = offset + torch.arange(need_n[r], device=dev)
                ops.append(td.P2POp(td.irecv, recv_all[roff:roff + need_n[r]], self._global_rank(r), group=self.group))
                offset += need_n[r]
                roff += need_n[r]
            if give_n[r]:
                ops.append(td.P2POp(td.isend, send_all[soff:soff + give_n[r]], self._global_rank(r), group=self.group))
                soff += give_n[r]
        if ops:
            for req in td.batch_isend_irecv(ops):
                req.wait()
        staged = torch.cat([rows, recv_all], 0) if n_remote else rows
        return staged, src_local.to(torch.int32)

    def _global_rank(self, r):
        if self.group is None:
            return r
        return self.td.get_global_rank(self.group, r)
