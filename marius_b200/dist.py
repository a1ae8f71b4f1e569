"""marius_b200.dist -- the embedding table sharded by node partition across the GPUs of one box.

Rank r owns the contiguous id range [r * rows_per_rank, (r+1) * rows_per_rank) of the global node id space, with its own
Adagrad-state shard.  One batch on rank r, given its SORTED unique global ids (what map_tensors / torch::_unique2 produce):

  1. bucket the unique ids by owner -- contiguous slices of the sorted list (integer divide, like storage.cpp:75);
  2. all-to-all #1: ids (int64)            -> every owner learns which of its rows each peer needs;
  3. owner gather (mb_gather_rows) + all-to-all #2: rows (d x fp32) back  -> the batch's [U, d] embedding matrix, in unique order;
  4. local forward / loss / backward on the fused kernels -> gradient rows [U, d] (Adagrad state is NOT shipped: the owner applies it);
  5. all-to-all #3: gradient rows to the owners; the owner sums rows that arrived for the same table row from different ranks
     (mb_reduce_rows_by_key: sorted, no atomics) and runs the fused Adagrad read-modify-write (mb_adagrad_update_rows);
  6. dense relation gradients: all-reduce (the reference's only collective, nn/model.cpp:136-159).

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing for the three exchanges; every byte of table data is
moved by the C-ABI kernels.  The row/compute backend is pluggable so that the routing logic is testable on CPU with world_size 2.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def owner_bounds(unique_ids: torch.Tensor, rows_per_rank: int, world: int) -> torch.Tensor:
    """positions in the sorted unique-id list where each owner's slice starts; shape [world + 1]."""
    edges = torch.arange(world + 1, dtype=torch.int64, device=unique_ids.device) * rows_per_rank
    edges[-1] = torch.iinfo(torch.int64).max
    b = torch.searchsorted(unique_ids, edges[:-1], right=False)
    return torch.cat([b, torch.tensor([unique_ids.numel()], dtype=torch.int64, device=unique_ids.device)])


class OpsBackend:
    """Rows and compute on the local GPU through the C ABI."""

    def __init__(self, table: torch.Tensor, state: torch.Tensor, ctx, precision: Optional[int] = None):
        from . import ops

        self.ops = ops
        self.table, self.state, self.ctx = table, state, ctx
        self.precision = ops.PREC_BF16X3 if precision is None else precision

    def gather(self, local_rows: torch.Tensor) -> torch.Tensor:
        return self.ops.gather_rows(self.table, local_rows)

    def train_batch(self, kind, emb, edges, rel, inv_rel, dst_negs, src_negs, reduction):
        return self.ops.train_batch(self.ctx, kind, emb, None, edges, rel, inv_rel, dst_negs, src_negs, 0.0, reduction, self.precision)

    def merge(self, local_rows: torch.Tensor, grads: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        # padded: unique ids past the count are -1 and are skipped by the update kernel -> no host synchronisation
        return self.ops.reduce_rows_by_key(self.ctx, local_rows, grads, max_id=self.table.size(0), padded=True)

    def update(self, local_rows: torch.Tensor, grads: torch.Tensor, lr: float) -> None:
        self.ops.adagrad_update_rows(self.table, self.state, local_rows, grads, lr)


class RoutePlan:
    """Split sizes of one batch's exchanges.  They depend only on the batch's unique ids, so -- like sampling and unique-id mapping --
    they can be prepared by the loader ahead of the step (on the host, counts exchanged over a CPU group) to keep the GPU step free
    of host round trips."""

    def __init__(self, send_counts, recv_counts, req_local=None):
        self.send_counts = [int(x) for x in send_counts]
        self.recv_counts = [int(x) for x in recv_counts]
        self.req_local = req_local  # optional: the owner-side list of requested local rows (all-to-all #1 done ahead of the step)


class ShardedTable:
    def __init__(self, rows_per_rank: int, backend, group=None, cpu_group=None):
        self.rows_per_rank = int(rows_per_rank)
        self.backend = backend
        self.group = group
        self.cpu_group = cpu_group  # optional gloo group for the tiny count exchange of make_plan()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.base = self.rank * self.rows_per_rank
        self.last_remote_rows = 0  # rows of the last batch that crossed a rank boundary (either direction is symmetric in size)

    # -- exchanges ------------------------------------------------------------------------------------------------
    def _a2a(self, x: torch.Tensor, in_splits, out_splits) -> torch.Tensor:
        out = x.new_empty((int(sum(out_splits)),) + tuple(x.shape[1:]))
        if self.world == 1:
            out.copy_(x)
            return out
        dist.all_to_all_single(out, x.contiguous(), output_split_sizes=list(out_splits), input_split_sizes=list(in_splits), group=self.group)
        return out

    def make_plan(self, unique_ids: torch.Tensor, ids_device: Optional[torch.Tensor] = None) -> RoutePlan:
        """Bucket the sorted unique ids by owner and exchange the bucket sizes.  `unique_ids` may be a host tensor (preferred: no
        device synchronisation; the counts travel over `cpu_group` when one was given) or a device tensor.  With `ids_device` (the
        same ids on the compute device) the id exchange itself (all-to-all #1, 8 B per remote row) is also done here, ahead of the
        step, so that the step only moves rows and gradients."""
        b = owner_bounds(unique_ids, self.rows_per_rank, self.world)
        send_counts = (b[1:] - b[:-1]).to(torch.int64)
        if self.world == 1:
            recv_counts = send_counts.clone()
        elif not unique_ids.is_cuda and (self.cpu_group is not None or dist.get_backend(self.group) == "gloo"):
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.cpu_group if self.cpu_group is not None else self.group)
        else:
            dev = unique_ids.device if unique_ids.is_cuda else torch.device("cuda", torch.cuda.current_device())
            sc = send_counts.to(dev)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc, group=self.group)
            recv_counts = rc.cpu()
        route = RoutePlan(send_counts.tolist(), recv_counts.tolist())
        if ids_device is not None:
            route.req_local = self._a2a(ids_device, route.send_counts, route.recv_counts) - self.base
        return route

    def plan(self, unique_ids: torch.Tensor, route: Optional[RoutePlan] = None):
        """(send_counts, recv_counts, requested_local_rows): who needs which of my rows for this batch."""
        if route is None:
            route = self.make_plan(unique_ids)
        send_l, recv_l = route.send_counts, route.recv_counts
        self.last_remote_rows = int(unique_ids.numel() - send_l[self.rank])
        if route.req_local is not None:
            return send_l, recv_l, route.req_local
        req = self._a2a(unique_ids, send_l, recv_l)  # all-to-all #1
        return send_l, recv_l, req - self.base

    def fetch_rows(self, unique_ids: torch.Tensor, route: Optional[RoutePlan] = None):
        send_l, recv_l, req_local = self.plan(unique_ids, route)
        rows = self.backend.gather(req_local)
        emb = self._a2a(rows, recv_l, send_l)  # all-to-all #2: rows come back in unique-id order
        return emb, (send_l, recv_l, req_local)

    def push_grads(self, grads: torch.Tensor, plan, lr: float) -> None:
        send_l, recv_l, req_local = plan
        g = self._a2a(grads, send_l, recv_l)  # all-to-all #3
        rows, gsum = self.backend.merge(req_local, g)
        self.backend.update(rows, gsum, lr)

    # -- one training batch -----------------------------------------------------------------------------------------
    def train_step(self, kind: int, unique_ids: torch.Tensor, edges: torch.Tensor, rel, inv_rel, dst_negs, src_negs, lr: float, reduction: int = 1,
                   route: Optional[RoutePlan] = None):
        """unique_ids: sorted GLOBAL ids; edges / negatives hold batch-local positions into unique_ids (the reference Batch layout).
        Returns the dict of Model::train_batch outputs (loss, rel_grad, inv_rel_grad); relation gradients are summed over ranks."""
        emb, plan = self.fetch_rows(unique_ids, route)
        out = self.backend.train_batch(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, reduction)
        self.push_grads(out["grad"], plan, lr)
        if self.world > 1:
            for k in ("rel_grad", "inv_rel_grad"):
                if out.get(k) is not None:
                    dist.all_reduce(out[k], group=self.group)
        return out
