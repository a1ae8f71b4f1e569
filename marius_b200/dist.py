"""marius_b200.dist -- the embedding table sharded by node partition across the GPUs of one box.

Rank r owns the contiguous id range [r * rows_per_rank, (r+1) * rows_per_rank) of the global node id space, with its own
Adagrad-state shard.  One batch on rank r, given its SORTED unique global ids (what map_tensors / torch::_unique2 produce):

  1. bucket the unique ids by owner -- contiguous slices of the sorted list (integer divide, like storage.cpp:75);
  2. all-to-all #1: ids (int64)            -> every owner learns which of its rows each peer needs;
  3. owner gather (mb_gather_rows) + all-to-all #2: rows (d x fp32) back  -> the batch's [U, d] embedding matrix, in unique order;
  4. local forward / loss / backward on the fused kernels -> gradient rows [U, d] (Adagrad state is NOT shipped: the owner applies it);
  5. all-to-all #3: gradient rows to the owners; the owner applies the contributions rank by rank with the fused Adagrad
     read-modify-write (mb_adagrad_update_rows; ids are unique within one contribution, so no atomics and a fixed order);
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

    def gather(self, local_rows: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.ops.gather_rows(self.table, local_rows, out)

    def train_batch(self, kind, emb, edges, rel, inv_rel, dst_negs, src_negs, reduction):
        return self.ops.train_batch(self.ctx, kind, emb, None, edges, rel, inv_rel, dst_negs, src_negs, 0.0, reduction, self.precision)

    def update(self, local_rows: torch.Tensor, grads: torch.Tensor, lr: float) -> None:
        self.ops.adagrad_update_rows(self.table, self.state, local_rows, grads, lr)


class RoutePlan:
    """Routing metadata of one batch.  It depends only on the batch's sorted unique ids, so -- like sampling and unique-id mapping --
    the loader can prepare it ahead of the step (on the host; bucket sizes and the 8-byte ids travel first) to keep the GPU step
    free of host round trips: the step itself then only moves rows and gradient rows."""

    def __init__(self, bounds, recv_counts, req_local):
        self.bounds = [int(x) for x in bounds]            # owner j's slice of the unique list is [bounds[j], bounds[j+1])
        self.recv_counts = [int(x) for x in recv_counts]  # rows peer j needs from me (0 for myself)
        self.req_local = req_local                        # my local row ids requested by the peers, concatenated in rank order

    @property
    def send_counts(self):
        return [self.bounds[j + 1] - self.bounds[j] for j in range(len(self.bounds) - 1)]


class ShardedTable:
    def __init__(self, rows_per_rank: int, backend, group=None, cpu_group=None):
        self.rows_per_rank = int(rows_per_rank)
        self.backend = backend
        self.group = group
        self.cpu_group = cpu_group  # optional gloo group for the tiny count exchange of make_plan()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.base = self.rank * self.rows_per_rank
        self.last_remote_rows = 0  # rows of the last batch fetched from other ranks

    # -- exchanges ------------------------------------------------------------------------------------------------
    def _exchange(self, send_parts, recv_parts):
        """all-to-all over lists of (possibly empty) contiguous tensors; entry `rank` is ignored (local rows never enter NCCL)."""
        if self.world == 1:
            return
        # grouped point-to-point (ncclGroupStart/End under NCCL, plain isend/irecv under gloo); empty parts are skipped on both ends
        ops_ = []
        for j in range(self.world):
            if j == self.rank:
                continue
            if recv_parts[j].numel() > 0:
                ops_.append(dist.P2POp(dist.irecv, recv_parts[j], j, self.group))
            if send_parts[j].numel() > 0:
                ops_.append(dist.P2POp(dist.isend, send_parts[j].contiguous(), j, self.group))
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()

    def make_plan(self, unique_ids: torch.Tensor, ids_device: Optional[torch.Tensor] = None) -> RoutePlan:
        """Bucket the sorted unique ids by owner (contiguous slices), exchange the bucket sizes and the ids (all-to-all #1).
        `unique_ids`: host tensor preferred (no device sync; counts over `cpu_group` when given).  `ids_device`: the same ids on the
        compute device (defaults to unique_ids)."""
        if ids_device is None:
            ids_device = unique_ids
        b = owner_bounds(unique_ids, self.rows_per_rank, self.world).cpu()
        send_counts = (b[1:] - b[:-1]).to(torch.int64)
        send_counts[self.rank] = 0  # my own slice is served locally
        if self.world == 1:
            recv_counts = send_counts.clone()
        elif not unique_ids.is_cuda and (self.cpu_group is not None or dist.get_backend(self.group) == "gloo"):
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.cpu_group if self.cpu_group is not None else self.group)
        else:
            sc = send_counts.to(ids_device.device)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc, group=self.group)
            recv_counts = rc.cpu()
        bl, rl = b.tolist(), recv_counts.tolist()
        req = ids_device.new_empty(int(sum(rl)))
        send_parts = [ids_device[bl[j]:bl[j + 1]] if j != self.rank else ids_device[:0] for j in range(self.world)]
        recv_parts = list(torch.split(req, rl)) if self.world > 1 else []
        self._exchange(send_parts, recv_parts)  # all-to-all #1: ids
        return RoutePlan(bl, rl, req - self.base)

    def fetch_rows(self, unique_ids: torch.Tensor, route: RoutePlan) -> torch.Tensor:
        r, bl = self.rank, route.bounds
        own = unique_ids[bl[r]:bl[r + 1]] - self.base
        if self.world == 1:
            self.last_remote_rows = 0
            return self.backend.gather(own)
        rows = self.backend.gather(route.req_local)  # what the peers asked of me
        emb = rows.new_empty((unique_ids.numel(), rows.size(1)))
        self.backend.gather(own, emb[bl[r]:bl[r + 1]])  # local rows: gathered straight into their slice of the batch matrix
        send_parts = list(torch.split(rows, route.recv_counts))
        recv_parts = [emb[bl[j]:bl[j + 1]] if j != r else emb[:0] for j in range(self.world)]
        self._exchange(send_parts, recv_parts)  # all-to-all #2: rows land in unique-id order
        self.last_remote_rows = int(unique_ids.numel() - (bl[r + 1] - bl[r]))
        return emb

    def push_grads(self, unique_ids: torch.Tensor, grads: torch.Tensor, route: RoutePlan, lr: float) -> None:
        """Gradient rows go to their owners (all-to-all #3).  The owner applies the contributions in rank order, each with the fused
        Adagrad read-modify-write (ids are unique within one contribution) -- the per-batch update the reference applies,
        dataloader.cpp:550-557."""
        r, bl = self.rank, route.bounds
        recv = grads.new_empty((int(sum(route.recv_counts)), grads.size(1)))
        if self.world > 1:
            send_parts = [grads[bl[j]:bl[j + 1]] if j != r else grads[:0] for j in range(self.world)]
            recv_parts = list(torch.split(recv, route.recv_counts))
            self._exchange(send_parts, recv_parts)
        own = unique_ids[bl[r]:bl[r + 1]] - self.base
        off = 0
        for j in range(self.world):
            if j == r:
                self.backend.update(own, grads[bl[r]:bl[r + 1]], lr)
            elif route.recv_counts[j] > 0:
                n = route.recv_counts[j]
                self.backend.update(route.req_local[off:off + n], recv[off:off + n], lr)
            if j != r:
                off += route.recv_counts[j]

    # -- one training batch -----------------------------------------------------------------------------------------
    def train_step(self, kind: int, unique_ids: torch.Tensor, edges: torch.Tensor, rel, inv_rel, dst_negs, src_negs, lr: float, reduction: int = 1,
                   route: Optional[RoutePlan] = None):
        """unique_ids: sorted GLOBAL ids; edges / negatives hold batch-local positions into unique_ids (the reference Batch layout).
        Returns the dict of Model::train_batch outputs (loss, rel_grad, inv_rel_grad); relation gradients are summed over ranks."""
        if route is None:
            route = self.make_plan(unique_ids)
        emb = self.fetch_rows(unique_ids, route)
        out = self.backend.train_batch(kind, emb, edges, rel, inv_rel, dst_negs, src_negs, reduction)
        self.push_grads(unique_ids, out["grad"], route, lr)
        if self.world > 1:
            for k in ("rel_grad", "inv_rel_grad"):
                if out.get(k) is not None:
                    dist.all_reduce(out[k], group=self.group)
        return out


class PeerShardedTable:
    """The sharded table with the rows of the peers mapped into this process (CUDA IPC over NVLink): the fused step
    (mb_train_step_sharded) gathers remote rows with plain loads and applies their Adagrad update with plain stores -- no staging,
    no collective on the row path.  Only the dense relation gradients are all-reduced (the caller does that, as in the reference).

    One process per GPU; every process must see all GPUs of the box (torchrun's default)."""

    def __init__(self, table: torch.Tensor, state: torch.Tensor, ctx, group=None, exchange_rows: int = 1 << 18):
        from . import ops

        self.ops, self.ctx, self.group = ops, ctx, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rows_per_rank = table.size(0)
        self.d = table.size(1)
        self.ld = table.stride(0)
        self.table, self.state = table, state
        tptr, sptr, eptr = [0] * self.world, [0] * self.world, [0] * self.world
        tptr[self.rank], sptr[self.rank] = table.data_ptr(), state.data_ptr()
        self.exchange_rows = int(exchange_rows)
        self.exchange = None
        if self.world > 1:
            # this rank's exchange area: barrier flags + one inbox of gradient rows per sender (include/marius_b200.h: mb_shards)
            self.exchange = torch.zeros(ops.shard_exchange_bytes(self.world, self.exchange_rows, self.d), dtype=torch.uint8, device=table.device)
            eptr[self.rank] = self.exchange.data_ptr()
            handles = [None] * self.world
            dist.all_gather_object(handles, (ops.ipc_export(table), ops.ipc_export(state), ops.ipc_export(self.exchange), tuple(table.shape), table.stride(0),
                                             self.exchange_rows), group=group)
            for r, (ht, hs, he, shape, ld, er) in enumerate(handles):
                if shape != tuple(table.shape) or ld != self.ld or er != self.exchange_rows:
                    raise ValueError("all shards must have the same shape, stride and exchange capacity")
                if r == self.rank:
                    continue
                # opened with MY device current (mb_ipc_import): loads / stores from my kernels reach the peer's HBM over NVLink
                tptr[r] = ops.ipc_import(ctx, *ht)
                sptr[r] = ops.ipc_import(ctx, *hs)
                eptr[r] = ops.ipc_import(ctx, *he)
            dist.barrier(group=group)
        self.shards = ops.make_shards_raw(tptr, sptr, self.rows_per_rank, self.rank, eptr, self.exchange_rows if self.world > 1 else 0)

    def error(self) -> int:
        """non-zero once a cross-rank barrier of a sharded step timed out on this rank"""
        return self.ops.shard_error(self.ctx, self.shards) if self.world > 1 else 0

    def train_step(self, kind, unique_ids, edges, rel, inv_rel, dst_negs, src_negs, lr, reduction=1, precision=None, loss=None, rel_grad=None,
                   inv_rel_grad=None):
        p = self.ops.PREC_BF16X3 if precision is None else precision
        return self.ops.train_step_sharded(self.ctx, kind, self.shards, self.ld, self.d, unique_ids, edges, rel, inv_rel, dst_negs, src_negs, lr,
                                           reduction, p, loss=loss, rel_grad=rel_grad, inv_rel_grad=inv_rel_grad)

    def train_step_host_async(self, kind, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, lr, reduction=1, precision=None, rel_grad=None,
                              inv_rel_grad=None):
        """enqueue only: returns (ticket, keep-alive); ops.train_step_host_wait(ctx, ticket) returns the loss"""
        p = self.ops.PREC_BF16X3 if precision is None else precision
        return self.ops.train_step_sharded_host_async(self.ctx, kind, self.shards, self.ld, self.d, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h,
                                                      src_negs_h, lr, reduction, p, rel_grad=rel_grad, inv_rel_grad=inv_rel_grad)

    def train_step_host(self, kind, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h, src_negs_h, lr, reduction=1, precision=None, rel_grad=None,
                        inv_rel_grad=None) -> float:
        p = self.ops.PREC_BF16X3 if precision is None else precision
        return self.ops.train_step_sharded_host(self.ctx, kind, self.shards, self.ld, self.d, unique_ids_h, edges_h, rel, inv_rel, dst_negs_h,
                                                src_negs_h, lr, reduction, p, rel_grad=rel_grad, inv_rel_grad=inv_rel_grad)
