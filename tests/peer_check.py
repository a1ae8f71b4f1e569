"""tests/peer_check.py -- 2+ rank check of the peer-memory sharded step (run under torchrun, one rank per GPU): bench.parity_sharded at
d = 400 / 1000 negatives, once with rows shared between the ranks' batches (owner-side ordered Adagrad apply) and once with disjoint rows."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    from marius_b200 import ops

    ctx = ops.Context(local)
    ok = True
    for overlap in (True, False):
        res = bench.parity_sharded(ops, ctx, dev, ops.PREC_BF16X3, rank, world, rows=32768, B=2000, overlap=overlap)
        ok = ok and bool(res.get("ok"))
        if rank == 0:
            print(f"PEER_CHECK overlap={overlap} {'OK' if res.get('ok') else 'FAIL'} world={world} max_err={res.get('max_err'):.2e} per_rank={res.get('per_rank')}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"PEER_CHECK {'OK' if int(flag.item()) == 1 else 'FAIL'} world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
