import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from marius_b200 import ops
ctx = ops.Context(local)
t = torch.full((1000, 64), float(rank + 1), device=dev)
handles = [None] * world
import ctypes as C
from marius_b200._lib import lib, check
dist.all_gather_object(handles, ops.ipc_export(t))
ptrs = []
for r, h in enumerate(handles):
    ptrs.append(t.data_ptr() if r == rank else ops.ipc_import(ctx, *h))
dist.barrier()
for r, p in enumerate(ptrs):
    idx = torch.arange(0, 1000, 7, device=dev)
    out = torch.empty(idx.numel(), 64, device=dev)
    try:
        check(lib.mb_gather_rows(C.c_void_p(p), 1000, 64, 64, C.c_void_p(idx.data_ptr()), idx.numel(), C.c_void_p(out.data_ptr()), 64, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        print(f"rank {rank}: gather from shard {r}: value {float(out[0,0])} ok", flush=True)
    except Exception as ex:
        print(f"rank {rank}: gather from shard {r} FAILED {ex}", flush=True)
dist.barrier()
dist.destroy_process_group()
