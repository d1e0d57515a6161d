"""One-off evidence: the row blocks a rank assembles for its z-slab (owner computes, no communication) are the same bits as
the same rows of the one-GPU matrix, at scale.  torchrun --nproc-per-node R scripts/slab_parity_multi_gpu.py [N]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import formoniq_b200 as fq
from formoniq_b200.dist import slab_of

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = fq.Context(local)
shape = [N, N, N]
part = fq.Mesh.kuhn(ctx, 3, shape, slab=slab_of(rank, world, N), jitter=0.2)
hb = fq.HodgeBlocks.symbolic(part, 1, sigma_rows=part.owned_range(0), u_rows=part.owned_range(1))
hb.numeric(part)
hb.numeric(part)
mine = [b.download() for b in hb.blocks]
ranges = [b.row_range for b in hb.blocks]
del hb, part
fq._lib.lib().fq_device_cache_trim()
full = fq.Mesh.kuhn(ctx, 3, shape, jitter=0.2)
hf = fq.HodgeBlocks.symbolic(full, 1)
hf.numeric(full)
hf.numeric(full)
ok, nnz = True, 0
for (rp, ci, va), (lo, hi), blk in zip(mine, ranges, hf.blocks):
    frp, fci, fva = blk.download()
    a, b = int(frp[lo]), int(frp[hi])
    ok = ok and np.array_equal(rp.astype(np.int64), frp[lo:hi + 1].astype(np.int64) - a)
    ok = ok and np.array_equal(ci, fci[a:b]) and np.array_equal(va.view(np.uint64), fva[a:b].view(np.uint64))
    nnz += b - a
flag = torch.tensor([1.0 if ok else 0.0, float(nnz)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
    dist.all_reduce(flag[1:], op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"check": "row blocks of the z-slabs == rows of the one-GPU matrix (pattern and value bits)", "N": N,
                      "ranks": world, "tets": 6 * N ** 3, "nnz_compared": int(flag[1].item()), "bitwise_equal": bool(flag[0].item() == 1.0)}))
if world > 1:
    dist.destroy_process_group()
