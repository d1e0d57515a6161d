"""Config 5 across GPUs: shift-invert Lanczos on the row-partitioned Hodge-Laplace pencil (dist.DistKktPencil).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/evp_multi_gpu.py [--grid 8]

Every rank owns a z-slab; rank 0 also solves the same pencil alone (one GPU, same code path with world = 1) and the
eigenvalues must agree to 1e-9.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import formoniq_b200 as fq
from formoniq_b200.dist import DistKktPencil

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=8, help="boxes per axis (global grid n x n x n)")
ap.add_argument("--shape", type=str, default="", help="boxes per axis as x,y,z (overrides --grid; z >= ranks)")
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--shift", type=float, default=5.0)
ap.add_argument("--precond", default="none", choices=["none", "afw"], help="preconditioner of the inner MINRES solves")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = fq.Context(local, stream=torch.cuda.current_stream().cuda_stream)
shape = [int(v) for v in args.shape.split(",")] if args.shape else [args.grid, args.grid, args.grid]
pencil = DistKktPencil(ctx, 3, shape, 1, rank, world, precond=args.precond)
torch.cuda.synchronize()
t0 = time.perf_counter()
vals, vecs = fq.shift_invert_lanczos(pencil, args.shift, args.k)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ref = None
if rank == 0:
    single = DistKktPencil(ctx, 3, shape, 1, 0, 1, precond=args.precond)
    ref, _ = fq.shift_invert_lanczos(single, args.shift, args.k)
if world > 1:
    dist.barrier()
if rank == 0:
    err = float(np.abs(vals - ref).max() / np.abs(ref).max())
    print(json.dumps({"workload": f"3-D Hodge-Laplace k=1 EVP, Kuhn grid {'x'.join(map(str, shape))}, {world} ranks", "eigenvalues": vals.tolist(),
                      "single_gpu_eigenvalues": ref.tolist(), "rel_diff_vs_single_gpu": err, "match_1e-9": err <= 1e-9,
                      "seconds": dt, "kkt_applies": pencil.applies, "inner_minres_iterations": pencil.inner_iterations, "precond": args.precond,
                      "afw_cg_iterations": pencil.afw_iterations,
                      "n_global": pencil.n_global}))
    assert err <= 1e-9
if world > 1:
    dist.destroy_process_group()
