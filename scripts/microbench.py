"""Per-kernel timing of the hot kernels on one block (development aid).
    python scripts/microbench.py [--n 128] [--block mass_u] [--reps 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import formoniq_b200 as fq

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--fused", action="store_true")
args = ap.parse_args()
ctx = fq.Context(0, stream=torch.cuda.current_stream().cuda_stream)
mesh = fq.Mesh.kuhn(ctx, 3, args.n)
ctx.set_timing(True)
if args.fused:
    hb = fq.HodgeBlocks.symbolic(mesh, 1)
    blocks = hb.blocks
    run = lambda: hb.numeric(mesh, True)
else:
    a = fq.WhitneyPairing.mass(3, 1).symbolic(mesh)
    blocks = [a]
    run = lambda: a.numeric(mesh, True)
for _ in range(3):
    run()
ctx.timing_report()
for _ in range(args.reps):
    run()
rep = ctx.timing_report()
for k, v in rep.items():
    print(f"{k:12s} {v['ms'] / args.reps:9.3f} ms/step ({v['count']} launches)")
a = max(blocks, key=lambda b: b.nnz)
x = fq.DeviceVector.from_numpy(ctx, np.cos(np.arange(a.shape[1]) ** 2 + 1.0))
y = fq.DeviceVector(ctx, a.shape[0])
for _ in range(3):
    a.apply(x, y)
ctx.timing_report()
for _ in range(args.reps):
    a.apply(x, y)
r = ctx.timing_report()["k4_spmv"]
ms = r["ms"] / r["count"]
print(f"spmv nnz={a.nnz} {ms:.4f} ms  {a.spmv_bytes / 1e9 / (ms / 1e3):.1f} GB/s ({a.spmv_bytes / 1e9 / (ms / 1e3) / 6555.2:.3f} of measured HBM peak)")
