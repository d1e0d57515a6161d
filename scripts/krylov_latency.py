"""CG latency: device-resident recurrence (CUDA graph) vs host scalars, Jacobi-CG on hdif_gram(1) of Kuhn cubes."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import formoniq_b200 as fq

ctx = fq.Context(0, stream=torch.cuda.current_stream().cuda_stream)
out = []
for n in (8, 16, 32, 64):
    wc = fq.WhitneyComplex(fq.Mesh.kuhn(ctx, 3, [n, n, n]))
    a = wc.hdif_gram(1)
    b = fq.DeviceVector.from_numpy(ctx, np.cos(np.arange(a.shape[0], dtype=np.float64) ** 2 + 1.0))
    row = {"n": n, "rows": a.shape[0], "nnz": a.nnz}
    for mode in ("host", "device"):
        if mode == "host":
            os.environ["FQ_KRYLOV_HOST"] = "1"
        else:
            os.environ.pop("FQ_KRYLOV_HOST", None)
        fq.cg(a, "jacobi", b, fq.StopCriterion(1e-10, 20))  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, rep = fq.cg(a, "jacobi", b, fq.StopCriterion(1e-10, 100000))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        row[mode] = {"iters": rep.iters, "seconds": dt, "us_per_iteration": 1e6 * dt / max(rep.iters, 1)}
    kkt = fq.HodgeBlocks.compute(wc.mesh, 1).mixed_hodge_laplacian(symmetrized=True)
    bk = fq.DeviceVector.from_numpy(ctx, ((7 * np.arange(kkt.shape[0])) % 13 - 6).astype(np.float64))
    for mode in ("host", "device"):
        if mode == "host":
            os.environ["FQ_KRYLOV_HOST"] = "1"
        else:
            os.environ.pop("FQ_KRYLOV_HOST", None)
        fq.minres(kkt, None, bk, fq.StopCriterion(1e-10, 20))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, rep = fq.minres(kkt, None, bk, fq.StopCriterion(1e-30, 3000))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        row["minres_kkt_" + mode] = {"iters": rep.iters, "seconds": dt, "us_per_iteration": 1e6 * dt / max(rep.iters, 1)}
    out.append(row)
    print(json.dumps(row), flush=True)
