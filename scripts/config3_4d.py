"""BASELINE config 3: 4-D Hodge-Laplace k = 2 on a Kuhn hypercube (arbitrary-dimension path).  python scripts/config3_4d.py [--grid 16]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import formoniq_b200 as fq

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
ctx = fq.Context(0, stream=torch.cuda.current_stream().cuda_stream)
n = args.grid
mesh = fq.Mesh.kuhn(ctx, 4, [n] * 4)
ctx.set_timing(True)
torch.cuda.synchronize(); t0 = time.perf_counter()
hb = fq.HodgeBlocks.symbolic(mesh, 2)
hb.numeric(mesh, True)
torch.cuda.synchronize(); t_first = time.perf_counter() - t0
first = ctx.timing_report()
for _ in range(2):
    hb.numeric(mesh, True)
ctx.timing_report()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(args.steps):
    hb.numeric(mesh, True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / args.steps
rep = ctx.timing_report()
nnz = sum(b.nnz for b in hb.blocks)
abytes = max(b.assembly_shared_bytes for b in hb.blocks) + sum(b.assembly_bytes - b.assembly_shared_bytes for b in hb.blocks)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
print(json.dumps({"workload": f"4-D Hodge-Laplace k=2, Kuhn grid {n}^4 ({mesh.ncells} pentatopes), four blocks (10x10 element matrices)",
                  "cells": mesh.ncells, "nnz": nnz, "ms_per_step": 1e3 * dt, "elements_per_s": mesh.ncells / dt, "nnz_per_s": nnz / dt,
                  "first_pass_ms": 1e3 * t_first, "first_pass_device_ms": {k: round(v["ms"], 2) for k, v in first.items()},
                  "kernels_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in rep.items()},
                  "algorithmic_bytes": abytes, "roofline": {"bound": "hbm", "achieved": abytes / 1e9 / dt, "peak": peak, "unit": "GB/s",
                                                           "frac": abytes / 1e9 / dt / peak}}))
