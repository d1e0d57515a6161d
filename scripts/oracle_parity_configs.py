"""One-off evidence beyond the test sizes: BASELINE configs 1 and 4 against the oracle at scale, and the SpMV of the
headline matrix against the oracle's serial product (bitwise).  Prints one JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import formoniq_b200 as fq
from oracle import oracle as O
from tests.util import kuhn_problem, same_bits_mod_zero_sign

ctx = fq.Context(0)


def compare(blk, ref):
    rp, ci, va = blk.download()
    erp, eci, eva = ref.arrays()
    return bool(np.array_equal(rp.astype(np.int64), erp) and np.array_equal(ci.astype(np.int64), eci)
                and same_bits_mod_zero_sign(va, eva)), len(eva)


# config 1: 2-D source problem blocks (M0, M1, dif_test(1), dif_both(2)) on a large Kuhn square, plain and jittered
for N in (1024,):
    for variant in ("plain", "jitter"):
        t0 = time.perf_counter()
        cx, s, *_ = kuhn_problem(2, [N, N], jitter=variant == "jitter")
        mesh = fq.Mesh.kuhn(ctx, 2, [N, N], jitter=0.2 if variant == "jitter" else 0.0)
        hb = fq.HodgeBlocks.symbolic(mesh, 1)
        hb.numeric(mesh)
        hb.numeric(mesh)
        ok, nnz = True, 0
        for blk, (kind, g) in zip(hb.blocks, [(O.MASS, 0), (O.MASS, 1), (O.DIF_TEST, 1), (O.DIF_BOTH, 2)]):
            good, m = compare(blk, cx.assemble(s, kind, g, nthreads=O.max_threads()))
            ok, nnz = ok and good, nnz + m
        print(json.dumps({"config": 1, "grid": f"{N}x{N}", "variant": variant, "cells": 2 * N * N, "nnz": nnz,
                          "pattern_bit_exact_and_values_bitwise": ok, "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
        del hb, mesh, cx, s

# config 4: 2+1 Minkowski Hodge-Dirac, the masses M0..M3 on the Lorentzian Kuhn box (signed lengths, det g < 0)
for N in (48,):
    t0 = time.perf_counter()
    cx, s, coords, diag, vmax = kuhn_problem(3, [N, N, N], minkowski=True)
    mesh = fq.Mesh.kuhn(ctx, 3, [N, N, N], vmax=vmax, ambient_diag=diag)
    lengths_equal = bool(np.array_equal(mesh.lengths(), s))
    ok, nnz = True, 0
    for g in range(4):
        form = fq.WhitneyPairing.mass(3, g)
        a = form.assemble(mesh)
        a.numeric(mesh)  # second pass: the tile path where there is one
        good, m = compare(a, cx.assemble(s, O.MASS, g, nthreads=O.max_threads()))
        ok, nnz = ok and good, nnz + m
    print(json.dumps({"config": 4, "grid": f"{N}^3 Minkowski", "cells": 6 * N ** 3, "nnz": nnz, "negative_lengths": int((s < 0).sum()),
                      "edge_lengths_bitwise": lengths_equal, "pattern_bit_exact_and_values_bitwise": ok,
                      "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
    del mesh, cx, s

# SpMV of the headline M1 (N = 128, 241.5 M non-zeros) against the oracle's serial product
N = int(os.environ.get("FQ_SPMV_N", "128"))
t0 = time.perf_counter()
cx, s, *_ = kuhn_problem(3, [N, N, N])
mesh = fq.Mesh.kuhn(ctx, 3, [N, N, N])
a = fq.WhitneyPairing.mass(3, 1).assemble(mesh)
ref = cx.assemble(s, O.MASS, 1, nthreads=O.max_threads())
n = a.shape[0]
x = ((7 * np.arange(n)) % 13 - 6).astype(np.float64)  # matfree.rs:205-207
y = a.apply(fq.DeviceVector.from_numpy(ctx, x)).to_numpy()
print(json.dumps({"spmv": "M1", "N": N, "rows": n, "nnz": a.nnz, "y_bitwise_equal_to_the_serial_cpu_product": bool(np.array_equal(y, ref.spmv(x))),
                  "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
