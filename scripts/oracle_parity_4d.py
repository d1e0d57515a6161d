"""One-off evidence: BASELINE config 3 (4-D k = 2) against the oracle beyond the test size: pattern bit-exact on the jittered
grid, values to 1e-12 (the 4 x 4 inverse of the third-party dependency is restated)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import formoniq_b200 as fq
from oracle import oracle as O
from tests.util import kuhn_problem

ctx = fq.Context(0)
for N in [int(a) for a in sys.argv[1:]] or [8]:
    t0 = time.perf_counter()
    shape = [N] * 4
    cx, s, *_ = kuhn_problem(4, shape, jitter=True)
    mesh = fq.Mesh.kuhn(ctx, 4, shape, jitter=0.2)
    lengths_equal = bool(np.array_equal(mesh.lengths(), s))
    hb = fq.HodgeBlocks.symbolic(mesh, 2)
    hb.numeric(mesh)
    hb.numeric(mesh)
    pattern, worst, nnz = True, 0.0, 0
    for blk, (kind, g) in zip(hb.blocks, [(O.MASS, 1), (O.MASS, 2), (O.DIF_TEST, 2), (O.DIF_BOTH, 3)]):
        ref = cx.assemble(s, kind, g, nthreads=O.max_threads())
        rp, ci, va = blk.download()
        erp, eci, eva = ref.arrays()
        pattern = pattern and np.array_equal(rp.astype(np.int64), erp) and np.array_equal(ci.astype(np.int64), eci)
        if pattern:
            worst = max(worst, float(np.abs(va - eva).max() / np.abs(eva).max()))
        nnz += len(eva)
    print(json.dumps({"config": 3, "grid": f"{N}^4", "pentatopes": 24 * N ** 4, "nnz": nnz, "edge_lengths_bitwise": lengths_equal,
                      "pattern_bit_exact": bool(pattern), "max_relative_value_difference": worst,
                      "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
