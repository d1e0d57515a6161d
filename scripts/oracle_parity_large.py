"""One-off evidence: the fused assembly against the oracle on large Kuhn cubes (pattern bit for bit, values bitwise).
usage: python scripts/oracle_parity_large.py N [N ...]   — prints one JSON line per N."""
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
for N in [int(a) for a in sys.argv[1:]]:
    for variant in os.environ.get("FQ_PARITY_VARIANTS", "plain,jitter").split(","):
        t0 = time.perf_counter()
        cx, s, *_ = kuhn_problem(3, [N, N, N], jitter=variant == "jitter")
        mesh = fq.Mesh.kuhn(ctx, 3, [N, N, N], jitter=0.2 if variant == "jitter" else 0.0)
        lengths_equal = bool(np.array_equal(mesh.lengths(), s))
        hb = fq.HodgeBlocks.symbolic(mesh, 1)
        ctx.set_timing(True)
        ctx.timing_report()
        hb.numeric(mesh)
        hb.numeric(mesh)
        fused = ctx.timing_report().get("k13_tile_fused", {}).get("count", 0)
        ctx.set_timing(False)
        ok, nnz = True, 0
        for blk, (kind, g) in zip(hb.blocks, [(O.MASS, 0), (O.MASS, 1), (O.DIF_TEST, 1), (O.DIF_BOTH, 2)]):
            ref = cx.assemble(s, kind, g, nthreads=O.max_threads())
            rp, ci, va = blk.download()
            erp, eci, eva = ref.arrays()
            ok = ok and np.array_equal(rp.astype(np.int64), erp) and np.array_equal(ci.astype(np.int64), eci)
            ok = ok and same_bits_mod_zero_sign(va, eva)
            nnz += len(eva)
            del ref, rp, ci, va, erp, eci, eva
        print(json.dumps({"N": N, "variant": variant, "tets": 6 * N ** 3, "nnz": nnz, "edge_lengths_bitwise": lengths_equal,
                          "fused_kernel_launches": fused, "pattern_bit_exact_and_values_bitwise": bool(ok),
                          "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
        del hb, mesh, cx, s
        fq._lib.lib().fq_device_cache_trim()
