"""Debug aid (used to find the b_empty phase-mixing race): tile pass vs slab pass on a mid-size jittered mesh, mismatch statistics per block."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import formoniq_b200 as fq

ctx = fq.Context(0)
shape = [22, 19, 25]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    mesh = fq.Mesh.kuhn(ctx, 3, shape, jitter=0.2)
    hb = fq.HodgeBlocks.symbolic(mesh, 1)
    hb.numeric(mesh)
    first = [blk.download() for blk in hb.blocks]
    for it in range(3):
        hb.numeric(mesh)
        for b, (blk, (rp0, ci0, va0)) in enumerate(zip(hb.blocks, first)):
            rp, ci, va = blk.download()
            bad = np.flatnonzero(va.view(np.uint64) != va0.view(np.uint64))
            if bad.size:
                rel = np.abs(va[bad] - va0[bad]) / np.maximum(np.abs(va0[bad]), 1e-300)
                rows = np.searchsorted(rp, bad, side="right") - 1
                print(f"rep {rep} pass {it} block {b}: {bad.size} of {va.size} differ, max rel {rel.max():.3e}, first idx {bad[:8]}, rows {rows[:8]}, "
                      f"vals {va[bad[:3]]} vs {va0[bad[:3]]}")
    print(f"rep {rep} done", flush=True)
