"""Where the end-to-end (host buffers in, host CSR out) time goes.  python scripts/e2e_phases.py [--n 128]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import formoniq_b200 as fq

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
args = ap.parse_args()
ctx = fq.Context(0, stream=torch.cuda.current_stream().cuda_stream)
DIM = 3
shape = [args.n] * 3
gen = fq.Mesh.kuhn(ctx, DIM, shape)
lengths = gen.lengths()
ns = fq.kuhn_counts(DIM, shape)
del gen

def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t

faces_t = [pinned(fq.kuhn_cell_faces_host(DIM, shape, j).view(np.int64)) for j in range(DIM)] + [None]
faces = [None if t is None else t.numpy().view(np.uint64) for t in faces_t]
len_t = pinned(lengths)
W = fq.WhitneyPairing
forms = [W.mass(DIM, 0), W.mass(DIM, 1), W.dif_test(DIM, 1), W.dif_both(DIM, 2)]

def T():
    torch.cuda.synchronize()
    return time.perf_counter()

ctx.set_timing(True)
for it in range(2):
    ctx.timing_report()
    t0 = T()
    m = fq.Mesh.from_arrays(ctx, DIM, ns, faces, len_t.numpy())
    t1 = T()
    print(f"iter {it}: mesh_create {1e3 * (t1 - t0):8.1f} ms")
    tot_sym = tot_num = tot_dl = 0.0
    for f in forms:
        a0 = T()
        a = f.symbolic(m)
        a1 = T()
        a.numeric(m, True)
        a2 = T()
        rp, ci, va = a.download()
        a3 = T()
        tot_sym += a1 - a0; tot_num += a2 - a1; tot_dl += a3 - a2
        print(f"   block: symbolic {1e3 * (a1 - a0):8.1f}  numeric {1e3 * (a2 - a1):8.1f}  download {1e3 * (a3 - a2):8.1f} ms  nnz {a.nnz}")
        del a, rp, ci, va
    print(f"   total: symbolic {1e3 * tot_sym:.1f}  numeric {1e3 * tot_num:.1f}  download {1e3 * tot_dl:.1f}  all {1e3 * (T() - t0):.1f} ms")
    print('   spans:', {k: round(v['ms'], 1) for k, v in ctx.timing_report().items()})
    del m
