"""CPU model of the record stream of one interior tile (no GPU): run lengths per (block, L), padding lanes, records, chunks,
and the first-half / second-half split the alternating kernel sees.  python scripts/tile_stream_model.py [bx by bz]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O

brick = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 3, 3)
CHUNK, CHDR, RHDR, NW = 1536, 16, 16, 16
shape = [brick[0] + 4, brick[1] + 4, brick[2] + 4]
cx = O.Complex.kuhn(3, shape)
nv = [n + 1 for n in shape]
vid = lambda x, y, z: x + nv[0] * (y + nv[1] * z)
owned = {vid(x, y, z) for x in range(2, 2 + brick[0]) for y in range(2, 2 + brick[1]) for z in range(2, 2 + brick[2])}
skel = [cx.skeleton(j) for j in range(4)]
blocks = [("M0", 0, 0), ("M1", 1, 1), ("dif_test", 0, 1), ("dif_both", 1, 1)]


def rec_lanes(L):
    m64 = ((CHUNK - 32) // 64 - 4) // 2
    m32 = ((CHUNK - 32) // 32 - 4) // 2
    return 64 if L <= m64 else (32 if L <= m32 else 16)


cells = {c for c in range(cx.ncells) if owned & set(skel[3][c].tolist())}
print(f"brick {brick}: {len(owned)} vertices, {len(cells)} cells ({len(cells) / len(owned):.2f} per vertex, recompute x{len(cells) / len(owned) / 6:.2f})")
tot_nnz = tot_lanes = tot_bytes = 0
off = 0
chunks_by_block = []
for name, tg, rg in blocks:
    ones = np.ones((cx.ncells, O.nlocal(3, tg), O.nlocal(3, rg)))
    m = O.assemble_from_elmats(cx, tg, rg, ones, drop_zeros=False)      # value = number of contributions L
    rows = [r for r in range(cx.nsimplices(tg)) if int(skel[tg][r].max()) in owned]   # top vertex = largest id (colex)
    Ls = np.concatenate([m.data[m.indptr[r]:m.indptr[r + 1]] for r in rows]).astype(int)
    start = off // CHUNK
    nnz = len(Ls)
    lanes = recs = 0
    for L in sorted(set(Ls.tolist())):
        n = int((Ls == L).sum())
        w = rec_lanes(L)
        k = -(-n // w)
        size = RHDR + w * (4 + 2 * L)
        for _ in range(k):
            if off % CHUNK == 0 or off + size > (off // CHUNK + 1) * CHUNK:
                if off % CHUNK:
                    off = (off // CHUNK + 1) * CHUNK
                off += CHDR
            off += size
        lanes += k * w
        recs += k
    end = -(-off // CHUNK)
    chunks_by_block.append((name, start, end))
    tot_nnz += nnz
    tot_lanes += lanes
    print(f"  {name:9s} rows {len(rows):4d} nnz {nnz:5d} contributions {int(Ls.sum()):6d} (L mean {Ls.mean():.2f}, max {Ls.max()}) "
          f"runs {len(set(Ls.tolist())):2d} records {recs:3d} lanes {lanes:5d} (padding {100 * (lanes - nnz) / lanes:.1f} %)")
nchunks = -(-off // CHUNK)
ideal = sum(1 for _ in range(1))
print(f"  total: nnz {tot_nnz}, lanes {tot_lanes} (padding {100 * (tot_lanes - tot_nnz) / tot_lanes:.1f} %), {nchunks} chunks of {CHUNK} B "
      f"= {nchunks * CHUNK} B ({nchunks * CHUNK / tot_nnz:.1f} B per non-zero), {nchunks / NW:.2f} chunks per warp "
      f"(quantisation: {100 * (1 - nchunks / (NW * -(-nchunks // NW))):.1f} % idle warp-chunks)")
y0 = chunks_by_block[-1][1]
print(f"  first-half chunks (M0, M1, dif_test): {y0}, second-half chunks (dif_both): {nchunks - y0}")
