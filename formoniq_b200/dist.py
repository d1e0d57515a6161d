"""Owner-computes row partition over ranks (one process per GPU) and the SpMV
halo exchange — the only inter-GPU traffic on the path (north star).

The Kuhn grid is cut into slabs of box layers along the last axis.  Colex
numbering sorts simplices by their top vertex and vertices are numbered with
the last axis slowest, so every rank owns a contiguous range of rows of every
grade and needs, for y = A x, the x entries of a contiguous window
[held_lo, held_hi) = lower halo | owned | upper halo.  The halos are owned by
the two neighbouring ranks.  The exchange is one send/recv pair per neighbour
(torch.distributed; NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from dataclasses import dataclass

from .api import kuhn_slab_ranges


def slab_of(rank: int, world: int, nlayers: int):
    """Box layers [begin, end) owned by `rank` (balanced contiguous split)."""
    base, rem = divmod(nlayers, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@dataclass
class SlabRanges:
    held_lo: int
    own_lo: int
    own_hi: int
    held_hi: int


class SlabPartition:
    """Id ranges of every rank for one grade (host-only, closed form)."""

    def __init__(self, dim: int, shape, world: int, grade: int):
        self.dim, self.shape, self.world, self.grade = dim, list(shape), world, grade
        nl = self.shape[dim - 1]
        if world > nl:
            raise ValueError("more ranks than box layers")
        self.slabs = [slab_of(r, world, nl) for r in range(world)]
        self.ranges = [SlabRanges(*kuhn_slab_ranges(dim, self.shape, s, grade)) for s in self.slabs]
        for r in range(world - 1):  # owned ranges tile the id space
            assert self.ranges[r].own_hi == self.ranges[r + 1].own_lo
        assert self.ranges[0].own_lo == 0

    def sends(self, rank: int):
        """[(peer, lo, hi)]: global id segments of my owned range a peer needs as halo."""
        me, out = self.ranges[rank], []
        for peer in (rank - 1, rank + 1):
            if 0 <= peer < self.world:
                pr = self.ranges[peer]
                for lo, hi in ((pr.held_lo, pr.own_lo), (pr.own_hi, pr.held_hi)):
                    a, b = max(lo, me.own_lo), min(hi, me.own_hi)
                    if a < b:
                        out.append((peer, a, b))
        return out

    def recvs(self, rank: int):
        """[(peer, lo, hi)]: halo segments of my window owned by a peer."""
        me, out = self.ranges[rank], []
        for peer in (rank - 1, rank + 1):
            if 0 <= peer < self.world:
                pr = self.ranges[peer]
                for lo, hi in ((me.held_lo, me.own_lo), (me.own_hi, me.held_hi)):
                    a, b = max(lo, pr.own_lo), min(hi, pr.own_hi)
                    if a < b:
                        out.append((peer, a, b))
        return out


def exchange_halo(window, part: SlabPartition, rank: int, group=None):
    """Fill the halo parts of `window` (a 1-D torch tensor covering
    [held_lo, held_hi) of this rank) from the neighbours' owned values."""
    import torch.distributed as dist

    if part.world == 1:
        return
    base = part.ranges[rank].held_lo
    ops = []
    for peer, lo, hi in part.sends(rank):
        ops.append(dist.P2POp(dist.isend, window[lo - base:hi - base], peer, group=group))
    for peer, lo, hi in part.recvs(rank):
        ops.append(dist.P2POp(dist.irecv, window[lo - base:hi - base], peer, group=group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
