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


class PeerHalo:
    """SpMV fused with the halo exchange over peer memory (NVLink 5 / NVSwitch).

    Every rank keeps x as a window vector [held_lo, held_hi) allocated by the library; the windows and two
    epoch flags per rank are mapped into the neighbouring processes with CUDA IPC.  `apply` runs ONE kernel:
    its first CTAs wait for the neighbours' published epoch, pull the halo segments out of their windows (P2P
    loads over NVLink) and signal `consumed`, while all other CTAs already multiply the row blocks that only need
    owned columns; the boundary row blocks wait on a device counter.  No separate exchange or flag launches.

        ph = PeerHalo(ctx, part, rank, x_window)        # collective: exchanges the IPC handles
        ph.publish()                                     # after this rank finished writing its owned x
        y = ph.apply(a, y)                               # waits for the neighbours' epoch, then the fused SpMV
        ph.release()                                     # before x is overwritten again: neighbours are done reading
    """

    def __init__(self, ctx, part: SlabPartition, rank: int, x_window, group=None, fused_flags=None):
        import ctypes as C
        import os

        # the kernel handles the epoch flags itself unless the direct-gather variant is selected
        self.fused_flags = (os.environ.get("FQ_PEER_DIRECT", "0") in ("", "0")) if fused_flags is None else fused_flags

        import torch.distributed as dist

        from . import _lib
        from .api import DeviceVector

        self.ctx, self.part, self.rank, self.x, self.group = ctx, part, rank, x_window, group
        self.epoch = 0
        self.ready = DeviceVector(ctx, 1)     # epoch of the last published x
        self.consumed = DeviceVector(ctx, 1)  # epoch this rank has finished reading from its neighbours
        L = _lib.lib()

        def export(v):
            buf = (C.c_ubyte * 64)()
            _lib.check(L.fq_vec_ipc_export(ctx._h, v._h, C.cast(buf, C.c_void_p)))
            return bytes(buf)

        mine = {"x": export(x_window), "ready": export(self.ready), "consumed": export(self.consumed), "n": len(x_window)}
        world = part.world
        table = [None] * world
        if world > 1:
            dist.all_gather_object(table, mine, group=group)
        else:
            table[0] = mine

        def imp(handle, n):
            h = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(handle)
            _lib.check(L.fq_vec_ipc_import(ctx._h, C.cast(buf, C.c_void_p), n, C.byref(h)))
            v = DeviceVector.__new__(DeviceVector)
            v.ctx, v._h, v.n = ctx, h, n
            return v

        self.peers = {}
        for peer in (rank - 1, rank + 1):
            if 0 <= peer < world:
                t = table[peer]
                self.peers[peer] = {"x": imp(t["x"], t["n"]), "ready": imp(t["ready"], 1), "consumed": imp(t["consumed"], 1)}
        if world > 1:
            dist.barrier(group=group)

    def publish(self):
        """This rank's owned x entries are final for the next epoch (stream-ordered)."""
        from . import _lib

        self.epoch += 1
        _lib.check(_lib.lib().fq_flag_signal(self.ctx._h, self.ready._h, float(self.epoch)))

    def apply(self, a, y):
        """y = A x with the neighbours' columns read over NVLink inside the SpMV kernel."""
        from . import _lib

        L = _lib.lib()
        r = self.part.ranges[self.rank]
        lo, hi = self.peers.get(self.rank - 1), self.peers.get(self.rank + 1)
        lo_base = self.part.ranges[self.rank - 1].held_lo if lo else 0
        hi_base = self.part.ranges[self.rank + 1].held_lo if hi else 0
        if self.fused_flags:
            _lib.check(L.fq_spmv_peer_epoch(self.ctx._h, a._h, self.x._h, r.held_lo, r.own_lo, r.own_hi,
                                            lo["x"]._h if lo else None, lo_base, lo["ready"]._h if lo else None,
                                            hi["x"]._h if hi else None, hi_base, hi["ready"]._h if hi else None,
                                            float(self.epoch), self.consumed._h, y._h))
            return y
        for p in self.peers.values():
            _lib.check(L.fq_flag_wait(self.ctx._h, p["ready"]._h, float(self.epoch)))
        _lib.check(L.fq_spmv_peer(self.ctx._h, a._h, self.x._h, r.held_lo, r.own_lo, r.own_hi,
                                  lo["x"]._h if lo else None, lo_base, hi["x"]._h if hi else None, hi_base, y._h))
        _lib.check(L.fq_flag_signal(self.ctx._h, self.consumed._h, float(self.epoch)))
        return y

    def release(self):
        """Wait until the neighbours have finished reading this rank's x of the current epoch."""
        from . import _lib

        for p in self.peers.values():
            _lib.check(_lib.lib().fq_flag_wait(self.ctx._h, p["consumed"]._h, float(self.epoch)))

    def check(self):
        from . import _lib

        _lib.check(_lib.lib().fq_flag_check(self.ctx._h))
