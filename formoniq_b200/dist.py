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

import numpy as np

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


@dataclass
class MeshPart:
    """One rank's share of an uploaded mesh under owner-computes: the vertex range it owns, the cells it holds (every cell
    touching an owned vertex, ascending global index) and the id range [lo, hi) it owns of every grade."""
    rank: int
    vertices: tuple
    cells: np.ndarray
    own: list


def partition_mesh(dim: int, nsimplices, cell_faces, world: int):
    """Owner-computes partition of ANY mesh in the reference's skeleton numbering, not only generated Kuhn grids.

    The reference numbers the simplices of every grade in colex order of their sorted vertex lists
    (simplicial/src/topology/skeleton.rs:50-86), i.e. by top vertex first: a contiguous vertex range therefore owns a
    contiguous id range of every grade, and every cell that contributes to an owned row contains that row's top vertex,
    which is owned — so a rank that holds the cells touching its vertices assembles its row block without communication,
    in the same cell order as one GPU (bit-identical rows).  Vertex ranges are balanced by the number of incident cells.
    cell_faces[j]: [ncells, C(dim+1, j+1)] global tables (FaceIncidence::faces_flat), grade 0 required; a grade whose
    table is missing gets the range (0, 0).  Returns one MeshPart per rank."""
    from math import comb
    from itertools import combinations

    ns = [int(v) for v in nsimplices]
    cells = np.asarray(cell_faces[0]).astype(np.int64).reshape(-1, dim + 1)
    if np.any(np.diff(cells, axis=1) <= 0):
        raise ValueError("cells must list their vertices in ascending order")
    nv = ns[0]
    weight = np.bincount(cells.ravel(), minlength=nv).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(weight)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, nv
    cuts = list(np.maximum.accumulate(cuts))
    # top vertex of every simplex of every grade, from the cells (colex subsets: the last position is the top one)
    tops = []
    for j in range(dim + 1):
        f = cell_faces[j] if j < len(cell_faces) else None
        if j == dim and f is None:
            tops.append(cells[:, dim].copy())
            continue
        if f is None:
            tops.append(None)
            continue
        f = np.asarray(f).astype(np.int64).reshape(-1, comb(dim + 1, j + 1))
        subsets = sorted(combinations(range(dim + 1), j + 1), key=lambda c: c[::-1])  # colex order of the local faces
        top = np.full(ns[j], -1, dtype=np.int64)
        for r, sub in enumerate(subsets):
            top[f[:, r]] = cells[:, sub[-1]]
        if np.any(top < 0) or np.any(np.diff(top) < 0):
            raise ValueError(f"grade {j}: simplices are not numbered by top vertex (not the reference's skeleton numbering)")
        tops.append(top)
    parts = []
    for r in range(world):
        v0, v1 = cuts[r], cuts[r + 1]
        held = np.nonzero(((cells >= v0) & (cells < v1)).any(axis=1))[0]
        own = []
        for j in range(dim + 1):
            if tops[j] is None:
                own.append((0, 0))
            else:
                own.append((int(np.searchsorted(tops[j], v0, side="left")), int(np.searchsorted(tops[j], v1, side="left"))))
        parts.append(MeshPart(r, (v0, v1), held, own))
    return parts


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


def all_reduce_scalar(value: float, op: str = "sum", group=None, world: int | None = None) -> float:
    """One scalar across the ranks (the only collective of the Krylov solvers besides the halo exchange).  `world` = 1
    marks a single-rank object inside a multi-rank job (no collective)."""
    import torch
    import torch.distributed as dist

    if world == 1 or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, group=group)
    return float(t.item())


class DistKktPencil:
    """The pencil of `elliptic::solve_evp` (problems/elliptic.rs:225-247) row-partitioned over the ranks:

        A = [[M_{k-1}, -dif_test(k)], [dif_test(k)^T, dif_both(k+1)]]   (hodge.rs:93-99),   B = diag(0, M_k).

    Every rank owns the sigma rows and the u rows of its z-slab (owner computes: its blocks are assembled from its own
    cells plus one halo layer, no assembly traffic).  The lower-left block is assembled as dif_trial(k) — the same
    operator as dif_test(k)^T with rows = the rank's u rows, so no transpose crosses ranks.  A vector is the pair
    (sigma_owned | u_owned); an operator application copies the two owned segments into column windows
    [held_lo, held_hi), refreshes their halos with one send/recv pair per neighbour and runs the four windowed SpMVs;
    inner products are completed by an all-reduce of one scalar.  Implements the pencil protocol of
    eigen.shift_invert_lanczos; the inner solves of the shift-invert step are MINRES on A - shift*B (SpMV-only).

    precond = "afw": MINRES is preconditioned by the AFW block diagonal diag(hdif_gram(k-1)^-1, hdif_gram(k)^-1)
    (problems/elliptic.rs:29-47) across the ranks: hdif_gram(j) = M_j + dif_both(j+1) (whitney_complex.rs:118-125) is
    assembled on the rank's rows, and each block solve is a Jacobi-preconditioned CG on the row-partitioned operator
    (halo exchange + windowed SpMV per iteration, all-reduced inner products) to `afw_rtol`, where the reference applies
    a sparse Cholesky factor.  The outer iteration count is then independent of the mesh width."""

    def __init__(self, ctx, dim: int, shape, grade: int, rank: int = 0, world: int = 1, group=None, jitter: float = 0.0,
                 inner_rtol: float = 1e-13, inner_max_iters: int = 200000, precond: str = "none", afw_rtol: float = 1e-12,
                 afw_max_iters: int = 20000):
        import torch

        from .api import DeviceVector, HodgeBlocks, Mesh, WhitneyPairing

        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.inner_rtol, self.inner_max_iters, self.inner_iterations, self.applies = inner_rtol, inner_max_iters, 0, 0
        shape = list(shape)
        slab = slab_of(rank, world, shape[dim - 1])
        self.mesh = Mesh.kuhn(ctx, dim, shape, slab=slab, jitter=jitter)
        self.part_s = SlabPartition(dim, shape, world, grade - 1)
        self.part_u = SlabPartition(dim, shape, world, grade)
        self.rs, self.ru = self.part_s.ranges[rank], self.part_u.ranges[rank]
        srows, urows = (self.rs.own_lo, self.rs.own_hi), (self.ru.own_lo, self.ru.own_hi)
        self.hb = HodgeBlocks.symbolic(self.mesh, grade, srows, urows)
        self.hb.numeric(self.mesh, True)
        self.dif_trial = WhitneyPairing.dif_trial(dim, grade).symbolic(self.mesh, *urows)
        self.dif_trial.numeric(self.mesh, True)
        self.ns, self.nu = srows[1] - srows[0], urows[1] - urows[0]
        self.n = self.ns + self.nu
        self.ns_global, self.nu_global = self.mesh.nsimplices(grade - 1), self.mesh.nsimplices(grade)
        self.n_global = self.ns_global + self.nu_global
        # column windows (torch owns the memory so that torch.distributed can send / receive slices of it)
        self.win_s_t = torch.zeros(self.rs.held_hi - self.rs.held_lo, dtype=torch.float64, device="cuda")
        self.win_u_t = torch.zeros(self.ru.held_hi - self.ru.held_lo, dtype=torch.float64, device="cuda")
        self.win_s, self.win_u = DeviceVector.from_torch(ctx, self.win_s_t), DeviceVector.from_torch(ctx, self.win_u_t)
        self.own_s = self.win_s.view(self.rs.own_lo - self.rs.held_lo, self.ns)
        self.own_u = self.win_u.view(self.ru.own_lo - self.ru.held_lo, self.nu)
        self.tmp_s, self.tmp_u = DeviceVector(ctx, max(self.ns, 1)).view(0, self.ns), DeviceVector(ctx, max(self.nu, 1)).view(0, self.nu)
        self._shift = 0.0
        self.precond, self.afw_rtol, self.afw_max_iters, self.afw_iterations = precond, afw_rtol, afw_max_iters, 0
        if precond == "afw":
            stiff_s = WhitneyPairing.dif_both(dim, grade).symbolic(self.mesh, *srows)   # d^T M_k d on the sigma rows
            stiff_s.numeric(self.mesh, True)
            self.h_s = self.hb.mass_sigma + stiff_s          # hdif_gram(k-1), rows = the rank's sigma rows
            self.h_u = self.hb.mass_u + self.hb.dif_both     # hdif_gram(k),   rows = the rank's u rows
            self.jac_s = self.h_s.inv_diagonal() if self.ns else None
            self.jac_u = self.h_u.inv_diagonal() if self.nu else None
        elif precond != "none":
            raise ValueError("precond must be 'none' or 'afw'")
        # inf-norms of the block rows (linalg/eigen.rs:359-368)
        hb = self.hb
        rows_s = hb.mass_sigma.row_abs_sums().to_numpy() + hb.dif_test.row_abs_sums().to_numpy()
        rows_u = self.dif_trial.row_abs_sums().to_numpy() + hb.dif_both.row_abs_sums().to_numpy()
        a_local = max(float(rows_s.max()) if rows_s.size else 0.0, float(rows_u.max()) if rows_u.size else 0.0)
        b_rows = hb.mass_u.row_abs_sums().to_numpy()
        self.a_norm = all_reduce_scalar(a_local, "max", group, world)
        self.b_norm = all_reduce_scalar(float(b_rows.max()) if b_rows.size else 0.0, "max", group, world)

    # -- windows
    def _load(self, x, sigma: bool = True, u: bool = True):
        """Owned segments of x into the column windows, halos refreshed from the neighbours."""
        if sigma:
            self.own_s.copy_from(x.view(0, self.ns))
            exchange_halo(self.win_s_t, self.part_s, self.rank, self.group)
        if u:
            self.own_u.copy_from(x.view(self.ns, self.nu))
            exchange_halo(self.win_u_t, self.part_u, self.rank, self.group)

    # -- pencil protocol
    def a_apply(self, x, y=None, shift: float = 0.0, symmetrized: bool = False):
        """y = A x - shift * B x on the rank's rows; symmetrized: the sigma rows negated (the form MINRES needs,
        problems/elliptic.rs:101-113)."""
        hb = self.hb
        y = x.zeros_like() if y is None else y
        self._load(x)
        ys, yu = y.view(0, self.ns), y.view(self.ns, self.nu)
        hb.mass_sigma.apply_window(self.win_s, self.rs.held_lo, ys)
        hb.dif_test.apply_window(self.win_u, self.ru.held_lo, self.tmp_s)
        ys.add_scaled(-1.0, self.tmp_s)
        if symmetrized:
            ys.scale(-1.0)
        self.dif_trial.apply_window(self.win_s, self.rs.held_lo, yu)
        hb.dif_both.apply_window(self.win_u, self.ru.held_lo, self.tmp_u)
        yu.add(self.tmp_u)
        if shift != 0.0:
            hb.mass_u.apply_window(self.win_u, self.ru.held_lo, self.tmp_u)
            yu.add_scaled(-shift, self.tmp_u)
        self.applies += 1
        return y

    def b_apply(self, x, y=None):
        y = x.zeros_like() if y is None else y
        self._load(x, sigma=False)
        y.view(0, self.ns).fill_zero()
        self.hb.mass_u.apply_window(self.win_u, self.ru.held_lo, y.view(self.ns, self.nu))
        return y

    def dot(self, x, y) -> float:
        return all_reduce_scalar(x.dot(y), "sum", self.group, self.world)

    def seed(self, s: int):
        import numpy as np

        from .api import DeviceVector
        from .eigen import pseudo_random

        v = np.concatenate([pseudo_random(s, self.ns, self.rs.own_lo),
                            pseudo_random(s, self.nu, self.ns_global + self.ru.own_lo)])
        return DeviceVector.from_numpy(self.ctx, v)

    def prepare(self, shift: float) -> float:
        self._shift = shift
        return shift

    def _reduce(self):
        return (lambda local: all_reduce_scalar(local, "sum", self.group, self.world)) if self.world > 1 else None

    def _block_solve(self, h, jac, n, window_t, window, part, ranges, own, r, z):
        """z = h^-1 r on one block of the AFW preconditioner: Jacobi-CG on the row-partitioned hdif_gram."""
        from .api import StopCriterion, cg_op

        def apply(x, y):
            own.copy_from(x)
            exchange_halo(window_t, part, self.rank, self.group)
            h.apply_window(window, ranges.held_lo, y)

        sol, rep = cg_op(self.ctx, n, apply, r, StopCriterion(self.afw_rtol, self.afw_max_iters),
                         precond=(lambda rr, zz: zz.assign_product(jac, rr)) if jac is not None else None, reduce=self._reduce())
        self.afw_iterations += rep.iters
        z.copy_from(sol)

    def afw_apply(self, r, z):
        """z = diag(hdif_gram(k-1)^-1, hdif_gram(k)^-1) r (problems/elliptic.rs:29-47)."""
        self._block_solve(self.h_s, self.jac_s, self.ns, self.win_s_t, self.win_s, self.part_s, self.rs, self.own_s,
                          r.view(0, self.ns), z.view(0, self.ns))
        self._block_solve(self.h_u, self.jac_u, self.nu, self.win_u_t, self.win_u, self.part_u, self.ru, self.own_u,
                          r.view(self.ns, self.nu), z.view(self.ns, self.nu))

    def solve(self, v):
        from .api import StopCriterion, minres_op
        from .eigen import EigenError

        rhs = v.clone()                      # S (A - shift B) x = S v with S = diag(-1 on sigma, +1 on u)
        rhs.view(0, self.ns).scale(-1.0)
        x, rep = minres_op(self.ctx, self.n, lambda xin, y: self.a_apply(xin, y, self._shift, True), rhs,
                           StopCriterion(self.inner_rtol, self.inner_max_iters),
                           precond=self.afw_apply if self.precond == "afw" else None, reduce=self._reduce())
        self.inner_iterations += rep.iters
        if not rep.converged:
            raise EigenError("SingularPencil", shift=self._shift, inner_residual=rep.residual)
        return x

    def gather(self, x):
        """The global vector (sigma | u) on every rank, for tests."""
        import numpy as np
        import torch
        import torch.distributed as dist

        mine = x.to_numpy()
        if self.world == 1:
            return mine
        parts = [None] * self.world
        dist.all_gather_object(parts, (self.rs.own_lo, self.ru.own_lo, mine[:self.ns], mine[self.ns:]), group=self.group)
        out = np.zeros(self.n_global)
        for slo, ulo, s, u in parts:
            out[slo:slo + len(s)] = s
            out[self.ns_global + ulo:self.ns_global + ulo + len(u)] = u
        return out
