"""2+ GPU check of the SpMV fused with the halo exchange over peer memory (run under torchrun):
the fused kernel must reproduce exchange + windowed SpMV bit for bit; prints both timings on rank 0.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/peer_spmv_check.py --size 64"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import formoniq_b200 as fq
from formoniq_b200.dist import PeerHalo, SlabPartition, exchange_halo, slab_of

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.current_stream()
ctx = fq.Context(local, stream=stream.cuda_stream)
DIM = 3
shape = [args.size, args.size, args.size * world]
mesh = fq.Mesh.kuhn(ctx, DIM, shape, slab=slab_of(rank, world, shape[2]))
ok = True
for name, form in (("mass_u", fq.WhitneyPairing.mass(DIM, 1)), ("dif_test", fq.WhitneyPairing.dif_test(DIM, 1)),
                   ("mass_sigma", fq.WhitneyPairing.mass(DIM, 0))):
    b, e = mesh.owned_range(form.test_grade())
    a = form.symbolic(mesh, b, e)
    a.numeric(mesh)
    part = SlabPartition(DIM, shape, world, form.trial_grade())
    r = part.ranges[rank]
    # owned entries only; halos poisoned: the fused kernel must never read them from the local window
    xw = torch.full((r.held_hi - r.held_lo,), float("nan"), device="cuda", dtype=torch.float64)
    ids = torch.arange(r.own_lo, r.own_hi, device="cuda", dtype=torch.float64)
    xw[r.own_lo - r.held_lo:r.own_hi - r.held_lo] = torch.cos(ids * ids + 1.0)
    xv = fq.DeviceVector(ctx, r.held_hi - r.held_lo)          # library-allocated: IPC-exportable
    xw_view = fq.DeviceVector.from_torch(ctx, xw)
    fq._lib.check(fq._lib.lib().fq_vec_copy(ctx._h, xv._h, xw_view._h))
    y_peer = fq.DeviceVector(ctx, e - b)
    ph = PeerHalo(ctx, part, rank, xv)
    ph.publish()
    ph.apply(a, y_peer)
    ph.release()
    ph.check()
    # reference: NCCL exchange into the window, then the windowed SpMV
    exchange_halo(xw, part, rank)
    y_ref = a.apply_window(fq.DeviceVector.from_torch(ctx, xw), r.held_lo)
    same = bool((torch.from_numpy(y_peer.to_numpy()) == torch.from_numpy(y_ref.to_numpy())).all())
    ok = ok and same

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            fn()
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    xin = fq.DeviceVector.from_torch(ctx, xw)
    y2 = fq.DeviceVector(ctx, e - b)

    def nccl_path():
        exchange_halo(xw, part, rank)
        a.apply_window(xin, r.held_lo, y2)

    def peer_path():
        ph.publish()
        ph.apply(a, y_peer)
        ph.release()

    t_nccl, t_peer = timed(nccl_path), timed(peer_path)
    ph.check()
    # break-down: the local kernels alone (no exchange, no flags) on the torch window and on the IPC-exported window
    t_local = timed(lambda: a.apply_window(xin, r.held_lo, y2))
    t_local_ipc = timed(lambda: a.apply_window(xv, r.held_lo, y2))

    def peer_kernel_only():
        lo, hi = ph.peers.get(rank - 1), ph.peers.get(rank + 1)
        fq._lib.check(fq._lib.lib().fq_spmv_peer(ctx._h, a._h, xv._h, r.held_lo, r.own_lo, r.own_hi,
                                                 lo["x"]._h if lo else None, part.ranges[rank - 1].held_lo if lo else 0,
                                                 hi["x"]._h if hi else None, part.ranges[rank + 1].held_lo if hi else 0, y_peer._h))

    t_kernel = timed(peer_kernel_only)
    if rank == 0:
        print(f"   spmv local {t_local:.4f}  local on ipc window {t_local_ipc:.4f}  fused kernel without flags {t_kernel:.4f} ms", flush=True)
    if rank == 0:
        print(f"{name}: bitwise_equal={same}  nccl exchange + spmv {t_nccl:.4f} ms   fused peer spmv {t_peer:.4f} ms  (nnz/rank {a.nnz})", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_SPMV_OK" if flag.item() == 1.0 else "PEER_SPMV_MISMATCH", flush=True)
dist.destroy_process_group()
