#!/usr/bin/env python
"""bench.py — the north-star benchmark: Galerkin assembly of the four
Hodge-Laplace blocks (M_{k-1}, M_k, dif_test(k), dif_both(k+1), k = 1) on a
synthetic 3-D Kuhn mesh of ~10 M tets per GPU, plus the CSR SpMV, reported as
elements/s, nnz/s and GB/s next to the CPU path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one numeric assembly pass of the four blocks over the rank's
cells (reference `!= 0.0` pattern semantics) with mesh tables and the
cell-slot->nnz maps resident in HBM.  Under torchrun every rank owns a slab
of box layers of a grid that grows with N (weak scaling); there is no
collective on the assembly path, the SpMV does one halo exchange per apply.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "assembly_elements_per_s"
UNIT = "elements/s"
DIM, GRADE = 3, 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first(self, timeout: float = 5.0):
        """Block until nvidia-smi delivers its first sample (its start-up can exceed a short timed region)."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """Samples from here on belong to the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hodge_forms(fq):
    W = fq.WhitneyPairing
    return [("mass_sigma", W.mass(DIM, GRADE - 1)), ("mass_u", W.mass(DIM, GRADE)), ("dif_test", W.dif_test(DIM, GRADE)),
            ("dif_both", W.dif_both(DIM, GRADE + 1))]


# --------------------------------------------------------------------------- CPU arm
def cpu_assembly_sample(sample_n: int | None, budget_s: float = 20.0):
    """The reference's CPU algorithm (oracle port: per-cell element matrices on
    all host threads, ordered concat, `!= 0.0` filter, serial COO->CSR) on a
    bounded Kuhn cube; returns elements/s over the four blocks."""
    import numpy as np

    from oracle import oracle as O

    threads = O.max_threads()

    def run(n):
        cx = O.Complex.kuhn(DIM, n)
        s = cx.edge_lengths_sq(O.kuhn_vertex_coords(DIM, n))
        total, nnz, par, ser = 0.0, 0, 0.0, 0.0
        for kind, k in ((O.MASS, GRADE - 1), (O.MASS, GRADE), (O.DIF_TEST, GRADE), (O.DIF_BOTH, GRADE + 1)):
            tm = np.zeros(2)
            a = cx.assemble(s, kind, k, nthreads=threads, times=tm)
            total += float(tm.sum())
            par += float(tm[0])
            ser += float(tm[1])
            nnz += a.nnz
        return cx.ncells, nnz, total, par, ser

    if sample_n is None:
        cells, _, t, _, _ = run(12)
        rate = cells / t
        sample_n = int(max(12, min(64, (budget_s * rate / 6.0) ** (1.0 / 3.0))))
        sample_n -= sample_n % 4
    cells, nnz, t, t_par, t_ser = run(sample_n)
    return {"value": cells / t, "nnz_per_s": nnz / t, "cores": threads, "seconds": t,
            "parallel_elmat_s": t_par, "serial_coo_to_csr_s": t_ser,
            "sample": f"3-D Kuhn cube N={sample_n} ({cells} tets), four Hodge blocks k=1: element matrices on {threads} threads "
                      f"({t_par:.1f} s) + serial COO->CSR as in the reference ({t_ser:.1f} s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_assembly_sample(args.sample_n, budget_s=5.0)
        if i >= args.warmup:
            vals.append(last)
        if args.sample_n is None:  # keep the calibrated size for the remaining steps
            args.sample_n = int(last["sample"].split("N=")[1].split(" ")[0])
    value = sum(v["value"] for v in vals) / len(vals)
    ms = 1e3 * sum(v["seconds"] for v in vals) / len(vals)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3-D Hodge-Laplace k=1 mixed (AFW) assembly, Kuhn unit cube, four blocks; CPU sample of the "
                               f"N={args.n} per-GPU workload", "sample": last["sample"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"],
                         "parallel_elmat_s": last["parallel_elmat_s"], "serial_coo_to_csr_s": last["serial_coo_to_csr_s"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "nnz_per_s": sum(v["nnz_per_s"] for v in vals) / len(vals),
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import formoniq_b200 as fq
    from formoniq_b200.dist import SlabPartition, exchange_halo, slab_of

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # host threads that widen the downloaded u32 index arrays to usize (e2e leg): the ranks of a node share its cores
    # (with fewer than 6 threads per rank the device-side widening is the faster one: 8 ranks on a 32-core host)
    widen = min(16, (os.cpu_count() or 8) // max(world, 1) - 1)
    os.environ.setdefault("FQ_HOST_WIDEN_THREADS", str(widen if widen >= 6 else 0))
    # stdout carries exactly ONE JSON line: anything a library prints (e.g. NCCL's
    # version banner) goes to stderr until the result is ready.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    ctx = fq.Context(local, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    shape = [n, n, n * world]  # weak scaling: one n^3 cube of boxes per GPU
    slab = slab_of(rank, world, shape[2])
    mesh = fq.Mesh.kuhn(ctx, DIM, shape, slab=slab)
    forms = hodge_forms(fq)
    ctx.set_timing(True)
    # HodgeBlocks plan: symbolic once (pattern + cell-slot -> nnz maps), rows owned by this rank
    hb = fq.HodgeBlocks.symbolic(mesh, GRADE, sigma_rows=mesh.owned_range(GRADE - 1), u_rows=mesh.owned_range(GRADE))
    mats = [(name, form, a) for (name, form), a in zip(forms, hb.blocks)]
    ctx.timing_report()
    owned_cells = mesh.nowned_cells
    # first numeric pass of a fresh HodgeBlocks: builds the tile plan (= the symbolic phase: structural patterns + record
    # streams), runs the fused kernel on the structural pattern, compacts to the reference's `!= 0.0` pattern and
    # retargets the streams.  Reported separately: this is what a one-shot assembly pays.
    torch.cuda.synchronize()
    t_first = time.perf_counter()
    hb.numeric(mesh, True)
    torch.cuda.synchronize()
    first_pass_wall_ms = 1e3 * (time.perf_counter() - t_first)
    first_report = ctx.timing_report()
    plan_build_ms = hb.blocks[1].plan_build_ms
    cell_visits = hb.blocks[1].plan_cell_visits
    # the same first pass once more on a second HodgeBlocks of the same mesh: the device-memory cache of the library is
    # warm now, so this is the plan build without the multi-GB cudaMalloc calls of a cold process
    hb_warm = fq.HodgeBlocks.symbolic(mesh, GRADE, sigma_rows=mesh.owned_range(GRADE - 1), u_rows=mesh.owned_range(GRADE))
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    hb_warm.numeric(mesh, True)
    torch.cuda.synchronize()
    warm_pass_wall_ms = 1e3 * (time.perf_counter() - t_warm)
    warm_report = ctx.timing_report()
    warm_plan_ms = hb_warm.blocks[1].plan_build_ms
    del hb_warm

    def step():
        hb.numeric(mesh, True)  # one fused element kernel + one segmented reduction per block

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up: nvidia-smi needs ~0.1-1 s to deliver its first sample
    for _ in range(args.warmup):
        step()
    ctx.timing_report()
    if rank == 0:
        sampler.wait_first()
    launches0 = ctx.launch_count
    barrier()
    if rank == 0:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    kern = ctx.timing_report()
    clocks = None
    if rank == 0:
        extended = False
        if len(sampler.lines) - sampler.first < 2:
            # the timed region was shorter than two sampling periods on this box: keep the same kernel running (untimed)
            # until the sampler has seen it, and say so
            extended = True
            t_ext = time.time()
            while len(sampler.lines) - sampler.first < 3 and time.time() - t_ext < 2.0:
                step()
                torch.cuda.synchronize()
        clocks = sampler.stop()
        if extended:
            clocks["note"] = "timed region shorter than the sampling period: sampled over extra untimed steps of the same kernel"
        ctx.timing_report()
    nnz_local = sum(a.nnz for _, _, a in mats)
    # SURVEY 8d per-block formula summed over the blocks, and the same with the inputs the fused launch shares (edge
    # lengths + cell -> edge ids) counted ONCE: the second one is what one launch of the fused kernel has to move
    asm_bytes_per_block_sum = sum(a.assembly_bytes for _, _, a in mats)
    shared_bytes = max(a.assembly_shared_bytes for _, _, a in mats)
    asm_bytes = shared_bytes + sum(a.assembly_bytes - a.assembly_shared_bytes for _, _, a in mats)

    # ---- SpMV of the KKT operator [[M0, -dif_test], [dif_test^T, dif_both]] that MINRES / Lanczos apply (1 GPU)
    kkt_entry = None
    if world == 1 and not args.no_kkt:
        try:
            kkt = hb.mixed_hodge_laplacian()
            xk = fq.DeviceVector.from_torch(ctx, torch.cos(torch.arange(kkt.shape[1], device="cuda", dtype=torch.float64) ** 2 + 1.0))
            yk = fq.DeviceVector(ctx, kkt.shape[0])
            for _ in range(max(args.warmup, 1)):
                kkt.apply(xk, yk)
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            for _ in range(args.steps):
                kkt.apply(xk, yk)
            k1.record(stream)
            torch.cuda.synchronize()
            kms = k0.elapsed_time(k1) / args.steps
            kkt_entry = {"nnz": kkt.nnz, "n": kkt.shape[0], "ms": kms, "gbs": kkt.spmv_bytes / 1e9 / (kms / 1e3),
                         "bytes": kkt.spmv_bytes}
            del kkt, xk, yk
        except Exception as exc:
            print(f"[bench] KKT SpMV skipped: {exc}", file=sys.stderr)

    # ---- SpMV: every block, K applies each (halo exchange included when world > 1)
    spmv = []
    for name, form, a in mats:
        tg = form.trial_grade()
        part = SlabPartition(DIM, shape, world, tg)
        r = part.ranges[rank]
        xw = torch.cos(torch.arange(r.held_lo, r.held_hi, device="cuda", dtype=torch.float64) ** 2 + 1.0)
        x = fq.DeviceVector.from_torch(ctx, xw)
        b, e = a.row_range
        y = fq.DeviceVector(ctx, e - b)
        for _ in range(max(args.warmup, 1)):
            exchange_halo(xw, part, rank)
            a.apply_window(x, r.held_lo, y)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.timing_report()
        s0.record(stream)
        for _ in range(args.steps):
            exchange_halo(xw, part, rank)
            a.apply_window(x, r.held_lo, y)
        s1.record(stream)
        barrier()
        kr = ctx.timing_report().get("k4_spmv", {"ms": 0.0, "count": 1})
        entry = {"block": name, "ms": s0.elapsed_time(s1) / args.steps, "kernel_ms": kr["ms"] / max(kr["count"], 1),
                 "bytes": a.spmv_bytes - 8 * a.shape[1] + 8 * (r.held_hi - r.held_lo), "nnz": a.nnz, "peer_ms": None}
        if world > 1 and not args.no_peer:
            # the same product with the exchange fused into the SpMV: neighbours' columns are loaded over NVLink
            # (CUDA IPC peer memory) by the gather itself; epoch flags order it against the producers of x
            try:
                from formoniq_b200.dist import PeerHalo

                xv = fq.DeviceVector(ctx, r.held_hi - r.held_lo)
                fq._lib.check(fq._lib.lib().fq_vec_copy(ctx._h, xv._h, x._h))
                ph = PeerHalo(ctx, part, rank, xv)
                for _ in range(max(args.warmup, 1)):
                    ph.publish(); ph.apply(a, y); ph.release()
                barrier()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record(stream)
                for _ in range(args.steps):
                    ph.publish(); ph.apply(a, y); ph.release()
                p1.record(stream)
                barrier()
                ph.check()
                entry["peer_ms"] = p0.elapsed_time(p1) / args.steps
                # the fused kernel must give the bits of exchange + windowed SpMV
                exchange_halo(xw, part, rank)
                y_ref = a.apply_window(x, r.held_lo).to_numpy()
                ph.publish(); ph.apply(a, y); ph.release()
                torch.cuda.synchronize()
                entry["peer_parity"] = bool(np.array_equal(y.to_numpy(), y_ref))
                del ph, xv
            except Exception as exc:  # never lose the bench line to the optional fused measurement
                print(f"[bench] fused peer SpMV skipped on rank {rank}: {exc}", file=sys.stderr)
                entry["peer_ms"] = None
        spmv.append(entry)
        del x, y, xw

    # ---- strong scaling: ONE fixed n^3 Kuhn cube (12.58 M tets at n = 128) cut into z-slabs over the ranks.  Assembly of
    # the four blocks (owner computes: the halo layer is recomputed, no collective) and the mixed-operator application
    # the Krylov / Lanczos solvers do per iteration: halo exchange of the two column windows (NCCL send/recv with the
    # z-neighbours) + the four windowed SpMVs.  The second one is the line with a collective in it.
    strong = None
    if not args.no_strong and n >= world:
        try:
            from formoniq_b200.dist import DistKktPencil

            del hb, mats
            fq._lib.lib().fq_device_cache_trim()
            pencil = DistKktPencil(ctx, DIM, [n, n, n], GRADE, rank, world)
            for _ in range(2):
                pencil.hb.numeric(pencil.mesh, True)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(args.steps):
                pencil.hb.numeric(pencil.mesh, True)
            a1.record(stream)
            barrier()
            asm_ms = a0.elapsed_time(a1) / args.steps
            xk = fq.DeviceVector.from_torch(ctx, torch.cos(torch.arange(pencil.n, device="cuda", dtype=torch.float64) ** 2 + 1.0))
            yk = fq.DeviceVector(ctx, pencil.n)
            for _ in range(3):
                pencil.a_apply(xk, yk)
            barrier()
            a0.record(stream)
            for _ in range(args.steps):
                pencil.a_apply(xk, yk)
            a1.record(stream)
            barrier()
            op_ms = a0.elapsed_time(a1) / args.steps
            blocks = [pencil.hb.mass_sigma, pencil.hb.dif_test, pencil.dif_trial, pencil.hb.dif_both]
            strong = {"asm_ms": asm_ms, "op_ms": op_ms, "op_bytes": sum(b.spmv_bytes for b in blocks),
                      "op_nnz": sum(b.nnz for b in blocks), "cells": 6 * n ** 3}
            del pencil, xk, yk
        except Exception as exc:
            print(f"[bench] strong-scaling section skipped on rank {rank}: {exc}", file=sys.stderr)
            strong = None

    # ---- reduce over ranks (max time, sum of work)
    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_total = allmax(ms_total)
    cells_all, nnz_all, bytes_all = allsum(owned_cells), allsum(nnz_local), allsum(asm_bytes)
    spmv_ms = [allmax(s["ms"]) for s in spmv]
    # a rank that skipped the fused measurement reports -1 so that every rank takes part in the reduction
    peer_all = [allmax(s["peer_ms"] if s["peer_ms"] is not None else -1.0) for s in spmv]
    peer_min = [-allmax(-(s["peer_ms"] if s["peer_ms"] is not None else -1.0)) for s in spmv]
    peer_ms = [pa if pm >= 0.0 else None for pa, pm in zip(peer_all, peer_min)]
    if world == 1:
        peer_ms = [None for _ in spmv]
    spmv_bytes = [allsum(s["bytes"]) for s in spmv]
    peer_parity = None
    if world > 1 and all("peer_parity" in s for s in spmv):
        peer_parity = allsum(sum(0 if s["peer_parity"] else 1 for s in spmv)) == 0
    secs = ms_total / 1e3
    value = cells_all * args.steps / secs
    ok_all = allsum(0 if strong is not None else 1) == 0
    strong_out = None
    if ok_all:
        s_asm, s_op = allmax(strong["asm_ms"]), allmax(strong["op_ms"])
        s_bytes, s_nnz = allsum(strong["op_bytes"]), allsum(strong["op_nnz"])
        strong_out = {"mesh": f"{n}^3 Kuhn cube ({strong['cells']} tets) over {world} z-slabs",
                      "assembly_ms": s_asm, "assembly_elements_per_s": strong["cells"] / (s_asm / 1e3),
                      "kkt_apply_ms": s_op, "kkt_apply_gbs": s_bytes / 1e9 / (s_op / 1e3), "kkt_nnz": int(s_nnz),
                      "kkt_apply": "halo exchange of the sigma / u windows (NCCL send/recv) + 4 windowed SpMVs per apply"
                                   if world > 1 else "4 windowed SpMVs per apply (1 GPU: no exchange)"}

    # ---- e2e through the C ABI with host buffers (rank-local; H2D + symbolic + numeric + D2H per step)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, fq, ctx, mesh, forms, shape, slab, rank, world, allmax, allsum, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    per_launch = {k: v["ms"] / max(v["count"], 1) for k, v in kern.items()}
    asm_kernel_ms = sum(v["ms"] for k, v in kern.items() if k.startswith(("k1", "k3"))) / args.steps  # k13_tile_fused included
    fused = kern.get("k13_tile_fused", {}).get("count", 0) > 0
    kernel_name = "tile_fused_kernel<n3_hodge1, 16+16 warps>" if fused else "k1_elmat + k3_gather (slab path)"
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r02_tile_traffic.json")
    if fused and os.path.exists(tp):  # dram__bytes_read+write per launch of the same kernel and workload (ncu --set full)
        tj = json.load(open(tp))
        if tj.get("n") == n and tj.get("cells") == owned_cells and tj.get("kernel") == kernel_name:
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    achieved = (asm_bytes / 1e9) / (asm_kernel_ms / 1e3) if asm_kernel_ms > 0 else 0.0
    big = max(range(len(spmv)), key=lambda i: spmv[i]["nnz"])
    fused_ms = per_launch.get("k13_tile_fused", 0.0)
    # FP64 pipe of K1: the fused 3-D k = 1 Hodge tape is 386 add/sub + 159 mul + 6 div + 1 sqrt per cell visit, never
    # contracted to FMA; the peak is the measured DADD/DMUL throughput (scripts/fp64_peak.cu -> profiles/r02_fp64_peak.json)
    fp64 = None
    fp = os.path.join(ROOT, "profiles", "r02_fp64_peak.json")
    if fused and fused_ms > 0 and os.path.exists(fp):
        pk = json.load(open(fp))
        flops = 552.0 * cell_visits
        fp64 = {"k1_fp64_ops_per_launch": flops, "cell_visits": cell_visits, "achieved_tflops": flops / (fused_ms / 1e3) / 1e12,
                "peak_nofma_tflops": pk["fp64_nofma_tflops"], "frac_of_nofma_peak": flops / (fused_ms / 1e3) / 1e12 / pk["fp64_nofma_tflops"],
                "peak_fma_tflops": pk["fp64_fma_tflops"]}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3-D Hodge-Laplace k=1 mixed (AFW): numeric assembly of M0, M1, dif_test(1), dif_both(2) on a Kuhn "
                               f"grid {shape[0]}x{shape[1]}x{shape[2]} ({int(cells_all)} tets), `!= 0.0` pattern semantics",
                   "cells_per_gpu": owned_cells, "l2_policy": "inputs_larger_than_l2",
                   "parallelism": f"owner-computes z-slabs x{world}, no assembly collective"},
        "nnz_per_s": nnz_all * args.steps / secs, "nnz": int(nnz_all),
        "spmv": {"block": spmv[big]["block"], "gbs": spmv_bytes[big] / 1e9 / (spmv_ms[big] / 1e3),
                 "ms": spmv_ms[big], "frac_of_hbm_peak": spmv_bytes[big] / 1e9 / (spmv_ms[big] / 1e3) / (peak * world),
                 "all_blocks_gbs": sum(spmv_bytes) / 1e9 / (sum(spmv_ms) / 1e3),
                 "halo_exchange": "nccl send/recv" if world > 1 else "none",
                 "fused_peer_ms": peer_ms[big],
                 "fused_peer_gbs": (spmv_bytes[big] / 1e9 / (peer_ms[big] / 1e3)) if peer_ms[big] else None,
                 "spmv_multi_gpu_parity": peer_parity,
                 "kkt": None if kkt_entry is None else {**kkt_entry, "frac_of_hbm_peak": kkt_entry["gbs"] / peak}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel_name,
                     "algorithmic_bytes_fused": asm_bytes, "algorithmic_bytes_per_block_sum": asm_bytes_per_block_sum,
                     "peak_source": peak_src, "kernel_ms_per_launch": per_launch, "fp64_pipe": fp64},
        "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in kern.items()},
        "first_pass": {"wall_ms": first_pass_wall_ms, "tile_plan_ms": plan_build_ms,
                       "device_ms": {k: round(v["ms"], 3) for k, v in first_report.items()},
                       "warm": {"wall_ms": warm_pass_wall_ms, "tile_plan_ms": warm_plan_ms,
                                "device_ms": {k: round(v["ms"], 3) for k, v in warm_report.items()},
                                "note": "the same pass for a second HodgeBlocks on the same mesh (device-memory cache warm)"},
                       "note": "first numeric pass after symbolic(): plan build (structural pattern + streams) + fused kernel "
                               "+ compaction + retarget; later passes run the fused kernel alone"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "strong_scaling": strong_out,
    }
    if fused and args.slab_ms:  # break-even of the plan against re-running the two-kernel slab path (FQ_NO_TILE=1 bench)
        out["first_pass"]["break_even_steps_vs_slab"] = plan_build_ms / max(args.slab_ms - fused_ms, 1e-9)
        out["first_pass"]["warm"]["break_even_steps_vs_slab"] = warm_plan_ms / max(args.slab_ms - fused_ms, 1e-9)
        out["first_pass"]["slab_path_ms_per_step"] = args.slab_ms
    if e2e is not None:
        out["e2e"] = e2e
    try:
        if args.no_cpu:
            raise RuntimeError("skipped (--no-cpu)")
        cb = cpu_assembly_sample(args.sample_n)
        out["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
                               "nnz_per_s": cb["nnz_per_s"], "parallel_elmat_s": cb["parallel_elmat_s"],
                               "serial_coo_to_csr_s": cb["serial_coo_to_csr_s"]}
    except Exception as exc:  # the oracle is only the checker; never let it break the GPU line
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, fq, ctx, mesh, forms, shape, slab, rank, world, allmax, allsum, barrier):
    """Same metric through the reference-facing C ABI with HOST buffers: per
    step the mesh tables and edge lengths go host->device from pinned memory
    (fq_mesh_create), the four blocks are assembled from scratch (symbolic +
    numeric, like BilinearForm::assemble) and the CSR arrays come back to the
    host as usize/usize/f64 (fq_csr_download)."""
    import numpy as np
    import torch

    n_e2e = min(args.e2e_n, args.n)
    eshape = [n_e2e, n_e2e, n_e2e]
    # host-side Complex tables by the closed form (untimed setup), lengths from the device generator
    gen = fq.Mesh.kuhn(ctx, DIM, eshape)
    lengths = gen.lengths()
    ns = fq.kuhn_counts(DIM, eshape)
    del gen

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    # FaceIncidence tables of the grades the four blocks use as rows / columns (0 and 1 for k = 1; the header allows NULL
    # for grades the caller will not use)
    used = sorted({g for _, f in forms for g in (f.test_grade(), f.trial_grade())})
    faces_t = [pinned(fq.kuhn_cell_faces_host(DIM, eshape, j).view(np.int64)) if j in used else None for j in range(DIM + 1)]
    faces = [None if t is None else t.numpy().view(np.uint64) for t in faces_t]
    len_t = pinned(lengths)
    h2d = sum(f.nbytes for f in faces if f is not None) + len_t.numpy().nbytes
    cells = ns[DIM]
    d2h = pcie = 0
    widen_threads = int(os.environ.get("FQ_HOST_WIDEN_THREADS", "0"))
    times = []
    # the caller's result buffers (the Vec<usize>/Vec<f64> of the four CsrMatrix): pinned, sized by the warm-up pass.
    # Each block's download is enqueued on the library's copy stream and overlaps the assembly of the next block;
    # the step ends when all four results are in host memory.
    outs = None
    sizes = [None] * len(forms)
    # blocks in HodgeBlocks order (issuing the two large ones first was measured: 262 instead of 234 ms, the widening
    # threads of the big blocks then compete with the host side of the remaining assemblies)
    order = list(range(len(forms)))
    if os.environ.get("FQ_E2E_ORDER"):
        order = [int(v) for v in os.environ["FQ_E2E_ORDER"].split(",")]
    for it in range(args.e2e_steps + 2):  # pass 0 sizes the result buffers, pass 1 warms the allocations up
        barrier()
        t0 = time.perf_counter()
        m = fq.Mesh.from_arrays(ctx, DIM, ns, faces, len_t.numpy())
        d2h = pcie = 0
        live = []
        for i in order:
            form = forms[i][1]
            a = form.assemble(m, True)
            if outs is None:
                rp, ci, va = a.download()
                sizes[i] = (rp.shape[0], ci.shape[0])
            else:
                rp, ci, va = a.download_async(outs[i])
                live.append(a)  # must outlive the copies
            d2h += rp.nbytes + ci.nbytes + va.nbytes
            pcie += (rp.nbytes + ci.nbytes) // (2 if widen_threads > 0 else 1) + va.nbytes
            del a, rp, ci, va
        ctx.wait_downloads()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        del m, live
        if it > 1:
            times.append(dt)
        elif it == 0:
            outs = [(torch.empty(nr, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64),
                     torch.empty(max(nz, 1), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64),
                     torch.empty(max(nz, 1), dtype=torch.float64, pin_memory=True).numpy()) for nr, nz in sizes]
    if os.environ.get("FQ_BENCH_VERBOSE"):
        print("e2e step times (ms):", [round(1e3 * x, 1) for x in times], file=sys.stderr)
    t = allmax(sum(times) / len(times))
    return {"value": allsum(cells) / t, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": t * 1e3, "d2h_pcie_bytes_per_step": int(pcie),
            "workload": f"per rank: fq_mesh_create + 4x fq_assemble + fq_csr_download on a Kuhn cube "
                                                f"N={n_e2e} ({cells} tets), pinned host arrays in and out, downloads overlapped with the next block, index arrays widened to usize by {os.environ.get('FQ_HOST_WIDEN_THREADS')} host threads (0 = on the device)"}


# --------------------------------------------------------------------------- side workloads (not the driver's line)
def run_side_workload(args):
    """--workload spmv: the KKT operator of the mixed problem (what MINRES / Lanczos apply) on ONE fixed n^3 mesh cut
    into z-slabs over the ranks, halo exchange included, GB/s on the algorithmic bytes of SURVEY 8d.
    --workload krylov: device-resident MINRES on the symmetrised KKT operator (1 GPU), iterations/s.
    --workload source: config 1 — the 2-D source problem assembled and solved on the device (AFW-preconditioned MINRES).
    --workload evp: config 5 — block shift-invert Lanczos on the row-partitioned pencil (dist.DistKktPencil)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import formoniq_b200 as fq
    from formoniq_b200.dist import DistKktPencil

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = fq.Context(local, stream=torch.cuda.current_stream().cuda_stream)

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return allmax(e0.elapsed_time(e1)) / steps

    base = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if args.workload == "spmv":
        n = args.n
        pencil = DistKktPencil(ctx, 3, [n, n, n], GRADE, rank, world)
        x = pencil.seed(1)
        y = x.zeros_like()
        ms = timed(lambda: pencil.a_apply(x, y), args.steps, max(args.warmup, 3))
        blocks = [pencil.hb.mass_sigma, pencil.hb.dif_test, pencil.dif_trial, pencil.hb.dif_both]
        t = torch.tensor([float(sum(b.spmv_bytes for b in blocks)), float(sum(b.nnz for b in blocks))], device="cuda",
                         dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        out = {**base, "metric": "kkt_spmv_gbs", "value": t[0].item() / 1e9 / (ms / 1e3), "unit": "GB/s", "ms_per_step": ms,
               "config": {"workload": f"KKT operator apply (4 windowed SpMVs + halo exchange of the sigma / u windows) on a "
                                      f"{n}^3 Kuhn cube over {world} z-slabs", "kkt_nnz": int(t[1].item()), "n": pencil.n_global}}
    elif args.workload == "krylov":
        n = min(args.n, 64)
        mesh = fq.Mesh.kuhn(ctx, 3, [n, n, n])
        kkt = fq.HodgeBlocks.compute(mesh, GRADE).mixed_hodge_laplacian(symmetrized=True)
        b = fq.DeviceVector.from_numpy(ctx, ((7 * np.arange(kkt.shape[0])) % 13 - 6).astype(np.float64))
        iters = 300
        ms = timed(lambda: fq.minres(kkt, None, b, fq.StopCriterion(1e-30, iters)), max(args.steps // 4, 2), 1)
        out = {**base, "metric": "minres_iterations_per_s", "value": iters / (ms / 1e3), "unit": "iterations/s", "ms_per_step": ms,
               "scaling": "none", "config": {"workload": f"device-resident MINRES (CUDA-graph replay, scalars on the device) on "
                                                         f"the symmetrised KKT operator of a {n}^3 Kuhn cube", "n": kkt.shape[0],
                                             "nnz": kkt.nnz, "iterations_per_step": iters}}
    elif args.workload == "source":
        # config 1: the mixed source problem on 1-forms of the unit square (problems/elliptic.rs:132-182), assembled and
        # solved on the device: HodgeBlocks + hdif_gram blocks, MINRES with the AFW block preconditioner (inner device-
        # resident Jacobi-CG solves instead of the reference's sparse Cholesky)
        n = args.source_n
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mesh = fq.Mesh.kuhn(ctx, 2, [n, n])
        hb = fq.HodgeBlocks.compute(mesh, 1)
        kkt = hb.mixed_hodge_laplacian(symmetrized=True)
        wc = fq.WhitneyComplex(mesh)
        blocks = [wc.hdif_gram(0), wc.hdif_gram(1)]
        ntot = hb.n_sigma + hb.n_u
        torch.cuda.synchronize()
        t_asm = time.perf_counter() - t0
        b = fq.DeviceVector.from_numpy(ctx, ((np.arange(ntot) % 7) - 3).astype(np.float64))  # the probe of elliptic.rs:270
        t1 = time.perf_counter()
        x, rep, inner = fq.minres_blockdiag(kkt, blocks, [0, hb.n_sigma, ntot], b, fq.StopCriterion(1e-10, 500),
                                            fq.StopCriterion(1e-13, 20000))
        torch.cuda.synchronize()
        t_solve = time.perf_counter() - t1
        y = kkt.apply(x)
        y.add_scaled(-1.0, b)
        out = {**base, "metric": "source_problem_seconds", "value": t_asm + t_solve, "unit": "s", "higher_is_better": False,
               "ms_per_step": 1e3 * (t_asm + t_solve), "steps": 1, "warmup": 0, "scaling": "none",
               "config": {"workload": f"2-D Hodge-Laplace source problem on 1-forms, Kuhn grid {n}x{n}: mesh + HodgeBlocks + hdif_gram "
                                      f"on the device, MINRES with the AFW block preconditioner to 1e-10", "unknowns": ntot,
                          "kkt_nnz": kkt.nnz},
               "assembly_seconds": t_asm, "solve_seconds": t_solve, "outer_minres_iterations": rep.iters, "converged": rep.converged,
               "inner_cg_iterations": int(inner), "true_relative_residual": y.norm() / b.norm()}
    else:
        g = args.evp_grid
        pencil = DistKktPencil(ctx, 3, [g, g, max(g, world)], GRADE, rank, world, precond=args.evp_precond)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals, _ = fq.shift_invert_lanczos(pencil, 5.0, 3)
        torch.cuda.synchronize()
        secs = allmax(time.perf_counter() - t0)
        out = {**base, "metric": "evp_seconds", "value": secs, "unit": "s", "higher_is_better": False, "ms_per_step": secs * 1e3,
               "steps": 1, "warmup": 0,
               "config": {"workload": f"3-D Hodge-Laplace k=1 EVP (3 eigenpairs nearest 5.0), Kuhn grid {g}x{g}x{max(g, world)} over "
                                      f"{world} ranks, inner solves MINRES (precond {args.evp_precond})", "n": pencil.n_global},
               "eigenvalues": [float(v) for v in vals], "kkt_applies": pencil.applies,
               "inner_minres_iterations": pencil.inner_iterations}
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=128, help="boxes per axis per GPU (128 -> 12.58 M tets)")
    ap.add_argument("--sample-n", type=int, default=64, help="Kuhn cube size of the CPU sample (0: calibrate to ~20 s)")
    ap.add_argument("--no-kkt", action="store_true", help="skip the KKT-operator SpMV")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling section (fixed n^3 mesh over the ranks)")
    ap.add_argument("--slab-ms", type=float, default=10.95,
                    help="ms per step of the two-kernel slab path on this workload (FQ_NO_TILE=1 run; default: the "
                         "round-1 measurement, profiles/r01_v11_bench_n128.json) for the plan break-even")
    ap.add_argument("--e2e-n", type=int, default=128)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (tuning runs only)")
    ap.add_argument("--no-peer", action="store_true", help="skip the fused peer-memory SpMV measurement (N > 1)")
    ap.add_argument("--workload", default="assembly", choices=["assembly", "spmv", "krylov", "evp", "source"],
                    help="assembly = the north-star line the driver reads; the others are side measurements")
    ap.add_argument("--evp-grid", type=int, default=6)
    ap.add_argument("--source-n", type=int, default=256)
    ap.add_argument("--evp-precond", default="none", choices=["none", "afw"])
    args = ap.parse_args()
    if args.sample_n == 0:
        args.sample_n = None
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "assembly":
        run_side_workload(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
