"""Summarise an .ncu-rep (raw page + source page hot spots).  usage: ncu_summary.py file.ncu-rep [ncells]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
ncells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_bytes.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get("Kernel Name", "")[:100])
    for k in keys:
        if k in d:
            extra = ""
            if ncells and k in ("smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
                v = float(d[k].replace(",", ""))
                u = units[hdr.index(k)]
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
                extra = f"   per cell: {v * scale / ncells:.1f}"
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}{extra}")
    stalls = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")}
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
tot = sum(int(r[iex] or 0) for r in data); tots = sum(int(r[ismp] or 0) for r in data)
print("SASS instrs", len(data), "executed", tot, "samples", tots)
top = sorted(range(len(data)), key=lambda i: -int(data[i][ismp] or 0))[:25]
for i in sorted(top):
    r = data[i]
    print(f"  {i:5d} {r[isrc][:70]:70s} ex={int(r[iex] or 0):>10d} samples={100 * int(r[ismp] or 0) / max(tots, 1):5.2f}%")

# ---- warp-specialised kernels: split the source counters by role at the setmaxnreg instructions (producers first)
marks = [i for i, r in enumerate(data) if "USETMAXREG" in r[isrc]]
if len(marks) == 2:
    def col(name):
        return hdr.index(name) if name in hdr else None
    iws, iwi = col("L1 Wavefronts Shared"), col("L1 Wavefronts Shared Ideal")
    stall_cols = [(h[len("stall_"):], i) for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
    def num(r, i):
        try:
            return float(r[i] or 0)
        except (TypeError, ValueError):
            return 0.0
    for name, lo, hi in (("producer warps", marks[0], marks[1]), ("consumer warps", marks[1], len(data))):
        rows_ = data[lo:hi]
        inst = sum(num(r, iex) for r in rows_)
        polls = sum(num(r, iex) for r in rows_ if "SYNCS" in r[isrc] and "NANOSLEEP" not in r[isrc])
        ws = sum(num(r, iws) for r in rows_) if iws is not None else 0
        wi = sum(num(r, iwi) for r in rows_) if iwi is not None else 0
        samples = sum(num(r, ismp) for r in rows_)
        stalls = sorted(((sum(num(r, i) for r in rows_), n) for n, i in stall_cols), reverse=True)[:7]
        print(f"  {name}: {inst / 1e6:.0f} M warp instructions ({polls / 1e6:.0f} M of them mbarrier polls), "
              f"shared-memory wavefronts {ws / 1e6:.0f} M (ideal {wi / 1e6:.0f} M)")
        print("      stall samples: " + ", ".join(f"{n} {100 * v / max(samples, 1):.1f}%" for v, n in stalls))
