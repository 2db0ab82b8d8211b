"""Summarise an .ncu-rep (read on the CPU box): one line per captured launch with the metrics the roofline tables quote.
usage: python tools/ncu_summary.py gpurun_out/<file>.ncu-rep > profiles/<name>.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel", 46), ("gpu__time_duration.sum", "us", 9), ("launch__grid_size", "grid", 7), ("launch__block_size", "block", 6),
        ("launch__registers_per_thread", "regs", 5), ("dram__bytes_read.sum", "dram_rd_MB", 11), ("dram__bytes_write.sum", "dram_wr_MB", 11),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 7),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 9),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%", 8),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%", 8), ("smsp__inst_executed.sum", "warp_instr", 12),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/instr", 9)]
print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised launches: compare shares and percentages, not absolute times)")
print(" ".join(f"{n:>{w}s}" if n != "kernel" else f"{n:{w}s}" for _, n, w in cols))
for r in rows[2:]:
    out = []
    for key, n, w in cols:
        v = r[ix[key]] if key in ix else ""
        if n == "kernel":
            v = v.replace("void ", "").replace("<unnamed>::", "")
            out.append(f"{v[:w]:{w}s}")
        else:
            try:
                f = float(v)
                out.append(f"{f:>{w}.1f}" if n not in ("grid", "block", "regs", "warp_instr") else f"{int(f):>{w}d}")
            except ValueError:
                out.append(f"{v[:w]:>{w}s}")
    print(" ".join(out))
