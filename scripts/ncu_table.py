"""Markdown table of the counters the north_star asks for (time, DRAM GB/s vs peak, bank conflicts, divergence) from .ncu-rep files (developer tool)."""
import csv, subprocess, sys, json
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
rows_out = []
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
    def col(r, k, scale=1.0):
        if k not in hdr: return float("nan")
        i = hdr.index(k); v = float(r[i].replace(",", "")) if r[i] else float("nan"); u = units[i]
        f = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return v * f * scale
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        t = col(r, "gpu__time_duration.sum")
        rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
        wf = col(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"); bc = col(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        rows_out.append((name, t * 1e3, rd / 1e9, wr / 1e9, (rd + wr) / t / 1e9, 100 * (rd + wr) / t / 1e9 / peak,
                         100 * bc / wf if wf == wf and wf > 0 else 0.0, col(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                         col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                         col(r, "smsp__inst_executed.sum") / 1e9, col(r, "launch__registers_per_thread")))
print("| kernel | ms | DRAM read GB | DRAM written GB | DRAM GB/s | %% of measured HBM peak (%.0f GB/s) | shared bank conflicts (%% of wavefronts) | active lanes / instruction | issue slots busy %% | warps active %% | G warp-instr | regs |" % peak)
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for x in rows_out:
    print("| %s | %.3f | %.2f | %.2f | %.0f | %.1f | %.1f | %.1f | %.0f | %.0f | %.2f | %d |" % x)
