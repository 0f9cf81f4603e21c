#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers the roofline report uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.txt]"""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    kn = hdr.index("Kernel Name")
    for r in data:
        out.append(f"== {r[kn]} (id {r[0]})")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"  {w:70s} {r[i]:>16s} {units[i]}")
        stalls = [(float(r[i] or 0), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        # issue / pipe utilisation: which unit an ALU-bound kernel saturates
        pipes = []
        for i, h in enumerate(hdr):
            if ("pipe" in h or "issue_active" in h or "inst_issued" in h) and "pct_of_peak_sustained_active" in h:
                try:
                    pipes.append((float(r[i]), h.replace(".pct_of_peak_sustained_active", "")))
                except ValueError:
                    pass
        pipes.sort(reverse=True)
        out.append("  busiest pipes / issue (% of peak, active cycles): " + ", ".join(f"{n}={v:.1f}" for v, n in pipes[:8]))
        stalls.sort(reverse=True)
        out.append("  warp stall reasons (avg warps stalled per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:7]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            ur = units[hdr.index("dram__bytes_read.sum")]
            out.append(f"  dram traffic read+write = {rd + wr:.6f} {ur}")
        except Exception:
            pass
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")

if __name__ == "__main__":
    main()
