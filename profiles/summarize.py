"""Summarise an Nsight Compute report into a small text file that can be committed.
    python profiles/summarize.py gpurun_out/prof_x.ncu-rep profiles/x_r1.txt ["command that produced it"]
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep.split('/')[-1]}\n")
        if cmd:
            f.write(f"# command: {cmd}\n")
        f.write("# (per-launch values; captured with --clock-control none; never a bench number)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\n== {name[:150]}\n")
            for i, h in enumerate(hdr):
                if h in KEYS and r[i]:
                    f.write(f"  {h:95s} {r[i]:>18s} {units[i]}\n")
            for i, h in enumerate(hdr):
                if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and r[i]:
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v >= 0.05:
                        f.write(f"  stall {h[len(STALLS):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue\n")


if __name__ == "__main__":
    main()
