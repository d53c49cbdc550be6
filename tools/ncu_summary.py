"""Extract the judged metrics from an .ncu-rep into a small text summary (profiles/).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt ["note ..."]
"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of %s" % rep, "# " + note, ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append("kernel: %s" % d.get("Kernel Name", "?"))
        for k in KEYS:
            if k in d:
                lines.append("  %-90s %s %s" % (k, d[k], units[hdr.index(k)]))
        stalls = sorted(((float(v), h) for h, v in d.items()
                         if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        lines.append("  warp stall reasons (warps per issue-active cycle):")
        for v, h in stalls[:8]:
            lines.append("    %-40s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
