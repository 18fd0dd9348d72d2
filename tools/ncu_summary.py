"""Condense an Nsight Compute report (.ncu-rep, `ncu --set full`) into a small per-kernel CSV for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x_summary.csv
Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps the metrics DESIGN.md / bench.py quote:
duration, DRAM bytes, L2/DRAM/SM throughput, tensor-pipe activity, registers, occupancy, top stall reasons.
"""
import csv, io, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [k for k in KEEP if k in hdr]
    agg = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, {"launches": 0, **{k: 0.0 for k in cols}})
        a["launches"] += 1
        for k in cols:
            try:
                a[k] += float(r[hdr.index(k)].replace(",", ""))
            except ValueError:
                pass
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches"] + [f"{k} [{units[hdr.index(k)]}] (mean per launch)" for k in cols])
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
            w.writerow([name, a["launches"]] + [f"{a[k] / a['launches']:.6g}" for k in cols])
    print(f"{out}: {len(agg)} kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
