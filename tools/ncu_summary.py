"""Print the headline metrics of .ncu-rep files (run here, no GPU needed):  python tools/ncu_summary.py gpurun_out/*.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed.sum"]


def main():
    for f in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, u = rows[0], rows[1]
        for v in rows[2:]:
            print("=== %s :: %s" % (f, v[h.index("Kernel Name")][:90]))
            for i, n in enumerate(h):
                if n in WANT or "warp_issue_stalled" in n and n.endswith("per_warp_active.pct") and float(v[i] or 0) > 5:
                    print("  %-82s %s %s" % (n, v[i], u[i]))


if __name__ == "__main__":
    main()
