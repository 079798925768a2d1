"""Key metrics of the kernels in an `ncu --set full` report, as text for profiles/.
    python tools/ncu_kernel_summary.py gpurun_out/r01_pair_sym_v3.ncu-rep > profiles/r01_pair_sym_v3_ncu_summary.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print("ncu --set full --clock-control none --import-source on ; report %s\n" % sys.argv[1].split("/")[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("Kernel Name".ljust(76), d.get("Kernel Name", "")[:120])
    for m in WANT:
        if m in d:
            print(m.ljust(76), units[hdr.index(m)].ljust(16), d[m])
    stalls = {h: d[h] for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    for h in sorted(stalls):
        print("  " + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")].ljust(28), stalls[h])
    print()
