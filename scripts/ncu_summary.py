"""print the handful of ncu metrics used in profiles/*.txt from a .ncu-rep (development aid)"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, rows = r[0], r[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for v in rows:
    d = dict(zip(h, v))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d: print(f"  {k:75s} {d[k]}")
    for k, x in d.items():
        if "smsp__average_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                if float(x) > 0.3: print(f"  stall {k[34:-23]:40s} {x}")
            except ValueError:
                pass
