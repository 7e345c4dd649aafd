"""Development aid: the handful of ncu raw metrics used when tuning (one line per metric). usage: python scripts/ncu_brief.py a.ncu-rep ..."""
import subprocess, csv, io, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_barriers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "l1tex__t_sector_hit_rate.pct"]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
    print(d["Kernel Name"][:60], d["Block Size"], d["Grid Size"])
    for k in KEYS: print("  ", k, d.get(k), u.get(k))
    st = [(k, float(v.replace(",", ""))) for k, v in d.items() if "smsp__average_warp" in k and "issue_stalled" in k and v not in ("", "n/a")]
    for k, v in sorted(st, key=lambda x: -x[1])[:10]: print("     ", k.replace("smsp__average_warps_issue_stalled_", ""), v)
