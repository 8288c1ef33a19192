"""Print the handful of ncu metrics that drive the blind-rotation design from a .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__sass_inst_executed_op_shuffle.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_xu.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
for n, vals in enumerate(rows[2:]):            # one block per profiled kernel launch
    if len(vals) < len(hdr):
        continue
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    if n:
        print("-----")
    for k in keys:
        hit = [h for h in hdr if h == k or h.endswith("." + k)]
        if hit:
            print(f"{k:75s} {d[hit[0]][0]:>22s} {d[hit[0]][1]}")
    print("--- stall reasons (warps per issue-active cycle)")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                v = float(d[h][0] or 0)
            except ValueError:
                continue
            if v > 0.05:
                print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.3f}")
