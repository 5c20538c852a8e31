"""Print a compact summary of an `ncu --page raw --csv` dump (one or more launches)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or (h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued') and v not in ('0', '')):
            print(f"{h:76s} {u:16s} {v}")
    print('-' * 60)
