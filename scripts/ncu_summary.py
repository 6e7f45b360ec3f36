"""Key metrics of every kernel launch in an ncu report: python scripts/ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print(f"==== {d.get('Kernel Name', '?')}  block {d.get('Block Size')}  grid {d.get('Grid Size')}")
    for h, u, v in zip(hdr, units, vals):
        if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.04):
            print(f"{h:90s} {u:12s} {v}")
