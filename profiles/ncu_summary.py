"""Print the metrics we track from an ncu report: python profiles/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__average_warps_issue_stalled', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warp_latency_issue_stalled', 'sm__sass_inst_executed_op_shared', 'smsp__inst_executed_op_shfl',
        'sass__inst_executed_local', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__pcsamp_warps_issue_stalled']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if any(h == w or h.startswith(w) for w in WANT) and r[i] not in ('', '0'):
            print(f"  {h:80s} {r[i]:>16s} {units[i]}")
