"""profiles/ summary of an ncu --set full capture: the metrics quoted in DESIGN.md, the instruction mix and the hot
regions of the SASS (tools/ncu_analyze.py).  usage: ncu_summary.py <report.ncu-rep> <title>"""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, vals = r[0], r[1], r[2]
keys = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed']
print("# ncu --set full --clock-control none: %s" % title)
for i, h in enumerate(hdr):
    if h in keys or ('smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')):
        print(h, units[i], vals[i])
print("# instruction mix and hot SASS regions (tools/ncu_analyze.py)")
out = subprocess.run([sys.executable, 'tools/ncu_analyze.py', rep], capture_output=True, text=True).stdout
print("\n".join(l for l in out.splitlines() if not l.startswith(('dram__', 'gpu__', 'l1tex__', 'launch__', 'lts__', 'sm__', 'smsp__'))))
