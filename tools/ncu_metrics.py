import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines()))
h,u,v=r[0],r[1],r[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active',
'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__t_set_accesses.avg.pct_of_peak_sustained_elapsed' ,'lts__t_sectors.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_sectors.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']
for i,x in enumerate(h):
    if x in want or ('average_warps_issue_stalled' in x and x.endswith('per_issue_active.ratio') and float(v[i] or 0)>0.15): print(x,u[i],v[i])
