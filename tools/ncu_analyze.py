import csv, re, collections, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/prof_walk.ncu-rep'
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines()))
hdr,units,vals=r[0],r[1],r[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__grid_size','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']
for i,h in enumerate(hdr):
    if h in want: print("%s [%s] = %s"%(h,units[i],vals[i]))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Source"); ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples"); it=hdr.index("Thread Instructions Executed")
tot=sum(int(x[ie]) for x in data); tott=sum(int(x[it]) for x in data)
print("total warp instr", tot, "sass lines", len(data), "avg threads", tott/tot)
ops=collections.Counter(); opsamp=collections.Counter()
for x in data:
    m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[ia]); op=m.group(2).split('.')[0] if m else '?'
    ops[op]+=int(x[ie]); opsamp[op]+=int(x[isamp])
ss=sum(opsamp.values())
for op,c in ops.most_common(22): print("%-10s %6.2f%% instr  %6.2f%% samples"%(op,100*c/tot,100*opsamp[op]/ss))
prev=None; start=0; acc=0; acct=0
print("regions (start line, exec count/instr, lines, share of instr):")
for i,x in enumerate(data+[None]):
    c=int(x[ie]) if x else -1
    if prev is not None and (x is None or abs(c-prev)>0.25*max(c,prev,1)):
        if acc/tot>0.01: print("  %4d-%4d count %12d lines %4d share %5.1f%% thr/inst %.1f  %s"%(start,i-1,prev,i-start,100*acc/tot, acct/max(acc,1), data[start][ia][:40]))
        start=i; acc=0; acct=0
    if x: acc+=c; acct+=int(x[it]); prev=c
