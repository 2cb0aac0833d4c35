import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
for r in rows[2:]:
    d=dict(zip(hdr,r)); u=dict(zip(hdr,units))
    print('KERNEL', d['Kernel Name'][:40])
    keys=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','sm__cycles_elapsed.avg','smsp__cycles_active.avg','l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum','l1tex__t_requests_pipe_lsu_mem_global_op_red.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','smsp__sass_thread_inst_executed_op_fp32_pred_on.sum','sm__sass_inst_executed_op_global_red.sum']
    for k in keys:
        if k in d: print('  ',k,'=',d[k],u[k])
    print('  stalls (avg warps per issue):')
    for h in hdr:
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try:
                v=float(d[h])
                if v>0.05: print('     ',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,3))
            except: pass
