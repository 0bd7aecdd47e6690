import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','sm__warps_active.avg.pct','launch__registers_per_thread ','sm__inst_executed.sum ','sm__inst_executed.sum.per_cycle_elapsed','smsp__issue_active.avg.pct','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.sum.pct','smsp__average_warps_issue_stalled','launch__grid_size','sm__cycles_elapsed.max ','smsp__warps_eligible.avg.per_cycle_active','smsp__average_warp_latency','smsp__inst_executed.sum ','sm__cycles_active.avg ','launch__occupancy_limit','sm__maximum_warps_per_active_cycle_pct','dram__bytes_read.sum','dram__bytes_write.sum']
for i,h in enumerate(hdr):
    if any(h.startswith(w.strip()) if w.endswith(' ') else (w in h) for w in want):
        try:
            v=float(vals[i].replace(',',''))
            if 'stalled' in h and v<0.05: continue
        except: pass
        print(h, units[i], vals[i])
