import csv,sys,subprocess
KEYS=['gpu__time_duration.sum','sm__throughput.avg.pct','sm__inst_executed_pipe_tensor','sm__pipe_tensor','tensor','sm__warps_active.avg.pct','smsp__inst_executed.sum','registers_per_thread','sm__issue_active.avg.pct','smsp__issue_active.avg.pct','dram__bytes_read.sum ','dram__bytes_write.sum ','lts__t_bytes.sum','smsp__average_warps_issue_stalled','launch__occupancy_limit','launch__waves','sm__inst_executed_pipe_xu','pipe_alu','pipe_fma','pipe_xu','pipe_lsu','launch__grid_size','dram__throughput','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for f in sys.argv[1:]:
    out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr=rows[0]; units=rows[1]
    for r in rows[2:]:
        d=dict(zip(hdr,r)); u=dict(zip(hdr,units))
        print('==',d['Kernel Name'][:70])
        for k in hdr:
            if any(s.strip() in k for s in KEYS) and d[k] not in ('','0','n/a') and 'pcsamp' not in k and 'not_issued' not in k:
                try:
                    v=float(d[k])
                    if 'stalled' in k and v<0.3: continue
                except: pass
                print(f'   {k:90s} {d[k]:>16s} {u[k]}')
