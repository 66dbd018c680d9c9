import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
def col(n): return hdr.index(n)
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__inst_executed.sum','lts__t_bytes.sum','l1tex__t_bytes.sum']
sel=sys.argv[2:] 
for r in rows[2:]:
    if sel and r[col('ID')] not in sel: continue
    print('====',r[col('ID')],r[col('Kernel Name')][:70])
    for w in want:
        if w in hdr: print('  ',w,'=',r[col(w)],units[col(w)])
    d=[]
    for i,h in enumerate(hdr):
        if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try: d.append((float(r[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    print('   stalls:', ' '.join('%s=%.2f'%(h,v) for v,h in sorted(d,reverse=True)[:9]))
