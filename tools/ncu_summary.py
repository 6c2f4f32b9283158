import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.avg.per_second','l1tex__t_bytes.sum','launch__waves_per_multiprocessor','launch__occupancy_per']
for r in rows[2:]:
    print('-----')
    d=dict(zip(hdr,zip(units,r)))
    for h in hdr:
        if any(h==w or h.startswith(w) for w in want):
            print(' ',h,d[h][0],d[h][1])
    st=[(float(d[h][1]),h) for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and d[h][1]]
    st.sort(reverse=True)
    print('  stalls:',', '.join('%s=%.2f'%(h.split('stalled_')[1].split('_per_issue')[0],v) for v,h in st[:8]))
