import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
# find header row
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr,units,data=rows[hi],rows[hi+1],rows[hi+2:]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_elapsed.max']
ki=hdr.index('Kernel Name')
seen={}
for r in data:
    name=r[ki].split('(')[0]
    seen.setdefault(name,[]).append(r)
for name,rs in seen.items():
    r=rs[-1]
    print('---',name,len(rs))
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"   {w:75s} {r[i]} {units[i]}")
