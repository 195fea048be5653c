import csv,sys,subprocess,collections,re
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines()))
h=r[0]; v=r[-1]
want=['gpu__time_duration.sum','dram__bytes_read.sum ','dram__bytes_write.sum ','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','smsp__issue_active.avg.pct','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','smsp__inst_executed_op_shared','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','stalled_long_scoreboard_per','stalled_short_scoreboard_per','stalled_wait_per','stalled_branch_resolving_per','stalled_lg_throttle_per','stalled_mio_throttle_per','stalled_no_instruction_per','stalled_not_selected_per','stalled_math_pipe','sm__cycles_elapsed.max ','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_active']
for i,n in enumerate(h):
    nn=n+' '
    if any(w in nn for w in want): print(f"{n} = {v[i]} {r[1][i]}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
# find header row
hi=[i for i,x in enumerate(rows) if 'Source' in x and any('Instructions Executed' in c for c in x)]
if hi:
    hdr=rows[hi[0]]; ci=hdr.index('Source'); 
    ie=[i for i,c in enumerate(hdr) if c=='# Instructions Executed' or c=='Instructions Executed'][0]
    ws=[i for i,c in enumerate(hdr) if 'Warp Stall Sampling (All' in c]
    agg=collections.Counter(); st=collections.Counter(); txt={}
    loc=[i for i,c in enumerate(hdr) if c in ('Address','#')]
    for x in rows[hi[0]+1:]:
        try: n=int(x[ie])
        except: continue
        key=x[ci].strip()[:110]
        agg[key]+=n
        if ws:
            try: st[key]+=int(x[ws[0]])
            except: pass
    tot=sum(agg.values()); tots=sum(st.values()) or 1
    print("total inst", tot)
    for k,n in agg.most_common(45): print(f"{100*n/tot:5.1f}% inst {100*st[k]/tots:5.1f}% stall | {k}")
