import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','smsp__warps_eligible.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active',
 'dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg','sm__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum']
def summarize(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr,units,vals=rows[0],rows[1],rows[2]
    d={h:(vals[i],units[i]) for i,h in enumerate(hdr)}
    lines=[]
    lines.append("kernel: "+d.get('Kernel Name',('',''))[0][:120])
    for k in KEYS:
        if k in d: lines.append("%-75s %s %s"%(k,d[k][0],d[k][1]))
    st=[]
    for h,(v,u) in d.items():
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try: st.append((float(v),h.split('issue_stalled_')[1].replace('_per_issue_active.ratio','')))
            except: pass
    lines.append("stall reasons (warps per issue-active cycle): "+", ".join("%s %.2f"%(n,v) for v,n in sorted(st,reverse=True)[:8]))
    return "\n".join(lines)
if __name__=='__main__':
    for p in sys.argv[1:]:
        print("=== "+p); print(summarize(p)); print()
