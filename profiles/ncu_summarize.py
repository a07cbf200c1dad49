import csv, collections, sys, subprocess
rep=sys.argv[1]
subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass 2>/dev/null > /tmp/both.csv", shell=True)
subprocess.run(f"ncu -i {rep} --page raw --csv 2>/dev/null > /tmp/raw.csv", shell=True)
rows=list(csv.reader(open('/tmp/raw.csv')))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_xu.sum']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(w, units[i], [r[i] for r in data])
for i,h in enumerate(hdr):
    if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and '_not_issued' not in h:
        v=[r[i] for r in data]
        try:
            if float(v[0])>0.25: print(h.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_issue_active.ratio',''), v)
        except: pass
rows=list(csv.reader(open('/tmp/both.csv')))
cur_file=None
agg=collections.defaultdict(lambda:[0,0])
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)>=8 and r[0].isdigit():
        try: inst=int(r[7]); samp=int(r[4])
        except: continue
        k=(cur_file,int(r[0]),r[1].strip()[:80]); agg[k][0]+=inst; agg[k][1]+=samp
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values())
print('total inst', tot, 'samples', tots)
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:n]:
    print(f"{v[0]/tot*100:5.1f}% inst {v[1]/tots*100:5.1f}% samp  {k[0]}:{k[1]}  {k[2]}")
