import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ci={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[ci['# Samples']].isdigit() and r[0].startswith('0x')]
tot=sum(int(r[ci['# Samples']]) for r in data)
B=int(sys.argv[2]) if len(sys.argv)>2 else 200
keys=['stall_long_sb','stall_barrier','stall_mio','stall_math','stall_short_sb','stall_lg','stall_wait','stall_not_selected','stall_selected','stall_branch_resolving','stall_no_inst','stall_dispatch','stall_membar','stall_sleep']
for b in range(0,len(data),B):
    seg=data[b:b+B]
    n=sum(int(r[ci['# Samples']]) for r in seg)
    st={k:sum(int(r[ci[k]] or 0) for r in seg) for k in keys}
    ops={}
    for r in seg:
        op=r[ci['Source']].strip().split()[0]
        if op.startswith('@'): op=r[ci['Source']].strip().split()[1]
        ops[op]=ops.get(op,0)+1
    topops=' '.join('%s:%d'%(k,v) for k,v in sorted(ops.items(),key=lambda kv:-kv[1])[:4])
    print('%5d-%5d %5.1f%%  %s | %s'%(b,b+B,100*n/tot,' '.join('%s=%d'%(k.replace('stall_',''),v) for k,v in sorted(st.items(),key=lambda kv:-kv[1])[:4]),topops))
