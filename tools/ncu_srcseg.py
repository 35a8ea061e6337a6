import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[ix['# Samples']].isdigit()]
tot=sum(int(r[ix['# Samples']]) for r in data)
print('samples',tot,'static',len(data))
seg=[];cur=dict(s=0,e=0,lsb=0,n=0,dp=0,ldg=0,imad=0,start=0)
for i,r in enumerate(data):
    src=r[ix['Source']]
    ex=int(r[ix['Instructions Executed']])
    cur['s']+=int(r[ix['# Samples']]);cur['e']+=ex;cur['lsb']+=int(r[ix['stall_long_sb']]);cur['n']+=1
    if re.search(r'\bD(FMA|MUL|ADD)\b',src): cur['dp']+=ex
    if re.search(r'\bIMAD',src): cur['imad']+=ex
    if 'LDG' in src: cur['ldg']+=ex
    if 'BAR.SYNC' in src or 'UCGABAR' in src or 'BAR.' in src:
        seg.append(cur);cur=dict(s=0,e=0,lsb=0,n=0,dp=0,ldg=0,imad=0,start=i+1)
seg.append(cur)
for s in seg:
    if s['e']>0 and s['s']>tot*0.004: print(f"lines {s['start']:6d}+{s['n']:5d} samples {s['s']/tot:6.1%} exec {s['e']:9d} dp {s['dp']:9d} imad {s['imad']:9d} ldg {s['ldg']:8d} long_sb {s['lsb']/max(1,s['s']):5.1%}")
for k in ['stall_barrier','stall_dispatch','stall_lg','stall_long_sb','stall_math','stall_membar','stall_mio','stall_no_inst','stall_not_selected','stall_selected','stall_short_sb','stall_wait','stall_branch_resolving']:
    print(k, f"{sum(int(r[ix[k]]) for r in data)/tot:.1%}")
