import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hdr_idx[0]]
col={n:i for i,n in enumerate(h)}
sec=[r for r in rows[hdr_idx[0]+1: hdr_idx[1]-1 if len(hdr_idx)>1 else None] if len(r)>10]
tot=sum(int(r[col['# Samples']]) for r in sec)
print('kernel', rows[0][1][:80], 'total samples', tot, 'instructions', len(sec))
stalls=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
# group consecutive instructions into regions by executed count
ex=[int(r[col['Instructions Executed']]) for r in sec]
# print top regions: sliding windows of 150 instrs by samples
cum=[0]
for r in sec: cum.append(cum[-1]+int(r[col['# Samples']]))
# segment where exec count changes by >30%
segs=[];start=0
for i in range(1,len(sec)):
    a,b=ex[i-1],ex[i]
    if (a==0)!=(b==0) or (a>0 and b>0 and (b>1.5*a or a>1.5*b)):
        segs.append((start,i)); start=i
segs.append((start,len(sec)))
big=sorted(segs,key=lambda s:-(cum[s[1]]-cum[s[0]]))[:10]
for s0,s1 in sorted(big):
    smp=cum[s1]-cum[s0]
    agg={st:sum(int(r[col[st]]) for r in sec[s0:s1]) for st in stalls}
    top=sorted(agg.items(), key=lambda x:-x[1])[:5]
    nm=sum('MUFU.EX2' in r[col['Source']] for r in sec[s0:s1])
    print(f'instrs {s0}-{s1} ({s1-s0}) exec~{ex[s0]} samples {smp} ({100*smp/tot:.1f}%) mufu {nm}', [(k.replace('stall_',''),v) for k,v in top])
if len(sys.argv)>2:
    s0,s1=int(sys.argv[2]),int(sys.argv[3])
    for r in sec[s0:s1]:
        st={s.replace('stall_',''):int(r[col[s]]) for s in stalls if int(r[col[s]])>0}
        print(r[col['Source']].strip()[:70].ljust(70), r[col['# Samples']].rjust(5), dict(sorted(st.items(), key=lambda x:-x[1])[:3]))
