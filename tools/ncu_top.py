"""Print the most stall-sampled SASS instructions of one kernel from an `ncu --page source --csv` dump."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
hdr=rows[1]
iS=hdr.index('# Samples'); iSrc=hdr.index('Source'); iEx=hdr.index('Instructions Executed')
data=[r for r in rows[2:] if len(r)==len(hdr) and r[iS].isdigit()]
stall=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[iS]) for r in data)
print('instructions',len(data),'total samples',tot)
agg={}
for r in data:
    for i in stall:
        agg[hdr[i]]=agg.get(hdr[i],0)+int(r[i] or 0)
print({k:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]})
top=sorted(enumerate(data), key=lambda kv:-int(kv[1][iS]))[:n]
for k,r in top:
    st={hdr[i]:int(r[i]) for i in stall if int(r[i] or 0)>0}
    st=dict(sorted(st.items(), key=lambda kv:-kv[1])[:3])
    print(k, r[iS], r[iEx], r[iSrc].strip()[:90], st)
