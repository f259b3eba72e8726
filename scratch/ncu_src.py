import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]
data=[r for r in rows[2:] if len(r)==len(h) and r[0].startswith('0x')]
si=h.index("# Samples"); src=h.index("Source")
tot=sum(int(r[si] or 0) for r in data)
print("total samples",tot, "rows", len(data))
agg=collections.Counter(); cnt=collections.Counter()
for r in data:
    t=r[src].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    agg[op]+=int(r[si] or 0); cnt[op]+=1
for op,v in agg.most_common(12): print(op, v, "%.1f%%"%(100*v/tot), cnt[op])
stall=[(i,n) for i,n in enumerate(h) if n.startswith('stall_')]
st=collections.Counter()
for r in data:
    for i,n in stall:
        try: st[n]+=int(r[i] or 0)
        except: pass
print([(n,v) for n,v in st.most_common(10)])
top=sorted(data,key=lambda r:-int(r[si] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 20]
for r in top:
    reasons=sorted([(int(r[i] or 0),n) for i,n in stall if (r[i] or '0')!='0'],reverse=True)[:2]
    print(r[si], r[src][:90], reasons)
