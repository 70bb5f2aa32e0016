import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for r in rows[hi+1:]:
    if len(r)>vi:
        try: v=float(r[vi].replace(',',''))
        except: continue
        k=r[ki].split('(')[0]; agg[k][0]+=1; agg[k][1]+=v; agg[k][2]=max(agg[k][2],v)
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:60s} n={v[0]:4d} total {v[1]/1e3:10.1f} us  {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]/1e3:8.1f} us max {v[2]/1e3:8.1f}")
