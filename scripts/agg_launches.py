import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki].split("(")[0][-44:]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{k:46s} n={v[0]:6d} total={v[1]/1e6:9.3f} ms share={v[1]/tot:.3f}")
