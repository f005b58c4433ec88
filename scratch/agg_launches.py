import csv, collections, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0]); n=0
gem=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    full=row['Kernel Name']; v=float(row['Metric Value'].replace(',',''))
    v = v/1000 if row['Metric Unit']=='ns' else v
    name=full.replace('(anonymous namespace)::','').replace('<unnamed>::',''); name=re.sub(r'\(.*','',name); name=re.sub(r'^void ','',name); m=re.match(r'([\w:]+)(<[\d, ]+>)?',name); name=(m.group(0) if m else name)[:60]
    agg[name][0]+=1; agg[name][1]+=v; n+=1
    if 'gemm_tf32x3' in full: gem[row['Grid Size']][0]+=1; gem[row['Grid Size']][1]+=v
tot=sum(v[1] for v in agg.values())
print(n,'launches, total us',round(tot,1))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f"{k:60s} {v[0]:6d} {v[1]:10.1f} {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]:.1f}")
print('--- gemm by grid')
for k,v in sorted(gem.items(), key=lambda kv:-kv[1][1]):
    print(k, v[0], round(v[1],1), 'avg', round(v[1]/v[0],1))
