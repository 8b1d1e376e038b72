import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)<len(H): continue
    d=dict(zip(H,r))
    name=d['Kernel Name'].split('(')[0]
    name=name.replace('void ','')
    if 'at::' in name: name=name.split('<')[0]
    else: name=name.split('<')[0]+('<'+name.split('<')[1][:28] if '<' in name and ('gemm_tc' in name or 'mha_tc' in name) else '')
    v=float(d['Metric Value']);  u=d['Metric Unit']
    if u=='ns': v/=1000
    elif u=='ms': v*=1000
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
print('total us', round(tot,1), 'launches', sum(v[0] for v in agg.values()))
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 45]:
    print(f'{k:60s} n={v[0]:4d} total={v[1]:9.1f}us avg={v[1]/v[0]:7.1f}')
