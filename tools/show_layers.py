import json,csv,sys
for f in ('gpurun_out/bench_tc_fp32.json','gpurun_out/bench_tc_bf16.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']), d['roofline']['kernel'], round(d['roofline']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
for f in ('gpurun_out/layers_fp32.csv','gpurun_out/layers_bf16.csv'):
    rows=list(csv.reader(open(f)))[1:]
    agg={}
    for r in rows:
        if r[0]=='total': print(f, 'total', r[2]); continue
        kind=r[1].split('.')[-1] if '.' in r[1] else r[1]
        agg[kind]=round(agg.get(kind,0)+float(r[2]),1)
    print(agg)
    print([ (r[1], round(float(r[2])), r[6]) for r in rows if r[0]!='total' and float(r[2])>40])
