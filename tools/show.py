import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        pc=d['roofline']['per_callback']
        print(f.split('/')[-1], round(d['value'],1), 'ms', round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in pc.items()}, round(d['roofline']['all_three']['frac'],3))
    except Exception as e:
        print(f,'ERR',e, open(f).read()[-600:])
