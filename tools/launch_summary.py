"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, sys, collections, re
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = None; agg = collections.OrderedDict(); cnt = collections.Counter()
    for r in rows:
        if r[0] == 'ID': hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        name = re.sub(r'<.*', '', d['Kernel Name']).replace('void ', '')
        name = re.sub(r'\(.*', '', name)
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        unit = d['Metric Unit']
        if unit == 'ns': v /= 1e3
        elif unit == 'ms': v *= 1e3
        elif unit in ('s', 'second'): v *= 1e6
        agg[name] = agg.get(name, 0) + v; cnt[name] += 1
    tot = sum(v for k, v in agg.items() if k.startswith('sb::') and 'precompute' not in k)
    print(f"== {f}  (sb:: kernels excluding precompute: {tot:.1f} us total)")
    for k, v in agg.items():
        print(f"   {k:34s} n={cnt[k]:3d} total={v:10.1f} us  avg={v/cnt[k]:9.1f} us  share={100*v/tot if k.startswith('sb::') and 'precompute' not in k else 0:5.1f}%")
