import json, sys
d = json.load(open(sys.argv[1]))
ak = d['roofline']['all_kernels']
rows = [(v['ms_per_step'], k, v['launches_per_step'], v.get('TFLOP/s'), v.get('GB/s')) for k, v in ak.items()]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows if r[1].startswith('all:'))
print('sum of all: %.1f ms   step %.1f ms   value %.1f %s' % (tot, d['ms_per_step'], d['value'], d['unit']))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in [r for r in rows if r[1].startswith('all:')][:n]:
    print('%8.2f ms  %-28s n=%4d' % (r[0], r[1], r[2]))
for r in [r for r in rows if not r[1].startswith('all:')]:
    print('%8.2f ms  %-34s n=%4d  TF=%s GB/s=%s' % (r[0], r[1], r[2], None if r[3] is None else round(r[3], 1), None if r[4] is None else round(r[4])))
