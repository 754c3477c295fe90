"""Pretty-print the interesting fields of a bench.py JSON line (development aid)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
keys = ('impl', 'value', 'unit', 'n_gpus', 'ms_per_step', 'kernel_us', 'e2e', 'full_forward', 'highres', 'train_step', 'clocks', 'cpu_baseline')
for k in keys:
    if k in d:
        print(k, '=', d[k])
if 'roofline' in d:
    r = d['roofline']; print('roofline', r['kernel'], r['bound'], 'achieved %.1f %s frac %.3f traffic %s' % (r['achieved'], r['unit'], r['frac'], r.get('traffic')))
if 'roofline_k1' in d:
    r = d['roofline_k1']; print('roofline_k1 achieved %.1f GB/s frac %.3f (survey bytes: %.1f GB/s)' % (r['achieved'], r['frac'], r['achieved_survey_bytes']))
