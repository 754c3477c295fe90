"""Summarise an ncu --page source export: instructions executed and stall samples per source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = None; cur_file = ''; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit(): continue
    d = dict(zip(hdr, r))
    def num(k):
        try: return float(d.get(k, '0') or 0)
        except ValueError: return 0.0
    stalls = {k[6:]: num(k) for k in hdr if k.startswith('stall_') and '(Not Issued)' not in k}
    out.append((num('Instructions Executed'), num('# Samples'), cur_file, int(r[0]), r[1].strip()[:90], stalls))
tot_i = sum(o[0] for o in out); tot_s = sum(o[1] for o in out)
print(f'total inst {tot_i:.0f}  samples {tot_s:.0f}')
agg = {}
for o in out:
    for k, v in o[5].items(): agg[k] = agg.get(k, 0) + v
print('stalls:', ', '.join(f'{k}={v:.0f}' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0))
for o in sorted(out, key=lambda o: -o[1])[:top]:
    s = ', '.join(f'{k}={v:.0f}' for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3] if v > 0)
    print(f'{o[0]/max(tot_i,1)*100:5.1f}%i {o[1]/max(tot_s,1)*100:5.1f}%s {o[2]}:{o[3]:<4d} {o[4]}  [{s}]')
