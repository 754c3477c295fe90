"""Summarise `ncu --set full` reports into the text files committed under profiles/.
usage: python tools/ncu_summary.py <tag> <report.ncu-rep> [...]   (reads here, on the CPU container; ncu -i needs no GPU)
Writes profiles/<kernel>_<tag>_summary.txt (headline metrics + per-source-line stall table) and updates
profiles/traffic.json (DRAM bytes per launch, consumed by bench.py's roofline.traffic)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def to_bytes(v, unit):
    m = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(v.replace(',', '')) * m.get(unit, 1)


def main():
    tag = sys.argv[1]
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for rep in sys.argv[2:]:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        kname = d['Kernel Name'][0].split('(')[0].split('<')[0].replace('void ', '').strip()
        lines = [f'# {os.path.basename(rep)}  kernel {d["Kernel Name"][0][:100]}', f'# ncu --set full --clock-control none --import-source on (one launch; cold-ish caches)']
        for k in KEYS:
            if k in d:
                lines.append(f'{k:75s} {d[k][0]:>16s} {d[k][1]}')
        rd, wr = to_bytes(*d['dram__bytes_read.sum']), to_bytes(*d['dram__bytes_write.sum'])
        traffic[kname] = rd + wr
        lines.append(f'dram bytes per launch (read+write) = {rd + wr:.0f}')
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
        tmp = f'/tmp/_ncu_src_{kname}.csv'
        open(tmp, 'w').write(src)
        tab = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), tmp, '30'], capture_output=True, text=True).stdout
        lines += ['', '# per source line: %instructions, %stall samples, line, top stall reasons', tab]
        out = os.path.join(ROOT, 'profiles', f'{kname}_{tag}_summary.txt')
        open(out, 'w').write('\n'.join(lines) + '\n')
        print('wrote', out)
    json.dump(traffic, open(tpath, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
