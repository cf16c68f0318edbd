#!/usr/bin/env python
"""Pull per-launch DRAM traffic and a few headline metrics out of `ncu --set full` captures into profiles/traffic.json.

usage: extract_traffic.py <workload> <kernel_family>=<file.ncu-rep>[:launch_index] ...
bench.py reads traffic.json to fill roofline.traffic (bytes per launch of the dominant kernel)."""
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}


def rows(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return r[0], r[1], r[2:]


def main():
    workload = sys.argv[1]
    p = os.path.join(HERE, 'traffic.json')
    data = json.load(open(p)) if os.path.exists(p) else {}
    data.setdefault(workload, {})
    for arg in sys.argv[2:]:
        fam, spec = arg.split('=')
        path, _, idx = spec.partition(':')
        hdr, units, rs = rows(path)
        r = rs[int(idx) if idx else 0]
        e = {'kernel': r[hdr.index('Kernel Name')].split('(')[0], 'capture': os.path.basename(path)}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                v = float(r[i].replace(',', ''))
                e[k] = v * SCALE.get(units[i], 1)
        e['dram_bytes_per_launch'] = e.get('dram__bytes_read.sum', 0) + e.get('dram__bytes_write.sum', 0)
        e['duration_us'] = e.pop('gpu__time_duration.sum', None)
        data[workload][fam] = e
    json.dump(data, open(p, 'w'), indent=1, sort_keys=True)
    print(json.dumps(data[workload], indent=1))


if __name__ == '__main__':
    main()
