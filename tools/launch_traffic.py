"""Developer tool: per kernel GROUP (the names bench.py uses in `kernel_ms_per_step`) launch count, time and DRAM traffic from an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.

    python tools/launch_traffic.py gpurun_out/launches.csv [steps_in_capture] > profiles/dominant_traffic.json
"""
import csv
import json
import re
import sys
from collections import defaultdict

EPI = {0: 'bias', 1: 'relu_stats', 2: 'relu_maxmin', 3: 'bnrelu_bwd'}


def group_of(name):
    m = re.search(r'gemm_nt_tc3_kernel<(?:\(int\))?(\d)', name)
    if m:
        return 'nt_gemm_nt[%s,plain]' % EPI[int(m.group(1))]
    m = re.search(r'gemm_nt_tc2?_kernel<(?:\(int\))?(\d), (?:\(int\))?(\d)', name)
    if m:
        return 'nt_gemm_nt[%s,%s]' % (EPI[int(m.group(2))], 'edge' if m.group(1) == '1' else 'plain')
    if 'gemm_tn_tc_kernel' in name or 'tn_reduce_kernel' in name:
        return 'nt_gemm_tn*'
    m = re.search(r'nt::(\w+?)(?:_v4)?_kernel', name)
    if m:
        return 'nt_' + m.group(1)
    return None


def main(path, steps=1):
    per_id = defaultdict(dict)
    names = {}
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[0].isdigit():
            names[r[0]] = r[4]
            try:
                per_id[r[0]][r[12]] = float(r[14].replace(',', ''))
            except ValueError:
                pass
    units = {}
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[0].isdigit():
            units[r[12]] = r[13]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'ms': 1e6}
    out = defaultdict(lambda: {'launches': 0, 'time_ns': 0.0, 'dram_bytes': 0.0})
    for kid, met in per_id.items():
        g = group_of(names[kid])
        if g is None:
            continue
        o = out[g]
        o['launches'] += 1
        o['time_ns'] += met.get('gpu__time_duration.sum', 0.0) * scale.get(units.get('gpu__time_duration.sum', 'ns'), 1.0)
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            o['dram_bytes'] += met.get(key, 0.0) * scale.get(units.get(key, 'byte'), 1.0)
    res = {}
    for g, o in sorted(out.items(), key=lambda kv: -kv[1]['time_ns']):
        res[g] = {'launches_per_step': o['launches'] / steps, 'dram_bytes_per_launch': o['dram_bytes'] / o['launches'],
                  'avg_launch_us_under_ncu': o['time_ns'] / o['launches'] / 1e3}
    res['_source'] = {'file': path, 'steps_in_capture': steps,
                      'how': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; '
                             'dram bytes = read + write, averaged over the launches of the group (cold L2 per launch)'}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
