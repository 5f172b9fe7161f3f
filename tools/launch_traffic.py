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
    m = re.search(r'gemm_nt_tc[34]_kernel<(?:\(int\))?(\d)', name)
    if m:
        return 'nt_gemm_nt[%s,plain]' % EPI[int(m.group(1))]
    m = re.search(r'gemm_nt_tc2?_kernel<(?:\(int\))?(\d), (?:\(int\))?(\d)', name)
    if m:
        return 'nt_gemm_nt[%s,%s]' % (EPI[int(m.group(2))], 'edge' if m.group(1) == '1' else 'plain')
    if 'gemm_tn_tc_kernel' in name or 'tn_reduce_kernel' in name:
        return 'nt_gemm_tn*'
    m = re.search(r'nt::(\w+?)(?:_v4|_x2)?_kernel', name)
    if m:
        return 'nt_' + m.group(1)
    return None


SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'nsecond': 1.0, 'usecond': 1e3, 'msecond': 1e6}


def main(path, steps=1):
    per_id = defaultdict(dict)
    names, order = {}, []
    for r in csv.reader(open(path)):
        if len(r) > 14 and r[0].isdigit():
            if r[0] not in names:
                order.append(r[0])
            names[r[0]] = r[4]
            try:
                per_id[r[0]][r[12]] = float(r[14].replace(',', '')) * SCALE.get(r[13], 1.0)      # units are per ROW in the ncu CSV
            except ValueError:
                pass
    out = defaultdict(lambda: {'launches': 0, 'time_ns': 0.0, 'dram_bytes': 0.0})

    def add(g, kid, count=1):
        o, met = out[g], per_id[kid]
        o['launches'] += count
        o['time_ns'] += met.get('gpu__time_duration.sum', 0.0)
        o['dram_bytes'] += met.get('dram__bytes_read.sum', 0.0) + met.get('dram__bytes_write.sum', 0.0)

    pending = []                                     # MN weight-gradient kernels waiting for their reduce kernel (tells the group)
    for kid in order:
        name = names[kid]
        if 'gemm_tn_mn_kernel' in name:
            pending.append(kid)
            continue
        if 'mn_reduce_kernel' in name:
            g = 'nt_gemm_tn_centered' if 'double' in name else 'nt_gemm_tn'
            for q in pending:
                add(g, q)                            # one library call = one MN kernel + its reduce kernel: counted as ONE launch
            pending = []
            add(g, kid, count=0)
            continue
        g = group_of(name)
        if g is not None:
            add(g, kid)
    res = {}
    for g, o in sorted(out.items(), key=lambda kv: -kv[1]['time_ns']):
        n = max(o['launches'], 1)
        res[g] = {'launches_per_step': o['launches'] / steps, 'dram_bytes_per_launch': o['dram_bytes'] / n,
                  'avg_launch_us_under_ncu': o['time_ns'] / n / 1e3}
    res['_source'] = {'file': path, 'steps_in_capture': steps,
                      'how': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; '
                             'dram bytes = read + write, averaged over the launches of the group (cold L2 per launch); a weight-gradient '
                             'call = its MN kernel + its reduce kernel'}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
