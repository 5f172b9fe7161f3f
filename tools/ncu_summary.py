"""Developer tool: per-kernel summary (selected raw metrics + top stall lines) of an .ncu-rep, run on the CPU box."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum']
STALLS = 'smsp__average_warps_issue_stalled_'


def main(path, top=0):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '')[:90])
        for k in KEYS:
            if k in d:
                print('   %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
        st = sorted(((float(d[h]), h[len(STALLS):-len('_per_issue_active.ratio')]) for h in hdr
                     if h.startswith(STALLS) and h.endswith('_per_issue_active.ratio') and d[h] not in ('', 'n/a')), reverse=True)
        print('   stalls/issue: ' + ', '.join('%s %.2f' % (n, v) for v, n in st[:7]))


if __name__ == '__main__':
    main(sys.argv[1])


def source_hot(path, kernel_index=0, top=40):
    """per CUDA source line: executed warp-instructions and stall samples (needs -lineinfo and --import-source on)."""
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--print-source', 'cuda', '--csv'], capture_output=True,
                         text=True).stdout
    # the CSV holds one block per kernel, each introduced by a "Kernel Name" row
    blocks, cur = [], None
    for r in csv.reader(raw.splitlines()):
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    b = blocks[kernel_index]
    hdr, rows = b['rows'][0], b['rows'][1:]
    isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    iline = hdr.index('#') if '#' in hdr else 0
    tot_e = sum(int(r[iex] or 0) for r in rows)
    tot_s = sum(int(r[isamp] or 0) for r in rows)
    print('==', b['name'][:100], 'inst', tot_e, 'samples', tot_s)
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][iex] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        print('%5s %5.1f%% inst %5.1f%% smp | %s' % (r[iline], 100.0 * int(r[iex] or 0) / max(tot_e, 1),
                                                 100.0 * int(r[isamp] or 0) / max(tot_s, 1), r[isrc].strip()[:110]))
