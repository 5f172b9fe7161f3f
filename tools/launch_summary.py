"""Developer tool: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import sys
from collections import defaultdict


def main(path, header=''):
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in csv.reader(open(path)):
        if len(r) > 5 and r[0].isdigit():
            if len(r) > 12 and r[12] != 'gpu__time_duration.sum':      # launch lists that carry several metrics per launch
                continue
            name = r[4]
            try:
                ns = float(r[-1])
            except ValueError:
                continue
            tot[name] += ns / 1000.0
            cnt[name] += 1
    total = sum(tot.values())
    if header:
        print(header)
    print('%-100s %6s %12s %7s' % ('kernel', 'count', 'total_us', 'share'))
    for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%-100s %6d %12.1f %6.1f%%' % (name[:100], cnt[name], us, 100.0 * us / total))
    print('%-100s %6d %12.1f' % ('TOTAL', sum(cnt.values()), total))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
