#!/usr/bin/env python
"""Where the stall samples of one kernel sit: ncu -i X --page source --csv --print-source sass, top lines.
Usage: ncu_stalls.py file.ncu-rep [min_pct]"""
import csv, subprocess, sys
from collections import Counter


def main(rep, min_pct=1.0):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print('kernel:', rows[0][1][:120])
    body = [r for r in rows[2:] if len(r) > 5 and r[2].isdigit() and r[5].isdigit()]
    tot_s = sum(int(r[2]) for r in body)
    tot_i = sum(int(r[5]) for r in body)
    print(f'stall samples {tot_s}, warp instructions executed {tot_i}, SASS lines {len(body)}')
    print('-- lines with >= %.1f %% of the samples: index, address, %% samples, times executed, instruction' % min_pct)
    for i, r in enumerate(body):
        if int(r[2]) >= tot_s * min_pct / 100:
            print(f'{i:5d} {r[0][-5:]} {100 * int(r[2]) / tot_s:5.1f} {int(r[5]):9d}  {r[1].strip()[:100]}')
    mix = Counter()
    for r in body:
        s = r[1].strip()
        op = (s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
        mix[op] += int(r[5])
    print('-- executed instruction mix (% of warp instructions)')
    print(', '.join(f'{k} {100 * v / tot_i:.1f}' for k, v in mix.most_common(16)))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
