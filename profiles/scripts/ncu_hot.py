#!/usr/bin/env python3
"""Top SASS lines of one kernel in an .ncu-rep (source page): samples and executed instructions."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[hi]
si, ss, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for i, r in enumerate(rows[hi + 1:]):
    if len(r) <= ie or r[0] == 'Address' or r[0] == 'Kernel Name':
        continue
    try:
        data.append((int(r[ss] or 0), int(r[ie] or 0), i, r[si]))
    except ValueError:
        pass
ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
print('total samples', ts, 'warp instructions', ti, 'sass lines', len(data))
print('-- by samples')
for d in sorted(data, reverse=True)[:top]:
    print(f'{d[2]:5d} samp {d[0]:6d} ({100 * d[0] / max(ts, 1):4.1f}%) inst {d[1]:10d} ({100 * d[1] / max(ti, 1):4.1f}%)  {d[3].strip()[:100]}')
print('-- by instructions')
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f'{d[2]:5d} samp {d[0]:6d} ({100 * d[0] / max(ts, 1):4.1f}%) inst {d[1]:10d} ({100 * d[1] / max(ti, 1):4.1f}%)  {d[3].strip()[:100]}')
