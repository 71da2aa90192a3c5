"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 2 --warmup 1`:
shares per kernel for one resident step and one end-to-end step (steps are delimited by the k_fill_i32
launches of sw_rewind / sw_reset).  usage: launch_summary.py launches.csv out.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
seq = [(r[ix['Kernel Name']].split('(')[0].replace('void ', ''), float(r[ix['Metric Value']].replace(',', '')), r[ix['Grid Size']])
       for r in rows[1:]]
fills = [i for i, (k, v, g) in enumerate(seq) if k == 'k_fill_i32']
groups = []
for i in fills:
    if groups and i == groups[-1][-1] + 1: groups[-1].append(i)
    else: groups.append([i])
out = ['# Launch list of `python bench.py --steps 2 --warmup 1` (64 members x 1M events, K=65536), round 1 final', '',
       '`ncu --metrics gpu__time_duration.sum --clock-control none -c 1500` (raw list: r01c_launches_c3.csv).  Per-launch times',
       'under ncu are cold-cache and SERIALISED (in the end-to-end step the can_see scan normally runs beside the round',
       'kernel, on the copy stream; here it is timed alone): read the shares, the absolute step time is bench.py\'s.', '']


def table(title, lo, hi):
    a = collections.OrderedDict()
    for k, v, g in seq[lo:hi]:
        a.setdefault(k, [0, 0.0]); a[k][0] += 1; a[k][1] += v
    tot = sum(v[1] for v in a.values())
    out.append('## %s (launches %d..%d): %d launches, %.3f ms of kernel time\n' % (title, lo, hi - 1, hi - lo, tot / 1e6))
    out.append('| kernel | launches | total us | share |\n|---|---|---|---|')
    for k, (n, t) in sorted(a.items(), key=lambda x: -x[1][1]):
        out.append('| %s | %d | %.1f | %.1f%% |' % (k, n, t / 1e3, 100 * t / tot))
    out.append('')


# resident steps come first (warm-up, 2 timed, the find_order pass), each after a group of rewind fills; the
# end-to-end steps follow, each after a group of reset fills
res = [g for g in groups if len(g) <= 6]
first_big = [i for i, (k, v, g) in enumerate(seq) if 'k_cs_local<2, 1>' in k]
table('resident step (timed step 1)', groups[1][-1] + 1, groups[2][0])
# an end-to-end step: the last interval between two fill groups that holds no find_order kernel
for a, b in reversed(list(zip(groups[:-1], groups[1:]))):
    names = {k for k, v, g in seq[a[-1] + 1:b[0]]}
    if 'k_rounds_batch<2, 1>' in names and 'k_order_rounds' not in names and b[0] - a[-1] > 150:
        table('end-to-end step (16 chunks; appends two chunks ahead)', a[-1] + 1, b[0])
        break
open(sys.argv[2], 'w').write('\n'.join(out))
print('\n'.join(out))
