"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share per kernel, and the sequence.
usage: python tools/launch_summary.py FILE [--seq N]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg, seq = collections.OrderedDict(), []
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    k = r[ki][:70]
    agg.setdefault(k, [0, 0]); agg[k][0] += v; agg[k][1] += 1
    seq.append((k, v, r[gi]))
tot = sum(a[0] for a in agg.values())
print(f"total {tot / 1e3:.1f} us over {len(seq)} launches")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    print('%6.2f%% %10.1f us %4d  %8.1f avg  %s' % (100 * a[0] / tot, a[0] / 1e3, a[1], a[0] / a[1] / 1e3, k))
if '--seq' in sys.argv:
    n = int(sys.argv[sys.argv.index('--seq') + 1])
    for s in seq[:n]:
        print('%10.1f us  %-18s %s' % (s[1] / 1e3, s[2], s[0]))
