"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the top launches.
usage: python tools/launch_agg.py profiles/x_launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    h = rows[hdr]
    ki, vi, gi, bi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
    out = []
    for r in rows[hdr + 1:]:
        try:
            out.append((int(r[0]), r[ki], r[gi], r[bi], float(r[vi].replace(',', '')) / 1e3))
        except (ValueError, IndexError):
            pass
    return out


if __name__ == "__main__":
    L = load(sys.argv[1])
    if len(sys.argv) > 3:
        L = [x for x in L if int(sys.argv[2]) <= x[0] <= int(sys.argv[3])]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in L:
        n = re.sub(r'<.*', '', x[1])[:60]
        agg[n][0] += 1
        agg[n][1] += x[4]
    tot = sum(v[1] for v in agg.values())
    print(f"{len(L)} launches, {tot:.1f} us")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{t:9.1f} us {c:5d} {t / tot * 100:5.1f}%  {n}")
    print("-- top launches")
    for x in sorted(L, key=lambda x: -x[4])[:40]:
        print(f"{x[0]:5d} {x[4]:8.1f} us grid{x[2]} blk{x[3]} {x[1][:110]}")
