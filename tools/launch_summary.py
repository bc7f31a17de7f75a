#!/usr/bin/env python
"""Per-step summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches and serialised
cold-cache microseconds per kernel, per step.   python tools/launch_summary.py launches.csv <steps>"""
import collections
import csv
import sys


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == 'ID':
            hdr, start = r, i + 1
            break
    idx = {h: j for j, h in enumerate(hdr)}
    per = collections.OrderedDict()
    tot, n = 0.0, 0
    for r in rows[start:]:
        if len(r) < len(hdr) or r[idx['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
        e = per.setdefault(r[idx['Kernel Name']], [0, 0.0])
        e[0] += 1
        e[1] += v
        tot += v
        n += 1
    ours = sum(c for k, (c, t) in per.items() if 'incagg::' in k)
    print(f"# {path}: {n} launches over {steps} steps = {n / steps:.1f} per step ({ours / steps:.1f} of this library, "
          f"{(n - ours) / steps:.1f} ATen), {tot / steps:.1f} us per step serialised (cold-cache ncu times: compare "
          f"shares, not absolutes)")
    print(f"{'launches/step':>14} {'us/step':>9} {'share':>6}  kernel")
    for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{c / steps:14.1f} {t / steps:9.1f} {100 * t / tot:5.1f}%  {k[:120]}")


if __name__ == "__main__":
    main()
