#!/usr/bin/env python3
"""Per-source-line executed instructions and stall samples from
`ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --print-source cuda,sass`."""
import collections
import csv
import sys


def main(path, thresh=0.012):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    ie, si, ln = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Line No')
    agg, src, cur = collections.OrderedDict(), {}, None
    for r in data:
        if r[ln].strip().isdigit():
            cur = int(r[ln])
            src[cur] = r[1].strip()[:100]
            continue
        try:
            e, s = int(r[ie]), int(r[si])
        except ValueError:
            continue
        a = agg.setdefault(cur, [0, 0])
        a[0] += e
        a[1] += s
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    print("total warp-instructions", tot, "stall samples", tots)
    for l, (e, s) in sorted(agg.items(), key=lambda x: (x[0] is None, x[0])):
        if e / tot > thresh or s / tots > thresh:
            print(f"{str(l):>5} {100 * e / tot:5.1f}% instr {100 * s / tots:5.1f}% stall  {src.get(l, '')}")


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.012)
