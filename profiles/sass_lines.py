#!/usr/bin/env python3
"""Per-source-line executed instructions and stall samples of one kernel, from the SASS-level `ncu --page source --csv`
export joined with the line table of the same build:

    cuobjdump -xelf all kmerseek_b200/libkmerseek_b200.so      # -> dense.sm_100a.cubin ...
    nvdisasm -g -c dense.sm_100a.cubin > dense.sass
    python profiles/sass_lines.py source.csv dense.sass dense_bucket_kernelILb0 'dense_bucket_kernel<(bool)0>' [thresh]

The two listings hold the same instruction sequence (the .so that ran is the one disassembled); instructions are joined
by their index inside the kernel and checked by opcode."""
import collections
import csv
import re
import sys


def sass_lines(path, mangled_part):
    """[(opcode, file, line)] per instruction of the first .text section whose name contains mangled_part"""
    out, on, cur = [], False, (None, None)
    for l in open(path):
        if l.startswith('.text.'):
            if on:
                break
            on = mangled_part in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
        if m:
            ins = m.group(1).strip()
            ins = re.sub(r'^@!?U?P\d+\s+', '', ins)
            out.append((ins.split()[0].split('.')[0], cur[0], cur[1]))
    return out


def ncu_rows(path, kernel_part):
    rows, on, hdr = [], False, None
    for r in csv.reader(open(path)):
        if r and r[0] == 'Kernel Name':
            if on and rows:
                break
            on = kernel_part in r[1]
            hdr = None
            continue
        if not on:
            continue
        if hdr is None:
            hdr = r
            continue
        rows.append(r)
    return hdr, rows


def main(csv_path, sass_path, mangled_part, kernel_part, thresh=0.012, src_root='kmerseek_b200/csrc/'):
    lines = sass_lines(sass_path, mangled_part)
    hdr, rows = ncu_rows(csv_path, kernel_part)
    ie, si, so = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source')
    if len(lines) != len(rows):
        print(f'warning: {len(lines)} instructions in the listing, {len(rows)} in the capture', file=sys.stderr)
    agg = collections.OrderedDict()
    bad = 0
    for (op, f, ln), r in zip(lines, rows):
        ins = re.sub(r'^@!?U?P\d+\s+', '', r[so].strip())
        if ins.split()[0].split('.')[0] != op:
            bad += 1
        a = agg.setdefault((f, ln), [0, 0])
        a[0] += int(r[ie] or 0)
        a[1] += int(r[si] or 0)
    if bad:
        print(f'warning: {bad} opcode mismatches', file=sys.stderr)
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    print('total warp-instructions', tot, 'stall samples', tots)
    cache = {}
    for (f, ln), (e, s) in sorted(agg.items(), key=lambda x: (str(x[0][0]), x[0][1] or 0)):
        if e / tot > thresh or s / tots > thresh:
            text = ''
            if f:
                try:
                    cache.setdefault(f, open(src_root + f).read().splitlines())
                    text = cache[f][ln - 1].strip()[:96]
                except (OSError, IndexError):
                    pass
            print(f'{str(f):>18}:{str(ln):<5} {100 * e / tot:5.1f}% instr {100 * s / tots:5.1f}% stall  {text}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]) if len(sys.argv) > 5 else 0.012)
