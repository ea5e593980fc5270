"""Aggregate an `ncu --page source --csv` (SASS view) dump by CUDA source line.

    python profiles/ncu_by_line.py src.csv file.cubin kernel_substring [top_n]

The SASS rows of ncu are matched, in order, with `nvdisasm -g` output of the same
cubin (built with -lineinfo), whose `//## File "...", line N` markers give the
source line of every instruction (innermost inlined frame).
"""
import csv
import re
import subprocess
import sys


def sass_lines(cubin, kernel):
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True,
                         text=True).stdout.splitlines()
    res, cur, active = [], ('?', 0), False
    for ln in out:
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            active = kernel in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln):
            res.append(cur)
    return res


def main():
    src_csv, cubin, kernel = sys.argv[1:4]
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    lines = sass_lines(cubin, kernel)
    if len(lines) != len(data):
        print('warning: %d SASS rows in ncu, %d in nvdisasm' % (len(data), len(lines)))
    iS = hdr.index('Warp Stall Sampling (All Samples)')
    iI = hdr.index('Instructions Executed')
    stalls = [(h, hdr.index(h)) for h in hdr
              if h.startswith('stall_') and 'Not Issued' not in h]

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    agg = {}
    for r, key in zip(data, lines):
        a = agg.setdefault(key, [0.0, 0.0, {}])
        a[0] += num(r[iS])
        a[1] += num(r[iI])
        for s, i in stalls:
            v = num(r[i])
            if v:
                a[2][s] = a[2].get(s, 0) + v
    totS = sum(a[0] for a in agg.values()) or 1
    totI = sum(a[1] for a in agg.values()) or 1
    print('samples %d, warp instructions %d' % (totS, totI))
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        st = sorted(a[2].items(), key=lambda x: -x[1])[:3]
        print('%5.1f%% smp %5.1f%% ins  %s:%d  %s' % (
            100 * a[0] / totS, 100 * a[1] / totI, key[0], key[1],
            {k.replace('stall_', ''): int(v) for k, v in st}))


if __name__ == '__main__':
    main()
