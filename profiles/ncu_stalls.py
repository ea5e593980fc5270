"""Top stalled SASS instructions from `ncu --page source --csv`:
python profiles/ncu_stalls.py src.csv [N]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    iS = hdr.index('Warp Stall Sampling (All Samples)')
    isrc = hdr.index('Source')
    stalls = [(h, hdr.index(h)) for h in hdr
              if h.startswith('stall_') and 'Not Issued' not in h]

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    tot = sum(num(r[iS]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    agg = {s: sum(num(r[i]) for r in data) for s, i in stalls}
    print({k: round(100 * v / tot, 1)
           for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
    order = sorted(range(len(data)), key=lambda k: -num(data[k][iS]))[:top_n]
    for k in sorted(order):
        r = data[k]
        st = {s: num(r[i]) for s, i in stalls if num(r[i]) > 0}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print('%4d %5.1f%% %-64s %s' % (k, 100 * num(r[iS]) / tot,
                                        r[isrc].strip()[:64], st))


if __name__ == '__main__':
    main()
