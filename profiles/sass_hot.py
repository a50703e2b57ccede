"""Top SASS instructions by stall samples from `ncu --page source --csv --print-source sass`."""
import csv
import sys


def main(path, n=30):
    rows = list(csv.reader(open(path)))
    h = rows[1]; data = rows[2:]
    ix = {k: i for i, k in enumerate(h)}
    S = ix['# Samples']; E = ix['Instructions Executed']
    tot = sum(int(r[S]) for r in data)
    print("total samples", tot, "n instr", len(data))
    stallcols = [k for k in h if k.startswith('stall_') and 'Not' not in k]
    agg = {}
    for r in data:
        for k in stallcols:
            agg[k] = agg.get(k, 0) + int(r[ix[k]] or 0)
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    top = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:n]
    for i in sorted(top):
        r = data[i]
        st = sorted([(int(r[ix[k]] or 0), k) for k in stallcols], reverse=True)[:2]
        print(i, r[S], r[E], r[ix['Source']].strip()[:64], st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
