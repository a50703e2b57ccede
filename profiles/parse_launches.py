"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one forward batch."""
import csv
import sys


def short(n):
    if "stem_tc" in n: return "stemTC"
    if "stem" in n: return "stem"
    if "head" in n: return "head"
    if "conv_halo" in n: return "halo" + n[n.index("conv_halo_kernel<") + 16:n.index(">") + 1]
    if "conv_gemm" in n: return "gemm" + n[n.index("conv_gemm_kernel<") + 16:n.index(">") + 1]
    return n[:24]


def main(path, per_batch=34):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = [(r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines)]
    starts = [i for i, s in enumerate(seq) if "stem" in s[0]]
    i0 = starts[1] if len(starts) > 1 else starts[0]
    batch = seq[i0:i0 + per_batch]
    tot = sum(t for _, t in batch)
    print(" ".join("%s:%.0f" % (short(n), t / 1000) for n, t in batch))
    print("total us %.0f" % (tot / 1000))


if __name__ == "__main__":
    main(sys.argv[1])
