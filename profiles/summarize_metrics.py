"""Per-launch table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --csv` log:
  python profiles/summarize_metrics.py <log.csv> <launches per batch> <instances> <flop per instance> > out.md
(the LAST `launches per batch` rows = one warm forward batch)."""
import collections
import csv
import sys


def main(path, per_batch, instances, flop):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    by = collections.OrderedDict()
    for r in rows:
        by.setdefault((r["ID"], r["Kernel Name"], r["Grid Size"]), {})[r["Metric Name"]] = (
            float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    items = list(by.items())[-per_batch:]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print("| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM TB/s | tensor pipe % |")
    print("|---|---|---|---|---|---|---|---|")
    tot_us = tot_b = w = 0.0
    for i, ((_, name, grid), m) in enumerate(items):
        t = m["gpu__time_duration.sum"]
        us = t[0] / 1000 if t[1] == "ns" else t[0]
        rd = m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
        wr = m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
        tp = m.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", (0.0, ""))[0]
        n = name.replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("cs::", "").split("(")[0]
        print("| %d | %s | %s | %.1f | %.1f | %.1f | %.2f | %.1f |" % (i, n[:48], grid.strip("()").split(",")[0], us, rd / 1e6,
                                                                  wr / 1e6, (rd + wr) / us / 1e6, tp))
        tot_us += us; tot_b += rd + wr; w += us * tp
    print("| | **total** | | %.1f | | | %.2f | %.1f (time-weighted) |" % (tot_us, tot_b / tot_us / 1e6, w / tot_us))
    print("\n%d instances: %.1f KB of DRAM traffic per instance; %.0f TFLOP/s on in-bounds FLOP over the summed "
          "(serialised, cold-cache) kernel time." % (instances, tot_b / instances / 1e3, flop * instances / tot_us / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]))
