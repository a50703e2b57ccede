"""Per-launch table (time, DRAM traffic, tensor-pipe utilisation) of one forward batch from an
`ncu --set full` report:  python profiles/summarize_full.py <report.ncu-rep> <instances> <out.md> <out.json>"""
import csv
import io
import json
import subprocess
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,"
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,"
           "sm__cycles_elapsed.max,lts__t_bytes.sum")


def main(rep, instances, out_md, out_json):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", METRICS],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units, data = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(h)}

    def val(r, k, unit_scale=None):
        v = float(r[ix[k]].replace(",", ""))
        u = units[ix[k]]
        if u == "Mbyte": v *= 1e6
        elif u == "Gbyte": v *= 1e9
        elif u == "Kbyte": v *= 1e3
        elif u == "ms": v *= 1e3          # -> us
        elif u == "ns": v *= 1e-3
        return v

    lines = ["| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | DRAM TB/s | tensor pipe % | smem wavefronts lsu / tc (M) |",
             "|---|---|---|---|---|---|---|---|---|"]
    tot = {"us": 0.0, "rd": 0.0, "wr": 0.0}
    for i, r in enumerate(data):
        name = r[ix["Kernel Name"]].replace("void unnamed>::", "").replace("unnamed>::", "").split("(")[0]
        us = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        tp = val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")
        lsu = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / 1e6
        tc = val(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum") / 1e6
        tot["us"] += us; tot["rd"] += rd; tot["wr"] += wr
        lines.append("| %d | %s | %s | %.1f | %.1f | %.1f | %.2f | %.1f | %.1f / %.1f |" % (
            i, name, r[ix["Grid Size"]].strip("()").split(",")[0], us, rd / 1e6, wr / 1e6,
            (rd + wr) / us / 1e6, tp, lsu, tc))
    lines.append("| | **total** | | %.1f | %.1f | %.1f | %.2f | | |" % (
        tot["us"], tot["rd"] / 1e6, tot["wr"] / 1e6, (tot["rd"] + tot["wr"]) / tot["us"] / 1e6))
    open(out_md, "w").write("\n".join(lines) + "\n")
    json.dump({"source": rep.split("/")[-1], "instances": int(instances), "launches": len(data),
               "sum_kernel_us": tot["us"], "dram_read_bytes": tot["rd"], "dram_write_bytes": tot["wr"],
               "dram_bytes_per_instance": (tot["rd"] + tot["wr"]) / int(instances)}, open(out_json, "w"), indent=1)
    print("\n".join(lines[-3:]))


if __name__ == "__main__":
    main(*sys.argv[1:5])
