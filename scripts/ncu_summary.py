"""Summarise ncu captures into profiles/: python scripts/ncu_summary.py <rep.ncu-rep> <launches.csv> <out.md> [traffic.json workload]
Reads the raw page of the report (ncu -i ... --page raw --csv) and the launch list, writes a markdown table with the metrics
the roofline arithmetic uses, and (optionally) the per-launch DRAM traffic of each kernel into a JSON file that bench.py
quotes under roofline.traffic."""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
SHORT = {"update_umma32_kernel": "grad", "update_tile_kernel": "grad", "rollout_kernel": "rollout",
         "loss_thread_kernel": "loss_kl", "gae_scan_kernel": "process_samples", "lfb_gram": "lfb_gram",
         "update_umma64_kernel": "fvp", "update_gemm_kernel": "grad"}


def raw_rows(rep):
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, launches, out_md = sys.argv[1:4]
    hdr, units, rows = raw_rows(rep)
    name_i = hdr.index("Kernel Name")
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    lines = ["| kernel | " + " | ".join(m for m, _ in cols) + " |", "|---|" + "---|" * len(cols)]
    traffic = OrderedDict()
    for r in rows:
        lines.append("| %s | " % r[name_i][:70] + " | ".join("%s %s" % (r[i], units[i]) for _, i in cols) + " |")
        try:
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            for k, short in SHORT.items():
                if k in r[name_i]:
                    if k == "update_umma32_kernel" and ", 0>" in r[name_i]:
                        short = "loss_kl"                      # forward-only mode of the same kernel
                    if k == "update_umma32_kernel" and ", 2>" in r[name_i]:
                        short = "fvp"
                    if k == "update_umma64_kernel" and ", 1>" in r[name_i]:
                        short = "grad"
                    traffic.setdefault(short, rd + wr)
        except Exception:
            pass
    agg = OrderedDict()
    with open(launches) as f:
        rd = [l for l in f if not l.startswith("==")]
    rows2 = list(csv.reader(io.StringIO("".join(rd))))
    h2 = rows2[0]
    ni, vi = h2.index("Kernel Name"), h2.index("Metric Value")
    for r in rows2[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ni][:60], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values()) or 1.0
    lines += ["", "Launch-list shares (all kernels of the process, gpu__time_duration.sum in ns):", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append("| %s | %d | %.1f | %.3f |" % (k, n, v / 1e3, v / tot))
    open(out_md, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 5:
        tj, wl = sys.argv[4:6]
        try:
            d = json.load(open(tj))
        except Exception:
            d = {}
        d[wl] = traffic
        json.dump(d, open(tj, "w"), indent=1)
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
