"""Summaries of ncu captures for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep  [kernel-substring]  -> JSON of the selected metrics
  python tools/ncu_summary.py launches gpurun_out/launches.csv                    -> markdown launch list (share of time)
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def full(path, pattern=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
    names, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = dict(zip(names, r))
        if pattern and pattern not in rec.get("Kernel Name", ""):
            continue
        m = {k: {"unit": units[names.index(k)], "value": rec[k]} for k in KEEP if k in rec}

        def nbytes(k):
            return float(m[k]["value"].replace(",", "")) * UNIT.get(m[k]["unit"], 1.0) if k in m else None
        rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
        out.append({"kernel": rec.get("Kernel Name"), "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None, "metrics": m})
    print(json.dumps(out[0] if len(out) == 1 else out, indent=1))


def launches(path):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    agg = collections.OrderedDict()
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
        a = agg.setdefault(r["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print(f"| {k[:110]} | {a[0]} | {a[1]:.3f} | {a[1] / tot * 100:.3f}% |")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        launches(sys.argv[2])
