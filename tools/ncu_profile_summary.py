"""Turn an .ncu-rep (ncu --set full) and a launch-list CSV into the committed summaries under profiles/.
usage: ncu_profile_summary.py <report.ncu-rep | raw.csv> <launches.csv> <out_prefix>"""
import collections
import csv
import json
import subprocess
import sys

rep, launches, out = sys.argv[1:4]
# a raw-page CSV exported on the GPU box (ncu -i x.ncu-rep --page raw --csv) is accepted in place of the report
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_active.avg",
    "sm__cycles_active.max", "sm__cycles_active.min",
]
idx = {w: hdr.index(w) for w in WANT if w in hdr}
kn = hdr.index("Kernel Name")
recs = []
for r in rows[2:]:
    d = {"kernel": r[kn]}
    for w, i in idx.items():
        try:
            d[w] = float(r[i].replace(",", ""))
        except ValueError:
            d[w] = r[i]
        d[w + "__unit"] = units[i]
    recs.append(d)
json.dump(recs, open(out + "_ncu_full.json", "w"), indent=1)


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


with open(out + "_ncu_full.md", "w") as f:
    f.write("| kernel | ms | regs | warps active % | FP64 pipe % | issue active % | lanes/inst | DRAM read GB | DRAM write GB |\n|---|---|---|---|---|---|---|---|---|\n")
    traffic = {}
    for d in recs:
        t = d.get("gpu__time_duration.sum", 0.0)
        tu = d.get("gpu__time_duration.sum__unit", "ms")
        ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1.0)
        rd = to_bytes(d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_read.sum__unit", "byte"))
        wr = to_bytes(d.get("dram__bytes_write.sum", 0.0), d.get("dram__bytes_write.sum__unit", "byte"))
        traffic.setdefault(d["kernel"], []).append(rd + wr)
        f.write(f"| `{d['kernel'][:60]}` | {ms:.3f} | {d.get('launch__registers_per_thread', 0):.0f} | "
                f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                f"{d.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
                f"{d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.1f} | {rd / 1e9:.2f} | {wr / 1e9:.2f} |\n")
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out + "_dram_traffic_bytes.json", "w"), indent=1)

# launch list: mean duration and share per kernel
lr = [r for r in csv.reader(open(launches)) if len(r) > 10]
h = lr[0]
agg = collections.OrderedDict()
for r in lr[1:]:
    if r[h.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    agg.setdefault(r[h.index("Kernel Name")], []).append(float(r[h.index("Metric Value")].replace(",", "")))
unit = lr[1][h.index("Metric Unit")]
ours = {k: v for k, v in agg.items() if "myqc::" in k and "dfma_peak" not in k}
tot = sum(sum(v) for v in ours.values())
with open(out + "_launches.md", "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 2 --warmup 1`; unit {unit}; share among the engine's own kernels\n\n")
    f.write("| kernel | launches | mean | share |\n|---|---|---|---|\n")
    for k, v in agg.items():
        sh = f"{sum(v) / tot:.3f}" if k in ours else "-"
        f.write(f"| `{k[:80]}` | {len(v)} | {sum(v) / len(v):.0f} | {sh} |\n")
