#!/bin/bash
# warp-cooperative (SP SP|SP SP) kernel v2 (padded records, wider reduction loop): bench + ncu capture with source view
mkdir -p gpurun_out
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2z_bench_${w}_$tag.json 2> gpurun_out/r2z_bench_${w}_$tag.err
  python - gpurun_out/r2z_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| {2,2} frac %.3f"%d["kernels"][-1]["frac"], "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for w in h2o_16 c20h42 h2o_64; do run warp $w MYQC_PP_KERNEL=warp; done
MYQC_PP_KERNEL=warp timeout 600 ncu --set full --import-source on --clock-control none -k regex:eri_ppw -c 1 -o /tmp/cap_ppw -f python bench.py --workload c20h42 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2z_ncu_ppw.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/cap_ppw.ncu-rep --page raw --csv > gpurun_out/r2z_ppw_raw.csv 2>/dev/null
ncu -i /tmp/cap_ppw.ncu-rep --page source --csv > gpurun_out/r2z_ppw_source.csv 2>/dev/null
ls -la gpurun_out/r2z_ppw_*.csv
