#!/bin/bash
# warp-cooperative (SP SP|SP SP) kernel (MYQC_PP_KERNEL=warp): parity, then the three bench configs against the slices
mkdir -p gpurun_out
MYQC_PP_KERNEL=warp timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2y_pytest_warp.log 2>&1; echo "pytest(warp) rc=$?"; tail -n 3 gpurun_out/r2y_pytest_warp.log
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_${w}_$tag.json 2> gpurun_out/r2y_bench_${w}_$tag.err
  python - gpurun_out/r2y_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| {2,2} frac %.3f"%d["kernels"][-1]["frac"], "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for w in h2o_16 c20h42 h2o_64; do run warp $w MYQC_PP_KERNEL=warp; run slices $w MYQC_PP_KERNEL=slices; done
