#!/bin/bash
# primitive split along the lane-side primitives: parity (forced), timings
mkdir -p gpurun_out
MYQC_SPLIT_FILL=1e9 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not h2o_64 and not c20h42" > gpurun_out/r2t_pytest_forced.log 2>&1; echo "pytest forced split rc=$?"; tail -n 3 gpurun_out/r2t_pytest_forced.log
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2t_bench_${w}_$tag.json 2> gpurun_out/r2t_bench_${w}_$tag.err
  python - gpurun_out/r2t_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for w in h2o_16 c20h42 h2o_8; do
run nosplit $w MYQC_SPLIT_MAXLG=0
run lg1 $w MYQC_SPLIT_MAXLG=1
run lg2 $w MYQC_SPLIT_MAXLG=2
run lg2_f4 $w MYQC_SPLIT_MAXLG=2 MYQC_SPLIT_FILL=4
run lg3 $w MYQC_SPLIT_MAXLG=3
done
