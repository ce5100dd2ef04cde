#!/bin/bash
# primitive split among lanes for small launches: parity (default + forced split on every launch), timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2r_pytest.log
MYQC_SPLIT_FILL=1e9 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not h2o_64 and not c20h42" > gpurun_out/r2r_pytest_forced.log 2>&1; echo "pytest forced split rc=$?"; tail -n 3 gpurun_out/r2r_pytest_forced.log
MYQC_SPLIT_FILL=1e9 MYQC_OUTPUT_MODE=compose timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "CO2 or h2o_8 or shard or c4h10" > gpurun_out/r2r_pytest_forced_compose.log 2>&1; echo "pytest forced split compose rc=$?"; tail -n 3 gpurun_out/r2r_pytest_forced_compose.log
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench_${w}_$tag.json 2> gpurun_out/r2r_bench_${w}_$tag.err
  python - gpurun_out/r2r_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for w in h2o_16 c20h42 CO2 h2o_8; do
run nosplit $w MYQC_SPLIT_MAXLG=0
run split $w MYQC_X=0
run split_f2 $w MYQC_SPLIT_FILL=2
run split_f05 $w MYQC_SPLIT_FILL=0.5
done
run split h2o_64 MYQC_X=0
