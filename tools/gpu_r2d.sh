#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2d_pytest.log | tail -n 12
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_$tag.json 2> gpurun_out/r2d_bench_$tag.err
  python - gpurun_out/r2d_bench_$tag.json "$tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "step %.3f ms" % d["ms_per_step"], "serial %.2f" % d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%.2f(%d)" % (k["ms"], k["tasks"]) for k in d["kernels"]), "| checksum", d["checksum"])
except Exception as e: print(sys.argv[2], "bench parse FAILED", e)
PY
}
run default MYQC_X=0
run l512 MYQC_TASK_ITEMS=512
run l2048 MYQC_TASK_ITEMS=2048
run l4096 MYQC_TASK_ITEMS=4096
run h160 MYQC_TASK_ITEMS_HEAVY=160
run h640 MYQC_TASK_ITEMS_HEAVY=640
run h1280 MYQC_TASK_ITEMS_HEAVY=1280
for w in h2o_16 c20h42; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err; python - gpurun_out/r2d_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "|", " ".join("%.3f" % k["ms"] for k in d["kernels"]))
except Exception as e: print("FAILED", e)
PY
done
