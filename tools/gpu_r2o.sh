#!/bin/bash
# does a streaming zero fill co-run with the class kernels once they only write the staging array?
mkdir -p gpurun_out
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2o_bench_${w}_$tag.json 2> gpurun_out/r2o_bench_${w}_$tag.err
  python - gpurun_out/r2o_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run fill_and_class h2o_64 MYQC_EXP_CORUN=1
run fill_only h2o_64 MYQC_EXP_CORUN=2
run class_only h2o_64 MYQC_EXP_CORUN=3
run fill1_and_class h2o_64 MYQC_EXP_CORUN=1 MYQC_FILL_CTAS=1
run fill4_and_class h2o_64 MYQC_EXP_CORUN=1 MYQC_FILL_CTAS=4
