#!/bin/bash
# merged parts (one launch per class and shard): parity, N=1 timings unchanged?
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2u_pytest.log
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2u_bench_${w}_$tag.json 2> gpurun_out/r2u_bench_${w}_$tag.err
  python - gpurun_out/r2u_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run merged h2o_64 MYQC_X=0
run merged h2o_16 MYQC_X=0
run merged c20h42 MYQC_X=0
