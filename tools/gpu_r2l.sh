#!/bin/bash
# compose kernel: per-column branch, batch/CTA sweep, ncu full capture of compose_kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "CO2 or h2o_8 or shard or c4h10" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2l_pytest.log
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2l_bench_${w}_$tag.json 2> gpurun_out/r2l_bench_${w}_$tag.err
  python - gpurun_out/r2l_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run b2 h2o_64 MYQC_COMPOSE_BATCH=2
run b4 h2o_64 MYQC_COMPOSE_BATCH=4
run b4c4 h2o_64 MYQC_COMPOSE_BATCH=4 MYQC_COMPOSE_CTAS=4
run b4c6 h2o_64 MYQC_COMPOSE_BATCH=4 MYQC_COMPOSE_CTAS=6
run b1c12 h2o_64 MYQC_COMPOSE_BATCH=1 MYQC_COMPOSE_CTAS=12
MYQC_COMPOSE_BATCH=4 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'compose' -c 1 -o gpurun_out/r2l_compose_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2l_ncu_full.log 2>&1; echo "ncu full rc=$?"
