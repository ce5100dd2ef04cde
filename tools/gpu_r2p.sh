#!/bin/bash
# what bounds the compose pass: gathers that hit L2 (2), no gathers at all (3), zeros only (1)
mkdir -p gpurun_out
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_${w}_$tag.json 2> gpurun_out/r2p_bench_${w}_$tag.err
  python - gpurun_out/r2p_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], [k["ms"] for k in d["kernels"] if k["kernel"]=="compose"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run s4_full h2o_64 MYQC_COMPOSE_SUB=4 MYQC_COMPOSE_CTAS=6
run s4_l2hot h2o_64 MYQC_COMPOSE_SUB=4 MYQC_COMPOSE_CTAS=6 MYQC_COMPOSE_ZERO=2
run s4_noload h2o_64 MYQC_COMPOSE_SUB=4 MYQC_COMPOSE_CTAS=6 MYQC_COMPOSE_ZERO=3
run s4_zero h2o_64 MYQC_COMPOSE_SUB=4 MYQC_COMPOSE_CTAS=6 MYQC_COMPOSE_ZERO=1
run s8_l2hot h2o_64 MYQC_COMPOSE_SUB=8 MYQC_COMPOSE_ZERO=2
run s8_noload h2o_64 MYQC_COMPOSE_SUB=8 MYQC_COMPOSE_ZERO=3
