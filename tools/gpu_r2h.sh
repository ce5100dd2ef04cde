#!/bin/bash
# Schwarz skip: full GPU suite, benches with tau = default / 0 / 1e-11, timeline of (H2O)_16
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2h_pytest.log | tail -n 12
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_${w}_$tag.json 2> gpurun_out/r2h_bench_${w}_$tag.err
  python - gpurun_out/r2h_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); sw=d["roofline"]["schwarz"]
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "fp64 frac %.3f (evaluated %.3f)"%(d["whole_step"]["fp64_frac_of_measured_dfma_peak"], sw["fp64_frac_evaluated"]), "pq ref %d eval %d"%(sw["prim_quartets_reference_rule"], sw["prim_quartets_evaluated_this_rank"]), "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for w in h2o_64 h2o_16 c20h42; do
  run default $w MYQC_X=0
  run off $w MYQC_SCHWARZ_TAU=0
  run 1e-11 $w MYQC_SCHWARZ_TAU=1e-11
done
MYQC_TIMELINE=1 timeout 200 python bench.py --workload h2o_16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep "myqc timeline" | tail -14
