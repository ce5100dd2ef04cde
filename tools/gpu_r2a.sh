#!/bin/bash
# round 2, first GPU pass of the owner-row strip engine: parity tests, then the headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2a_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2a_bench.json"))
    print("step %.3f ms value %.4g" % (d["ms_per_step"], d["value"]), "roofline", {k: d["roofline"][k] for k in ("bound","frac","hbm_frac","fp64_frac_measured_peak","serialised_launch_sum_ms")})
    for k in d["kernels"]: print(k["kernel"], "%.3f ms"%k["ms"], "tasks", k["tasks"], "fp64 frac %.3f"%(k["fp64_frac"] or 0), k["prim_quartets"])
    print("e2e", d["e2e"]); print("checksum", d["checksum"])
except Exception as e: print("bench parse FAILED", e)
PY
for w in h2o_16 c20h42; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; python - gpurun_out/r2a_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "checksum", d["checksum"])
    for k in d["kernels"]: print("  ", k["kernel"], "%.4f ms"%k["ms"], "fp64 frac %.3f"%(k["fp64_frac"] or 0))
except Exception as e: print("FAILED", e)
PY
done
