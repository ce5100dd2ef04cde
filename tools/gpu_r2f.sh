#!/bin/bash
# lane-per-primitive mode of the (S SP|SP SP) / (SP SP|SP SP) kernels: parity, benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2f_pytest.log | tail -n 8
for w in h2o_64 h2o_16 c20h42 CO2; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; python - gpurun_out/r2f_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "|", " ".join("%s %.3f(%.2f)" % (k["kernel"][-5:], k["ms"], k["frac"] or 0) for k in d["kernels"]), "| checksum", d["checksum"])
except Exception as e: print("FAILED", e)
PY
done
