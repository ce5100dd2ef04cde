#!/bin/bash
# adaptive task size + one stream per launch; new parity tests; compute-sanitizer on the small examples
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=8 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2g_pytest.log | tail -n 22
for w in h2o_64 h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2g_bench_$w.json 2> gpurun_out/r2g_bench_$w.err; python - gpurun_out/r2g_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "|", " ".join("%s %.3f(%.2f)" % (k["kernel"][-5:], k["ms"], k["frac"] or 0) for k in d["kernels"]), "| checksum", d["checksum"])
except Exception as e: print("FAILED", e)
PY
done
# compute-sanitizer on the drop-in int2e program (CO2 and (H2O)_4 job directories)
python - <<'PY'
import os, sys
sys.path.insert(0, "tests")
import myqc_b200 as Q
from conftest import example_zmat, INPUTS
for name in ("CO2", "h2o_4"):
    d = os.path.join("gpurun_out", "san_" + name)
    os.makedirs(d, exist_ok=True)
    for f in ("XX", "error"):
        if os.path.exists(os.path.join(d, f)): os.remove(os.path.join(d, f))
    Q.make_job(d, example_zmat(name), INPUTS)
PY
for name in CO2 h2o_4; do
  for tool in memcheck racecheck; do
    rm -f gpurun_out/san_$name/XX
    ( cd gpurun_out/san_$name && timeout 600 compute-sanitizer --tool $tool --print-limit 20 ../../myqc_b200/csrc/int2e > ../r2_sanitizer_${tool}_$name.log 2>&1 ); echo "$tool $name rc=$?"; tail -n 3 gpurun_out/r2_sanitizer_${tool}_$name.log
  done
done
rm -rf gpurun_out/san_CO2 gpurun_out/san_h2o_4
