#!/bin/bash
# parity, bench at a few task sizes, one ncu --set full capture of every strip launch of one (H2O)_64 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2c_pytest.log | tail -n 12
for us in 0 200 800; do
  if [ $us = 0 ]; then unset MYQC_TASK_US; else export MYQC_TASK_US=$us; fi
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_$us.json 2> gpurun_out/r2c_bench_$us.err; echo "bench task_us=$us rc=$?"
  python - gpurun_out/r2c_bench_$us.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("step %.3f ms" % d["ms_per_step"], "serial sum %.2f" % d["roofline"]["serialised_launch_sum_ms"], "checksum", d["checksum"])
    print("  " + " | ".join("%s %.2f (%d)" % (k["kernel"].replace("eri_strip",""), k["ms"], k["tasks"]) for k in d["kernels"]))
except Exception as e: print("bench parse FAILED", e)
PY
done
unset MYQC_TASK_US
for w in h2o_16 c20h42; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err; python - gpurun_out/r2c_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "checksum", d["checksum"])
    print("  " + " | ".join("%s %.3f" % (k["kernel"].replace("eri_strip",""), k["ms"]) for k in d["kernels"]))
except Exception as e: print("FAILED", e)
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:eri_strip -c 9 -o gpurun_out/r2c_eri_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"
