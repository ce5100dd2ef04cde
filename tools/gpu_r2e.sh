#!/bin/bash
# class engine restored (+ grouped step B, one-shot cache, slab dense): full GPU suite, smoke, benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2e_pytest.log | tail -n 8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2e_smoke.log
MYQC_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; grep "myqc trace\] shard" gpurun_out/r2e_bench.err | tail -4
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2e_bench.json"))
    print("step %.3f ms value %.4g" % (d["ms_per_step"], d["value"]), {k: d["roofline"][k] for k in ("bound","frac","hbm_frac","fp64_frac_measured_peak","serialised_launch_sum_ms")})
    for k in d["kernels"]: print(k["kernel"], "%.3f ms"%k["ms"], "frac %.3f"%(k["frac"] or 0))
    print("family", d["roofline"]["dominant_family"]); print("e2e", d["e2e"]); print("checksum", d["checksum"])
except Exception as e: print("bench parse FAILED", e)
PY
for w in h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 3 > gpurun_out/r2e_bench_$w.json 2> gpurun_out/r2e_bench_$w.err; python - gpurun_out/r2e_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "e2e ms", d["e2e"].get("ms_per_step"), "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]))
except Exception as e: print("FAILED", e)
PY
done
