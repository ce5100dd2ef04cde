#!/bin/bash
# parity, bench, and one ncu --set full capture of every strip launch of one (H2O)_64 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2b_pytest.log | tail -n 12
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2b_bench.json"))
    print("step %.3f ms value %.4g" % (d["ms_per_step"], d["value"]), "roofline", {k: d["roofline"][k] for k in ("bound","frac","hbm_frac","fp64_frac_measured_peak","serialised_launch_sum_ms")})
    for k in d["kernels"]: print(k["kernel"], "%.3f ms"%k["ms"], "tasks", k["tasks"], "fp64 frac %.3f"%(k["fp64_frac"] or 0), k["prim_quartets"])
    print("checksum", d["checksum"])
except Exception as e: print("bench parse FAILED", e)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:eri_strip -c 9 -o gpurun_out/r2b_eri_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -n 3 gpurun_out/r2b_ncu_full.log
ls -la gpurun_out/r2b_*
