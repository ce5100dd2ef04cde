#!/bin/bash
# final evidence of round 2, session 3 (one B200): all GPU tests, smoke, sanitizer runs of the warp-cooperative kernels,
# the three bench lines, ncu launch lists and full captures (raw pages exported to CSV on the box)
mkdir -p gpurun_out
P=gpurun_out/r2g
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 2 ${P}_pytest_gpu.log
MYQC_PP_KERNEL=slices MYQC_SP_KERNEL=class timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > ${P}_pytest_class_kernels.log 2>&1; echo "pytest(class kernels forced) rc=$?"; tail -n 2 ${P}_pytest_class_kernels.log
timeout 300 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 ${P}_smoke.log
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
    ( cd gpurun_out/san_$name && MYQC_PP_KERNEL=warp MYQC_SP_KERNEL=warp timeout 600 compute-sanitizer --tool $tool --print-limit 20 ../../myqc_b200/csrc/int2e > ../r2g_sanitizer_${tool}_${name}_warp_kernels.log 2>&1 ); echo "$tool $name rc=$?"; tail -n 2 ${P}_sanitizer_${tool}_${name}_warp_kernels.log
  done
done
rm -rf gpurun_out/san_CO2 gpurun_out/san_h2o_4
timeout 600 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
for w in h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 5 > ${P}_bench_$w.json 2> ${P}_bench_$w.err; done
for w in h2o_64 h2o_16 c20h42; do python - ${P}_bench$( [ $w = h2o_64 ] || echo _$w ).json $w <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "e2e %.1f ms"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
for w in h2o_64 h2o_16 c20h42; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_${w}_launches.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > ${P}_ncu_launches_$w.log 2>&1; echo "ncu launches $w rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eri_class|eri_ppw|fill_zero' -c 10 -o /tmp/cap_$w -f python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > ${P}_ncu_full_$w.log 2>&1; echo "ncu full $w rc=$?"
ncu -i /tmp/cap_$w.ncu-rep --page raw --csv > ${P}_${w}_raw.csv 2>/dev/null
rm -f /tmp/cap_$w.ncu-rep
done
du -sh gpurun_out
