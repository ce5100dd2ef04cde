#!/bin/bash
# round-2 end-style run on ONE GPU: full GPU test suite, smoke, both bench arms, ncu evidence for the three bench configs
mkdir -p gpurun_out
P=gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 ${P}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 ${P}_smoke.log
timeout 600 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"; head -c 1500 ${P}_bench.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err; echo "ref rc=$?"; head -c 600 ${P}_bench_reference.json; echo
for w in h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 5 > ${P}_bench_$w.json 2> ${P}_bench_$w.err; python - ${P}_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "e2e", d["e2e"].get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("FAILED", e)
PY
done
MYQC_OUTPUT_MODE=compose timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > ${P}_bench_h2o_64_compose.json 2> ${P}_bench_h2o_64_compose.err; echo "compose bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_h2o64_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > ${P}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eri_class|fill_zero' -c 10 -o ${P}_h2o64_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > ${P}_ncu_full.log 2>&1; echo "ncu full rc=$?"
MYQC_OUTPUT_MODE=compose timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eri_class|compose' -c 10 -o ${P}_h2o64_compose_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > ${P}_ncu_compose_full.log 2>&1; echo "ncu compose full rc=$?"
for w in h2o_16 c20h42; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_${w}_launches.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > ${P}_ncu_launches_$w.log 2>&1; echo "ncu launches $w rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eri_class|fill_zero' -c 10 -o ${P}_${w}_full -f python bench.py --workload $w --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > ${P}_ncu_full_$w.log 2>&1; echo "ncu full $w rc=$?"
done
ls -la ${P}_* | head -40
