#!/bin/bash
# round-end style run: full GPU test suite, smoke, both bench arms, ncu evidence, the next-row benches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/f_smoke.log
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; head -c 2500 gpurun_out/f_bench.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "ref rc=$?"; head -c 600 gpurun_out/f_bench_reference.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eri_class|fill_zero' -c 10 -o gpurun_out/f_eri_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 400 python tools/bench_fock.py h2o_64 3 > gpurun_out/f_fock.txt 2>&1; tail -n 8 gpurun_out/f_fock.txt
timeout 600 python tools/bench_ao2mo.py h2o_16 h2o_32 h2o_64 > gpurun_out/f_ao2mo.jsonl 2> gpurun_out/f_ao2mo.err; cat gpurun_out/f_ao2mo.jsonl
for w in h2o_16 c20h42; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 5 > gpurun_out/f_bench_$w.json 2> gpurun_out/f_bench_$w.err; python - gpurun_out/f_bench_$w.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["config"]["workload"], "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "fp64 frac %.3f"%d["whole_step"]["fp64_frac_of_measured_dfma_peak"], "e2e", d["e2e"].get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e: print("FAILED", e)
PY
done
ls -la gpurun_out/f_*
