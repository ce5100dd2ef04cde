#!/bin/bash
# Schwarz default: parity file, fill overlap experiments, final ncu captures of the class engine
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^ \|^$" gpurun_out/r2i_pytest.log | tail -n 6
run() { tag=$1; w=$2; shift; shift; env "$@" timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_${w}_$tag.json 2> gpurun_out/r2i_bench_${w}_$tag.err
  python - gpurun_out/r2i_bench_${w}_$tag.json "$w $tag" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step %.4f"%d["ms_per_step"], "serial %.3f"%d["roofline"]["serialised_launch_sum_ms"], "|", " ".join("%s %.3f" % (k["kernel"][-5:], k["ms"]) for k in d["kernels"]), "| checksum %.12f"%d["checksum"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run default h2o_64 MYQC_X=0
run regions2 h2o_64 MYQC_FILL_REGIONS=2
run regions4 h2o_64 MYQC_FILL_REGIONS=4
run screened h2o_64 MYQC_FILL_MODE=screened
run screened_nopace h2o_64 MYQC_FILL_MODE=screened MYQC_FILL_NOPACE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_h2o64_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2i_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'eri_class|fill_zero' -c 10 -o gpurun_out/r2_eri_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2i_ncu_full.log 2>&1; echo "ncu full rc=$?"
