#!/bin/bash
# one GPU call: full gpu test suite, ao2mo cross-check + bench, ERI store-hint variants, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x --deselect tests/test_gpu_ao2mo.py > gpurun_out/c1_pytest_core.log 2>&1; echo "core tests rc=$?"
timeout 600 python -m pytest tests/test_gpu_ao2mo.py -m gpu -q --maxfail=30 > gpurun_out/c1_pytest_ao2mo.log 2>&1; echo "ao2mo tests rc=$?"
MYQC_AO2MO_GEMM=simt timeout 600 python -m pytest tests/test_gpu_ao2mo.py -m gpu -q --maxfail=30 > gpurun_out/c1_pytest_ao2mo_simt.log 2>&1; echo "ao2mo simt tests rc=$?"
tail -3 gpurun_out/c1_pytest_core.log gpurun_out/c1_pytest_ao2mo.log gpurun_out/c1_pytest_ao2mo_simt.log
timeout 600 python tools/bench_ao2mo.py h2o_16 h2o_32 h2o_64 > gpurun_out/c1_ao2mo_bench.jsonl 2> gpurun_out/c1_ao2mo_bench.err; echo "ao2mo bench rc=$?"; cat gpurun_out/c1_ao2mo_bench.jsonl; tail -3 gpurun_out/c1_ao2mo_bench.err
timeout 300 python tools/bench_fock.py h2o_64 3 > gpurun_out/c1_fock_bench.txt 2>&1; cat gpurun_out/c1_fock_bench.txt | tail -3
for v in "" stcs stcg stwt; do
  if [ -n "$v" ]; then export MYQC_LIB=$PWD/myqc_b200/csrc/variants/libmyqc_eri_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_var_${v:-base}.json 2> gpurun_out/c1_var_${v:-base}.err
  python - "${v:-base}" gpurun_out/c1_var_${v:-base}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "| step %.2f ms |"%d["ms_per_step"], " ".join("%s %.2f"%(k["kernel"].replace("eri_class",""),k["ms"]) for k in d["kernels"]), "| checksum", d["checksum"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
unset MYQC_LIB
timeout 600 python bench.py > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err; echo "bench rc=$?"; cat gpurun_out/c1_bench_default.json | head -c 3000
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c1_bench_reference.json 2> gpurun_out/c1_bench_reference.err; echo "ref rc=$?"; head -c 1500 gpurun_out/c1_bench_reference.json
