#!/bin/bash
mkdir -p gpurun_out
nproc > gpurun_out/c4_nproc.txt; lscpu | head -20 >> gpurun_out/c4_nproc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sparse or plan_api" > gpurun_out/c4_pytest.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/c4_pytest.log
for cfg in "MYQC_SPARSE_D2H=0" "MYQC_HOST_THREADS=16" "MYQC_HOST_THREADS=8" "MYQC_HOST_THREADS=4" "MYQC_HOST_THREADS=32"; do
  env $cfg MYQC_TRACE=1 timeout 300 python tools/bench_e2e.py h2o_64 3 > gpurun_out/c4_e2e.tmp 2> gpurun_out/c4_e2e.err; cat gpurun_out/c4_e2e.tmp; grep "myqc trace" gpurun_out/c4_e2e.err | tail -n 1
  cat gpurun_out/c4_e2e.tmp >> gpurun_out/c4_e2e.txt; grep "myqc trace" gpurun_out/c4_e2e.err | tail -n 2 >> gpurun_out/c4_e2e.txt
done
MYQC_AO2MO_TRACE=1 AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 600 python tools/bench_ao2mo.py h2o_64 > gpurun_out/c4_ao2mo_trace.jsonl 2> gpurun_out/c4_ao2mo_trace.err; echo "trace rc=$?"
cat gpurun_out/c4_ao2mo_trace.jsonl; grep "trace" gpurun_out/c4_ao2mo_trace.err | tail -n 6
timeout 300 python -m pytest tests/test_gpu_ao2mo.py -m gpu -q -k "transform_matches or many_panels" > gpurun_out/c4_pytest_ao2mo.log 2>&1; echo "ao2mo tests rc=$?"; tail -n 2 gpurun_out/c4_pytest_ao2mo.log
