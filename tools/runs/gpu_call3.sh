#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ao2mo.py -m gpu -q --maxfail=10 > gpurun_out/c3_pytest.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/c3_pytest.log
MYQC_AO2MO_TRACE=1 AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 600 python tools/bench_ao2mo.py h2o_64 > gpurun_out/c3_ao2mo_trace.jsonl 2> gpurun_out/c3_ao2mo_trace.err; echo "trace rc=$?"
cat gpurun_out/c3_ao2mo_trace.jsonl; grep "trace" gpurun_out/c3_ao2mo_trace.err | tail -12
timeout 600 python tools/bench_ao2mo.py h2o_16 h2o_32 h2o_64 > gpurun_out/c3_ao2mo_bench.jsonl 2> gpurun_out/c3_ao2mo_bench.err; echo "bench rc=$?"; cat gpurun_out/c3_ao2mo_bench.jsonl
AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_ao2mo_launches.csv python tools/bench_ao2mo.py h2o_32 > gpurun_out/c3_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'gemm_f64|panel_' -c 8 -o gpurun_out/c3_ao2mo_full -f python tools/bench_ao2mo.py h2o_32 > gpurun_out/c3_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/c3*
