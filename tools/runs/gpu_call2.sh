#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fock.py tests/test_gpu_ao2mo.py -m gpu -q --maxfail=10 > gpurun_out/c2_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c2_pytest.log
MYQC_AO2MO_TRACE=1 AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 600 python tools/bench_ao2mo.py h2o_32 h2o_64 > gpurun_out/c2_ao2mo_trace.jsonl 2> gpurun_out/c2_ao2mo_trace.err; echo "trace rc=$?"
cat gpurun_out/c2_ao2mo_trace.jsonl; grep "trace" gpurun_out/c2_ao2mo_trace.err | tail -48
AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_f64 -c 4 -o gpurun_out/c2_ao2mo_gemm -f python tools/bench_ao2mo.py h2o_32 > gpurun_out/c2_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
AO2MO_BENCH_KINDS=mma AO2MO_BENCH_REPS=1 timeout 900 ncu --set full --clock-control none -k regex:unpack -c 3 -o gpurun_out/c2_ao2mo_unpack -f python tools/bench_ao2mo.py h2o_32 > gpurun_out/c2_ncu_unpack.log 2>&1; echo "ncu unpack rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 400 python tools/bench_fock.py h2o_64 3 > gpurun_out/c2_fock_bench.txt 2>&1; tail -8 gpurun_out/c2_fock_bench.txt
